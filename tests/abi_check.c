/* Compiled (not run) by tests/test_host_cpu.py with a C compiler: include/pil.h must be a plain C header, and the
 * entry points must have exactly the signatures a C caller (or a cgo / JNI / ctypes binding) is told about. */
#include "pil.h"

typedef int (*fwd_fn)(const void*, const void*, int64_t, int64_t, int64_t, int, int, int, const PilParams*, double*, float*,
                      void*, size_t, void*);
typedef int (*bwd_fn)(const void*, const void*, void*, int64_t, int64_t, int64_t, int, int, int, const PilParams*,
                      const double*, int64_t, const float*, float, void*);
typedef int (*step_fn)(const void*, const void*, void*, int64_t, int64_t, int64_t, int, int, int, const PilParams*, double*,
                       float*, void*, size_t, void*);

int abi_check(void) {
    fwd_fn f = pil_forward;
    bwd_fn b = pil_backward;
    step_fn s = pil_loss_fwd_bwd;
    PilParams p = {0.5, 0.5, 1e-4, 1e-4, 5.0, 0.5, 0.05, 1e-6};
    PilExchange ex;
    PilLaunchInfo info;
    ex.rank = 0;
    ex.world = 1;
    ex.epoch = 0;
    ex.flags = PIL_XCHG_DEFER_FINALIZE;
    ex.mailbox[PIL_MAX_RANKS - 1] = 0;
    info.kernels_launched = 0;
    (void)info;
    (void)ex;
    return (f != 0) + (b != 0) + (s != 0) + pil_validate_params(&p) + (int)sizeof(double[PIL_NSUMS]) + (int)sizeof(float[PIL_NOUT]) +
           (int)sizeof(double[PIL_NMOMENTS]) + PIL_F32 + PIL_X_LOGITS_TANH + PIL_ERR_EXCHANGE + PIL_SESSION_GRAD_ON_DEVICE;
}
