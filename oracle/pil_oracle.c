/*
 * pil_oracle.c -- CPU oracle for the physics-prior loss (Stage II fine-tuning loss).
 *
 * ======================= TEST INFRASTRUCTURE -- NOT PRODUCT CODE =======================
 * A plain-C restatement of the reference algorithm (src/pde.py, src/loss.py of
 * seemapoudel58/Physics_informed_image_segmentation), used ONLY as the checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing under
 * physics_informed_image_segmentation_b200/ may import, link or call it: the product path is the
 * CUDA library behind include/pil.h and fails loudly without it.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this
 * oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build container by
 * tests/golden/make_golden.py (imports /root/reference/src/{pde,loss}.py by path) and committed as
 * tests/golden/ref_cases.npz, plus the known-answer stencil vectors of SURVEY.md section 4.
 * tests/test_oracle.py checks both (fp32 and fp64 instantiations).
 * ========================================================================================
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC) -> oracle/libpil_oracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct PiloParams {
    double dice_weight;        /* src/loss.py:88 */
    double bce_weight;         /* src/loss.py:89 */
    double pde_weight;         /* src/loss.py:90, gate at :150 */
    double phase_field_weight; /* src/loss.py:91, gate at :155 */
    double diffusion_coeff;    /* src/pde.py:9  */
    double reaction_threshold; /* src/pde.py:10 */
    double epsilon;            /* src/loss.py:95, src/pde.py:183 */
    double smooth;             /* src/loss.py:92 */
} PiloParams;

#define REAL float
#define SUF f32
#define EXP expf
#define LOG logf
#define TANH tanhf
#include "pil_oracle_impl.inc"
#undef REAL
#undef SUF
#undef EXP
#undef LOG
#undef TANH

#define REAL double
#define SUF f64
#define EXP exp
#define LOG log
#define TANH tanh
#include "pil_oracle_impl.inc"
#undef REAL
#undef SUF
#undef EXP
#undef LOG
#undef TANH

/*
 * Assemble the scalar loss from (all-reduced) sums: src/loss.py:134-160.
 *   out[0] total, out[1] dice_loss, out[2] bce, out[3] L_rd (mean r^2), out[4] L_pf
 * The gates are the reference's Python `> 0` tests on the weights.
 */
void pilo_finalize(const double* sums, int64_t n_global, const PiloParams* p, double* out) {
    const double I = sums[0], P = sums[1], T = sums[2], s = p->smooth;
    const double N = (double)n_global;
    const double dice_loss = 1.0 - (2.0 * I + s) / (P + T + s);
    const double bce = sums[3] / N, rd = sums[4] / N, pf = sums[5] / N;
    double total = p->dice_weight * dice_loss + p->bce_weight * bce;
    if (p->pde_weight > 0.0) total += p->pde_weight * rd;
    if (p->phase_field_weight > 0.0) total += p->phase_field_weight * pf;
    out[0] = total;
    out[1] = dice_loss;
    out[2] = bce;
    out[3] = rd;
    out[4] = pf;
}
