// pil_session.cu -- host-buffer entry of include/pil.h (pil_session_*).
//
// What `loss = criterion(outputs, masks); loss.backward()` (reference src/train.py:117,:163) is to a
// caller whose maps live in HOST memory: one call moves the maps to the GPU, runs the fused forward
// and backward kernels of pil_kernels.cu and brings the loss report (and optionally the gradient)
// back.  Built only on the public C ABI plus one 8-thread kernel that adds per-chunk sums.
//
// Pipeline (2 copy streams + 1 compute stream, events in between):
//   H2D chunk c (x,t)  ->  forward(chunk c) -> ... -> add chunk sums -> finalize
//                                                  -> backward(chunk c) -> D2H grad chunk c
// The forward of chunk c overlaps the H2D of chunk c+1; the D2H of gradient chunk c overlaps the
// backward of chunk c+1.  The two phases cannot overlap each other: the Dice term makes every
// gradient depend on sums over the whole batch (reference src/loss.py:134-138).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <new>

#include "pil.h"

namespace {

constexpr int kMaxChunks = 16;

__global__ void pil_add_chunk_sums(const double* chunk_sums, int chunks, double* out) {
    const int c = threadIdx.x;
    if (c < PIL_NSUMS) {
        double v = 0.0;
        for (int k = 0; k < chunks; ++k) v += chunk_sums[k * PIL_NSUMS + c];
        out[c] = v;
    }
}

size_t dsize(int d) { return d == PIL_F32 ? 4 : (d == PIL_BF16 ? 2 : 1); }

}  // namespace

struct PilSession {
    int device;
    int64_t max_B, H, W;
    int x_dtype, t_dtype;
    void *dx, *dt, *dg;
    double* dsums;        // [kMaxChunks + 1][PIL_NSUMS]; last row = global
    float* dout;          // [PIL_NOUT]
    float* hout;          // pinned [PIL_NOUT]
    void* ws[kMaxChunks];
    size_t ws_bytes;
    cudaStream_t s_h2d, s_comp, s_d2h;
    cudaEvent_t ev_in[kMaxChunks], ev_bwd[kMaxChunks], ev_done;
};

#define PIL_TRY(expr)                          \
    do {                                       \
        cudaError_t e__ = (expr);              \
        if (e__ != cudaSuccess) return (int)e__; \
    } while (0)

extern "C" {

int pil_session_destroy(PilSession* s) {
    if (!s) return PIL_ERR_NULL;
    cudaSetDevice(s->device);
    cudaDeviceSynchronize();
    cudaFree(s->dx);
    cudaFree(s->dt);
    cudaFree(s->dg);
    cudaFree(s->dsums);
    cudaFree(s->dout);
    cudaFreeHost(s->hout);
    for (int c = 0; c < kMaxChunks; ++c) {
        cudaFree(s->ws[c]);
        if (s->ev_in[c]) cudaEventDestroy(s->ev_in[c]);
        if (s->ev_bwd[c]) cudaEventDestroy(s->ev_bwd[c]);
    }
    if (s->ev_done) cudaEventDestroy(s->ev_done);
    if (s->s_h2d) cudaStreamDestroy(s->s_h2d);
    if (s->s_comp) cudaStreamDestroy(s->s_comp);
    if (s->s_d2h) cudaStreamDestroy(s->s_d2h);
    delete s;
    return PIL_OK;
}

int pil_session_create(PilSession** out, int device, int64_t max_B, int64_t H, int64_t W, int x_dtype, int t_dtype) {
    if (!out) return PIL_ERR_NULL;
    if (max_B < 1 || H < 2 || W < 2) return PIL_ERR_SHAPE;
    if (!(x_dtype == PIL_F32 || x_dtype == PIL_BF16)) return PIL_ERR_DTYPE;
    if (!(t_dtype == PIL_F32 || t_dtype == PIL_BF16 || t_dtype == PIL_U8)) return PIL_ERR_DTYPE;
    PIL_TRY(cudaSetDevice(device));
    PilSession* s = new (std::nothrow) PilSession();
    if (!s) return (int)cudaErrorMemoryAllocation;
    *s = PilSession{};
    s->device = device;
    s->max_B = max_B;
    s->H = H;
    s->W = W;
    s->x_dtype = x_dtype;
    s->t_dtype = t_dtype;
    const size_t n = (size_t)max_B * H * W;
    s->ws_bytes = pil_workspace_bytes(max_B, H, W);
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) {
        if (e == cudaSuccess) e = r;
    };
    ok(cudaMalloc(&s->dx, n * dsize(x_dtype)));
    ok(cudaMalloc(&s->dt, n * dsize(t_dtype)));
    ok(cudaMalloc(&s->dg, n * dsize(x_dtype)));
    ok(cudaMalloc((void**)&s->dsums, sizeof(double) * PIL_NSUMS * (kMaxChunks + 1)));
    ok(cudaMalloc((void**)&s->dout, sizeof(float) * PIL_NOUT));
    ok(cudaMallocHost((void**)&s->hout, sizeof(float) * PIL_NOUT));
    ok(cudaStreamCreateWithFlags(&s->s_h2d, cudaStreamNonBlocking));
    ok(cudaStreamCreateWithFlags(&s->s_comp, cudaStreamNonBlocking));
    ok(cudaStreamCreateWithFlags(&s->s_d2h, cudaStreamNonBlocking));
    for (int c = 0; c < kMaxChunks && e == cudaSuccess; ++c) {
        ok(cudaMalloc(&s->ws[c], s->ws_bytes));
        if (e == cudaSuccess) ok(cudaMemset(s->ws[c], 0, s->ws_bytes));
        ok(cudaEventCreateWithFlags(&s->ev_in[c], cudaEventDisableTiming));
        ok(cudaEventCreateWithFlags(&s->ev_bwd[c], cudaEventDisableTiming));
    }
    ok(cudaEventCreateWithFlags(&s->ev_done, cudaEventDisableTiming));
    if (e != cudaSuccess) {
        pil_session_destroy(s);
        return (int)e;
    }
    *out = s;
    return PIL_OK;
}

// flags: PIL_SESSION_GRAD_ON_DEVICE -- run the backward too and leave the gradient in the session's device
// buffer (pil_session_grad_ptr), the way a training step consumes it; nothing but the loss report is
// copied back.  This path uses the training split of pil_kernels.cu: the pointwise forward of chunk c
// overlaps the H2D of chunk c+1, then ONE backward over the whole batch accumulates the stencil sums and
// finalises the loss.
static int session_run_impl(PilSession* s, const void* x_host, const void* t_host, void* grad_host, int64_t B, int x_kind,
                            const PilParams* p, float* loss_out_host, int flags, const PilExchange* ex = nullptr,
                            int64_t n_global_in = 0, float grad_scale = 1.0f) {
    if (!s || !x_host || !t_host || !p || !loss_out_host) return PIL_ERR_NULL;
    if (B < 1 || B > s->max_B) return PIL_ERR_SESSION;
    int st = pil_validate_params(p);
    if (st != PIL_OK) return st;
    PIL_TRY(cudaSetDevice(s->device));

    const int64_t img = s->H * s->W;
    const size_t xs = dsize(s->x_dtype), ts = dsize(s->t_dtype);
    const int chunks = (int)(B < kMaxChunks ? B : kMaxChunks);
    const int64_t n_global = B * img;
    auto first = [&](int c) { return (B * c) / chunks; };
    const bool device_grad = ((flags & PIL_SESSION_GRAD_ON_DEVICE) != 0 && grad_host == nullptr) || ex != nullptr;

    // phase 1: H2D + forward per chunk
    for (int c = 0; c < chunks; ++c) {
        const int64_t b0 = first(c), nb = first(c + 1) - b0;
        const size_t off = (size_t)b0 * img;
        PIL_TRY(cudaMemcpyAsync((char*)s->dx + off * xs, (const char*)x_host + off * xs, (size_t)nb * img * xs,
                                cudaMemcpyHostToDevice, s->s_h2d));
        PIL_TRY(cudaMemcpyAsync((char*)s->dt + off * ts, (const char*)t_host + off * ts, (size_t)nb * img * ts,
                                cudaMemcpyHostToDevice, s->s_h2d));
        PIL_TRY(cudaEventRecord(s->ev_in[c], s->s_h2d));
        PIL_TRY(cudaStreamWaitEvent(s->s_comp, s->ev_in[c], 0));
        if (device_grad) {
            st = pil_forward_pointwise((char*)s->dx + off * xs, (char*)s->dt + off * ts, nb, s->H, s->W, s->x_dtype, s->t_dtype,
                                       x_kind, p, s->dsums + (size_t)c * PIL_NSUMS, s->ws[c], s->ws_bytes, s->s_comp);
        } else {
            st = pil_forward((char*)s->dx + off * xs, (char*)s->dt + off * ts, nb, s->H, s->W, s->x_dtype, s->t_dtype, x_kind,
                             p, s->dsums + (size_t)c * PIL_NSUMS, nullptr, s->ws[c], s->ws_bytes, s->s_comp);
        }
        if (st != PIL_OK) return st;
    }
    double* gs = s->dsums + (size_t)kMaxChunks * PIL_NSUMS;
    pil_add_chunk_sums<<<1, 32, 0, s->s_comp>>>(s->dsums, chunks, gs);
    PIL_TRY(cudaGetLastError());
    if (ex != nullptr) {
        // data parallel: hand the shard's pointwise sums to every rank, then one backward fed from the mailbox; its last
        // block swaps the stencil sums and assembles the GLOBAL loss report
        st = pil_exchange_push(ex, 0, gs, s->s_comp);
        if (st != PIL_OK) return st;
        st = pil_backward_accumulate_xchg(s->dx, s->dt, s->dg, B, s->H, s->W, s->x_dtype, s->t_dtype, x_kind, p, ex, n_global_in,
                                          nullptr, grad_scale, s->dsums, s->dout, nullptr, s->ws[0], s->ws_bytes, s->s_comp);
        if (st != PIL_OK) return st;
        PIL_TRY(cudaMemcpyAsync(s->hout, s->dout, sizeof(float) * PIL_NOUT, cudaMemcpyDeviceToHost, s->s_comp));
        PIL_TRY(cudaStreamSynchronize(s->s_comp));
        for (int k = 0; k < PIL_NOUT; ++k) loss_out_host[k] = s->hout[k];
        return PIL_OK;
    }
    if (device_grad) {
        // one backward over the whole batch: gradient stays on the device, stencil sums + loss in the same kernel
        st = pil_backward_accumulate(s->dx, s->dt, s->dg, B, s->H, s->W, s->x_dtype, s->t_dtype, x_kind, p, gs, n_global,
                                     nullptr, 1.0f, s->dsums /* scratch: chunk rows are consumed */, s->dout, s->ws[0],
                                     s->ws_bytes, s->s_comp);
        if (st != PIL_OK) return st;
        PIL_TRY(cudaMemcpyAsync(s->hout, s->dout, sizeof(float) * PIL_NOUT, cudaMemcpyDeviceToHost, s->s_comp));
        PIL_TRY(cudaStreamSynchronize(s->s_comp));
        for (int k = 0; k < PIL_NOUT; ++k) loss_out_host[k] = s->hout[k];
        return PIL_OK;
    }
    st = pil_finalize(gs, n_global, p, s->dout, s->s_comp);
    if (st != PIL_OK) return st;
    PIL_TRY(cudaMemcpyAsync(s->hout, s->dout, sizeof(float) * PIL_NOUT, cudaMemcpyDeviceToHost, s->s_comp));

    // phase 2: backward + D2H per chunk
    if (grad_host) {
        for (int c = 0; c < chunks; ++c) {
            const int64_t b0 = first(c), nb = first(c + 1) - b0;
            const size_t off = (size_t)b0 * img;
            st = pil_backward((char*)s->dx + off * xs, (char*)s->dt + off * ts, (char*)s->dg + off * xs, nb, s->H, s->W,
                              s->x_dtype, s->t_dtype, x_kind, p, gs, n_global, nullptr, 1.0f, s->s_comp);
            if (st != PIL_OK) return st;
            PIL_TRY(cudaEventRecord(s->ev_bwd[c], s->s_comp));
            PIL_TRY(cudaStreamWaitEvent(s->s_d2h, s->ev_bwd[c], 0));
            PIL_TRY(cudaMemcpyAsync((char*)grad_host + off * xs, (char*)s->dg + off * xs, (size_t)nb * img * xs,
                                    cudaMemcpyDeviceToHost, s->s_d2h));
        }
        PIL_TRY(cudaEventRecord(s->ev_done, s->s_d2h));
        PIL_TRY(cudaStreamWaitEvent(s->s_comp, s->ev_done, 0));
    }
    PIL_TRY(cudaStreamSynchronize(s->s_comp));
    for (int k = 0; k < PIL_NOUT; ++k) loss_out_host[k] = s->hout[k];
    return PIL_OK;
}

int pil_session_run(PilSession* s, const void* x_host, const void* t_host, void* grad_host, int64_t B, int x_kind,
                    const PilParams* p, float* loss_out_host) {
    return session_run_impl(s, x_host, t_host, grad_host, B, x_kind, p, loss_out_host, 0);
}

int pil_session_run_ex(PilSession* s, const void* x_host, const void* t_host, void* grad_host, int64_t B, int x_kind,
                       const PilParams* p, float* loss_out_host, int flags) {
    return session_run_impl(s, x_host, t_host, grad_host, B, x_kind, p, loss_out_host, flags);
}

int pil_session_run_xchg(PilSession* s, const void* x_host, const void* t_host, int64_t B, int x_kind, const PilParams* p,
                         const PilExchange* ex, int64_t n_global, float grad_scale, float* loss_out_host) {
    if (!ex) return PIL_ERR_NULL;
    return session_run_impl(s, x_host, t_host, nullptr, B, x_kind, p, loss_out_host, PIL_SESSION_GRAD_ON_DEVICE, ex, n_global,
                            grad_scale);
}

void* pil_session_grad_ptr(PilSession* s) { return s ? s->dg : nullptr; }

}  // extern "C"
