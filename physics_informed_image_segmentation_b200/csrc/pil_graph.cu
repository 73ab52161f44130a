// pil_graph.cu -- one training-step evaluation of the loss as ONE CUDA-graph launch (pil_step_graph_* of include/pil.h).
//
// At the reference's real batch shape (8 x 1 x 128 x 128, src/dataset.py:18) and for the shards a strong-scaled
// data-parallel batch leaves per GPU, the two fused kernels of a step take a few microseconds each and the cost of a
// step is the host's: two C-ABI calls, two launches with their attribute blocks, the occupancy and tiling
// arithmetic.  The pair is therefore captured once -- with its programmatic-dependent-launch edge, so the backward's
// blocks still slide in while the forward drains -- and replayed with a single cudaGraphLaunch per step.
// The captured launches take their arguments by value: the graph is bound to the buffers it was created with.  The
// data-parallel form uses the device-resident exchange epoch (PIL_XCHG_DEVICE_EPOCH), the only per-step argument
// that would otherwise change.
#include <cuda_runtime.h>
#include <stdint.h>

#include <new>

#include "pil.h"

struct PilStepGraph {
    int device;
    cudaStream_t capture_stream;
    cudaGraph_t graph;
    cudaGraphExec_t exec;
};

namespace {

struct StepArgs {
    const void* x;
    const void* t;
    void* grad;
    int64_t B, H, W;
    int x_dtype, t_dtype, x_kind;
    PilParams p;
    double* sums;
    float* loss_out;
    void* workspace;
    size_t workspace_bytes;
    const PilExchange* ex;
    int64_t n_global;
    const float* upstream;
    float grad_scale;
    double* stencil_sums;
    double* total_sums;
};

// the two launches of a step on stream s (same calls a caller would make one by one)
int enqueue_step(const StepArgs& a, cudaStream_t s) {
    int st;
    if (a.ex != nullptr) {
        st = pil_forward_pointwise_xchg(a.x, a.t, a.B, a.H, a.W, a.x_dtype, a.t_dtype, a.x_kind, &a.p, a.sums, a.workspace,
                                        a.workspace_bytes, a.ex, s);
        if (st != PIL_OK) return st;
        return pil_backward_accumulate_xchg(a.x, a.t, a.grad, a.B, a.H, a.W, a.x_dtype, a.t_dtype, a.x_kind, &a.p, a.ex, a.n_global,
                                            a.upstream, a.grad_scale, a.stencil_sums, a.loss_out, a.total_sums, a.workspace,
                                            a.workspace_bytes, s);
    }
    if (a.upstream == nullptr && a.grad_scale == 1.0f && a.total_sums == a.sums)
        return pil_loss_fwd_bwd(a.x, a.t, a.grad, a.B, a.H, a.W, a.x_dtype, a.t_dtype, a.x_kind, &a.p, a.sums, a.loss_out, a.workspace,
                                a.workspace_bytes, s);
    st = pil_forward_pointwise(a.x, a.t, a.B, a.H, a.W, a.x_dtype, a.t_dtype, a.x_kind, &a.p, a.sums, a.workspace, a.workspace_bytes, s);
    if (st != PIL_OK) return st;
    return pil_backward_accumulate(a.x, a.t, a.grad, a.B, a.H, a.W, a.x_dtype, a.t_dtype, a.x_kind, &a.p, a.sums, a.n_global, a.upstream,
                                   a.grad_scale, a.stencil_sums, a.loss_out, a.workspace, a.workspace_bytes, s);
}

// the sweep step (BASELINE config 4): one pass for the moments, one block for the n_params loss reports
struct SweepArgs {
    const void* x;
    const void* t;
    int64_t B, H, W;
    int x_dtype, t_dtype, x_kind;
    double* moments;
    void* workspace;
    size_t workspace_bytes;
    const PilExchange* ex;
    int64_t n_global;
    const PilParams* params;
    int n_params;
    float* loss_out;
    double* moments_out;
};
int enqueue_sweep(const SweepArgs& a, cudaStream_t s) {
    if (a.ex != nullptr) {
        const int st = pil_forward_moments_xchg(a.x, a.t, a.B, a.H, a.W, a.x_dtype, a.t_dtype, a.x_kind, a.moments, a.workspace,
                                                a.workspace_bytes, a.ex, s);
        if (st != PIL_OK) return st;
        return pil_sweep_finalize_xchg(a.ex, a.n_global, a.params, a.n_params, a.loss_out, a.moments_out, s);
    }
    const int st = pil_forward_moments(a.x, a.t, a.B, a.H, a.W, a.x_dtype, a.t_dtype, a.x_kind, a.moments, a.workspace, a.workspace_bytes, s);
    if (st != PIL_OK) return st;
    return pil_sweep_finalize(a.moments, a.n_global, a.params, a.n_params, a.loss_out, s);
}

// run `enqueue` once for real on the caller's stream (validates, fills the per-device caches outside of capture and -- data
// parallel -- is a step like any other), then capture it on a private stream and instantiate
template <typename F>
int capture_graph(PilStepGraph** out, F enqueue, cudaStream_t user) {
    int st = enqueue(user);
    if (st != PIL_OK) return st;
    cudaError_t e = cudaStreamSynchronize(user);
    if (e != cudaSuccess) return (int)e;
    PilStepGraph* g = new (std::nothrow) PilStepGraph();
    if (!g) return (int)cudaErrorMemoryAllocation;
    *g = PilStepGraph{};
    cudaGetDevice(&g->device);
    e = cudaStreamCreateWithFlags(&g->capture_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamBeginCapture(g->capture_stream, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) {
        pil_step_graph_destroy(g);
        return (int)e;
    }
    st = enqueue(g->capture_stream);
    e = cudaStreamEndCapture(g->capture_stream, &g->graph);  // always end the capture, also after a failed enqueue
    if (st == PIL_OK && e == cudaSuccess) e = cudaGraphInstantiate(&g->exec, g->graph, 0);
    if (st != PIL_OK || e != cudaSuccess) {
        pil_step_graph_destroy(g);
        cudaGetLastError();
        return st != PIL_OK ? st : (int)e;
    }
    *out = g;
    return PIL_OK;
}

}  // namespace

extern "C" {

int pil_step_graph_destroy(PilStepGraph* g) {
    if (!g) return PIL_ERR_NULL;
    cudaSetDevice(g->device);
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    if (g->capture_stream) cudaStreamDestroy(g->capture_stream);
    delete g;
    return PIL_OK;
}

int pil_step_graph_create(PilStepGraph** out, const void* x, const void* t, void* grad, int64_t B, int64_t H, int64_t W, int x_dtype,
                          int t_dtype, int x_kind, const PilParams* p, double* sums, float* loss_out, double* stencil_sums,
                          void* workspace, size_t workspace_bytes, const PilExchange* ex, int64_t n_global, const float* upstream,
                          float grad_scale, double* total_sums, void* stream) {
    if (!out || !x || !t || !grad || !p || !sums || !loss_out || !stencil_sums || !workspace) return PIL_ERR_NULL;
    if (ex != nullptr && !(ex->flags & PIL_XCHG_DEVICE_EPOCH)) return PIL_ERR_EXCHANGE;  // a host epoch would be frozen into the graph
    StepArgs a = {x, t, grad, B, H, W, x_dtype, t_dtype, x_kind, *p, sums, loss_out, workspace, workspace_bytes, ex, n_global,
                  upstream, grad_scale, stencil_sums, total_sums};
    // One real step first, on the caller's stream (data parallel: taken by every rank -- creation is collective), then
    // the capture: see capture_graph.
    return capture_graph(out, [&](cudaStream_t s) { return enqueue_step(a, s); }, (cudaStream_t)stream);
}

int pil_sweep_graph_create(PilStepGraph** out, const void* x, const void* t, int64_t B, int64_t H, int64_t W, int x_dtype, int t_dtype,
                           int x_kind, double* moments, void* workspace, size_t workspace_bytes, const PilExchange* ex,
                           int64_t n_global, const PilParams* params, int n_params, float* loss_out, double* moments_out,
                           void* stream) {
    if (!out || !x || !t || !moments || !workspace || !params || !loss_out) return PIL_ERR_NULL;
    if (ex != nullptr && !(ex->flags & PIL_XCHG_DEVICE_EPOCH)) return PIL_ERR_EXCHANGE;  // a host epoch would be frozen into the graph
    const SweepArgs a = {x, t, B, H, W, x_dtype, t_dtype, x_kind, moments, workspace, workspace_bytes, ex, n_global, params, n_params,
                         loss_out, moments_out};
    return capture_graph(out, [&](cudaStream_t s) { return enqueue_sweep(a, s); }, (cudaStream_t)stream);
}

int pil_step_graph_launch(PilStepGraph* g, void* stream) {
    if (!g || !g->exec) return PIL_ERR_NULL;
    return (int)cudaGraphLaunch(g->exec, (cudaStream_t)stream);
}

}  // extern "C"
