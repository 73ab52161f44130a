// pil_tail.cu -- the model tail fused with the loss (pil_tail_* of include/pil.h).
//
// Reference: the U-Net ends in a 1x1 convolution from its 64-channel full-resolution feature map to one channel
// (src/unet.py:157 `self.out_conv = nn.Conv2d(base_channels, out_channels, kernel_size=1)`, applied at :205) followed by
// the output activation (:208-214); the loss then reads the probabilities (src/train.py:114-117).  In eager PyTorch
// that is: conv forward (read 64 ch, write z), sigmoid (read z, write u), the loss passes, sigmoid backward, and a conv
// backward that reads the gradient map twice and the features once more (grad_input, grad_weight, grad_bias kernels).
//
// Here the tail is two streaming kernels around the fused backward kernel of pil_bwd.cu:
//   T1  pil_tail_forward    z = sum_c w_c feat_c + b for 4 pixels per thread, written once as fp32 logits, AND the
//                           pointwise loss sums (I, P, T, BCE, double well) of pil_forward_pointwise on the same
//                           registers -- the K1L pass over the logits disappears.
//   K2  pil_backward_accumulate (unchanged) reads z, t and writes g = dL/dz plus the stencil sums.
//   T2  pil_tail_backward   ONE pass over the features: dL/dfeat_c = w_c g (written), dL/dw_c = sum g feat_c and
//                           dL/db = sum g (per-thread fp32 partials -> per-block doubles -> last block, fixed order).
// Both are bound by the 64-channel feature traffic (256 B/px read in T1; 256 B/px read + 256 B/px written in T2, fp32):
// what the fusion saves is the logits/probability round trips (~20 of ~810 B/px) and four launches, not the feature
// passes -- stated in DESIGN.md, measured in bench.py's tail leg.
//
// Layout: feat is a contiguous NCHW tensor (B, C, H, W), fp32 or bf16 (the reference trains in fp32; bf16 is what
// autocast hands over); weight is the C floats of the (1, C, 1, 1) kernel; H*W must be a multiple of 4 and the bases
// 16-byte aligned for the vector path, otherwise a scalar path runs.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pil_fwdrow.cuh"

namespace pil {

constexpr int kTailThreads = 256;
constexpr int kTailMaxC = 128;

struct TailFwdArgs {
    const void* feat;
    const void* t;
    float* logits;
    long long hw;      // H * W
    long long n;       // B * H * W
    int C;
    const float* w;     // device, C floats: the (1, C, 1, 1) kernel of the output convolution
    const float* bias;  // device, 1 float, or null
    double* partials;
    unsigned int* ticket;
    double* sums;
    PilParams p;
    XchgDev X;
};

template <typename FT>
__device__ __forceinline__ float4 ldf4(const FT* p);
template <>
__device__ __forceinline__ float4 ldf4<float>(const float* p) {
    return __ldcs(reinterpret_cast<const float4*>(p));  // read once
}
template <>
__device__ __forceinline__ float4 ldf4<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 raw = __ldcs(reinterpret_cast<const uint2*>(p));
    float4 r;
    r.x = __uint_as_float(raw.x << 16);
    r.y = __uint_as_float(raw.x & 0xffff0000u);
    r.z = __uint_as_float(raw.y << 16);
    r.w = __uint_as_float(raw.y & 0xffff0000u);
    return r;
}
template <typename FT>
__device__ __forceinline__ float ldf1(const FT* p) {
    if constexpr (std::is_same<FT, float>::value) return *p;
    else return __bfloat162float(*p);
}

// T1: 1x1 convolution + pointwise loss sums.  A "unit" is 4 adjacent pixels of one image (VEC) or one pixel (!VEC).
template <int KIND, typename FT, typename TT, bool VEC>
__global__ void __launch_bounds__(kTailThreads) pil_tail_fwd_kernel(const TailFwdArgs A) {
    pdl_wait();
    pdl_launch_dependents();
    __shared__ float s_wt[kTailMaxC];
    for (int c = threadIdx.x; c < A.C; c += kTailThreads) s_wt[c] = A.w[c];
    const float bias = A.bias != nullptr ? __ldg(A.bias) : 0.f;
    __syncthreads();
    FwdRow<KIND, true> fr;
#pragma unroll
    for (int k = 0; k < 8; ++k) fr.acc[k] = 0.f;
    fr.init_packed();
    const FT* feat = reinterpret_cast<const FT*>(A.feat);
    const TT* t = reinterpret_cast<const TT*>(A.t);
    constexpr int V = VEC ? 4 : 1;
    const long long units = A.n / V, hwu = A.hw / V;
    const long long chw = (long long)A.C * A.hw;
    for (long long q = (long long)blockIdx.x * kTailThreads + threadIdx.x; q < units; q += (long long)gridDim.x * kTailThreads) {
        const long long b = q / hwu, r = (q - b * hwu) * V;        // image, first pixel inside the image
        const FT* f = feat + b * chw + r;
        if constexpr (VEC) {
            float4 z = make_float4(bias, bias, bias, bias);
#pragma unroll 8
            for (int c = 0; c < A.C; ++c) {
                const float4 v = ldf4<FT>(f + (long long)c * A.hw);
                const float wc = s_wt[c];
                z.x = fmaf(wc, v.x, z.x);
                z.y = fmaf(wc, v.y, z.y);
                z.z = fmaf(wc, v.z, z.z);
                z.w = fmaf(wc, v.w, z.w);
            }
            const long long px = b * A.hw + r;
            *reinterpret_cast<float4*>(A.logits + px) = z;          // re-read by the backward kernel: default policy
            fr.point4(z, ld4<TT>(t + px), true);
        } else {
            float z = bias;
            for (int c = 0; c < A.C; ++c) z = fmaf(s_wt[c], ldf1<FT>(f + (long long)c * A.hw), z);
            const long long px = b * A.hw + r;
            A.logits[px] = z;
            const float tv = ld1<TT>(t + px);
            fr.point2(make_float2(z, z), make_float2(tv, tv), true);  // the pixel twice; halved below
        }
    }
    fr.fold_packed();
    if constexpr (!VEC) {
#pragma unroll
        for (int k = 0; k < 8; ++k) fr.acc[k] *= 0.5f;
    }
    double raw[PIL_NSUMS];
    if (!reduce_to_last_block<kTailThreads, float>(fr.acc, A.partials, A.ticket, raw)) return;
    __shared__ double s_push[PIL_NSUMS];
    if (threadIdx.x == 0) {
        double sv[PIL_NSUMS];
        sums_from_raw(raw, A.p.epsilon, (double)A.n, sv);
#pragma unroll
        for (int k = 0; k < PIL_NSUMS; ++k) {
            A.sums[k] = sv[k];
            s_push[k] = sv[k];
        }
    }
    if (A.X.world > 0) {
        __syncthreads();
        xchg_push(A.X, 0, s_push);
    }
}

struct TailBwdArgs {
    const void* feat;
    const float* g;        // dL/dlogits (B, 1, H, W)
    void* dfeat;
    long long hw, n;
    int C;
    const float* w;        // device, C floats
    double* partials;      // [blocks][C + 1]
    unsigned int* ticket;
    float* dweight;        // [C]
    float* dbias;          // [1] or null
};

template <typename FT>
__device__ __forceinline__ void stf4(FT* p, float4 v) {
    if constexpr (std::is_same<FT, float>::value) {
        __stcs(reinterpret_cast<float4*>(p), v);
    } else {
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 raw;
        raw.x = *reinterpret_cast<uint32_t*>(&lo);
        raw.y = *reinterpret_cast<uint32_t*>(&hi);
        __stcs(reinterpret_cast<uint2*>(p), raw);
    }
}

// T2: dL/dfeat, dL/dweight and dL/dbias in one pass over the features.  CB channels are processed per sweep so that the
// per-thread partials of dL/dw stay in registers (CB accumulators); C = 64 takes two sweeps of 32 over the same pixels,
// the second of which finds g in L1/L2.
template <typename FT, bool VEC, int CB>
__global__ void __launch_bounds__(kTailThreads) pil_tail_bwd_kernel(const TailBwdArgs A) {
    const FT* feat = reinterpret_cast<const FT*>(A.feat);
    FT* dfeat = reinterpret_cast<FT*>(A.dfeat);
    constexpr int V = VEC ? 4 : 1;
    const long long units = A.n / V, hwu = A.hw / V;
    const long long chw = (long long)A.C * A.hw;
    __shared__ float s_w[kTailThreads / 32][CB + 1];
    __shared__ float s_wt[kTailMaxC];
    for (int c = threadIdx.x; c < A.C; c += kTailThreads) s_wt[c] = A.w[c];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c0 = 0; c0 < A.C; c0 += CB) {
        float dw[CB];
#pragma unroll
        for (int j = 0; j < CB; ++j) dw[j] = 0.f;
        float db = 0.f;
        for (long long q = (long long)blockIdx.x * kTailThreads + threadIdx.x; q < units; q += (long long)gridDim.x * kTailThreads) {
            const long long b = q / hwu, r = (q - b * hwu) * V;
            const long long px = b * A.hw + r, fo = b * chw + r;
            if constexpr (VEC) {
                const float4 g = *reinterpret_cast<const float4*>(A.g + px);
                db += (g.x + g.y) + (g.z + g.w);
#pragma unroll
                for (int j = 0; j < CB; ++j) {
                    if (c0 + j < A.C) {
                        const long long o = fo + (long long)(c0 + j) * A.hw;
                        const float4 v = ldf4<FT>(feat + o);
                        const float wc = s_wt[c0 + j];
                        stf4<FT>(dfeat + o, make_float4(wc * g.x, wc * g.y, wc * g.z, wc * g.w));
                        dw[j] = fmaf(g.x, v.x, fmaf(g.y, v.y, fmaf(g.z, v.z, fmaf(g.w, v.w, dw[j]))));
                    }
                }
            } else {
                const float g = A.g[px];
                db += g;
#pragma unroll
                for (int j = 0; j < CB; ++j) {
                    if (c0 + j < A.C) {
                        const long long o = fo + (long long)(c0 + j) * A.hw;
                        const float v = ldf1<FT>(feat + o);
                        if constexpr (std::is_same<FT, float>::value) dfeat[o] = s_wt[c0 + j] * g;
                        else dfeat[o] = __float2bfloat16_rn(s_wt[c0 + j] * g);
                        dw[j] = fmaf(g, v, dw[j]);
                    }
                }
            }
        }
        // block partials of this channel sweep (and of dL/db in the first sweep): warp shuffles -> shared -> doubles
#pragma unroll
        for (int j = 0; j < CB; ++j) {
            float v = dw[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s_w[warp][j] = v;
        }
        {
            float v = db;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s_w[warp][CB] = v;
        }
        __syncthreads();
        if (threadIdx.x <= CB) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < kTailThreads / 32; ++w) v += (double)s_w[w][threadIdx.x];
            const int slot = (threadIdx.x == CB) ? A.C : c0 + (int)threadIdx.x;
            if (threadIdx.x < CB ? (c0 + (int)threadIdx.x < A.C) : (c0 == 0))
                A.partials[(long long)blockIdx.x * (A.C + 1) + slot] = v;
        }
        __syncthreads();
    }
    // last block: every channel's partials in block order (bit-reproducible)
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(A.ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int c = threadIdx.x; c <= A.C; c += kTailThreads) {
        double v = 0.0;
        for (unsigned int b = 0; b < gridDim.x; ++b) v += __ldcg(A.partials + (long long)b * (A.C + 1) + c);
        if (c < A.C) A.dweight[c] = (float)v;
        else if (A.dbias != nullptr) A.dbias[0] = (float)v;
    }
    if (threadIdx.x == 0) *A.ticket = 0u;
}

static int tail_blocks(long long units) {
    long long b = (units + kTailThreads - 1) / kTailThreads;
    const long long cap = (long long)sm_count() * 8;
    if (b > cap) b = cap;
    if (b > kMaxPointBlocks) b = kMaxPointBlocks;
    return (int)(b < 1 ? 1 : b);
}

template <int KIND, typename FT>
static cudaError_t launch_tail_fwd(int t_dtype, const TailFwdArgs& a, bool vec, int blocks, cudaStream_t s) {
#define PIL_TAIL_GO(TT)                                                                                          \
    return vec ? launch_pdl(pil_tail_fwd_kernel<KIND, FT, TT, true>, blocks, kTailThreads, 0, s, a)              \
               : launch_pdl(pil_tail_fwd_kernel<KIND, FT, TT, false>, blocks, kTailThreads, 0, s, a)
    switch (t_dtype) {
        case PIL_F32: PIL_TAIL_GO(float);
        case PIL_BF16: PIL_TAIL_GO(__nv_bfloat16);
        default: PIL_TAIL_GO(uint8_t);
    }
#undef PIL_TAIL_GO
}

}  // namespace pil

using namespace pil;

extern "C" {

size_t pil_tail_workspace_bytes(int64_t C) {
    if (C < 1 || C > kTailMaxC) return 0;
    // ticket + partials of the forward sums (8 doubles per block) and of the backward (C + 1 doubles per block)
    // (the forward's partials are tagged 16-byte slots, pil_common.cuh)
    return 256 + (size_t)kMaxPointBlocks * (size_t)(C + 1 > 2 * PIL_NSUMS ? C + 1 : 2 * PIL_NSUMS) * sizeof(double);
}

int pil_tail_forward(const void* feat, int feat_dtype, const float* weight, const float* bias, const void* t, int t_dtype, int64_t B,
                     int64_t C, int64_t H, int64_t W, int x_kind, const PilParams* p, float* logits_out, double* sums,
                     void* workspace, size_t workspace_bytes, const PilExchange* ex, void* stream) {
    if (!feat || !weight || !t || !p || !logits_out || !sums || !workspace) return PIL_ERR_NULL;
    if (B < 1 || H < 2 || W < 2 || C < 1 || C > kTailMaxC || B * C * H * W > ((int64_t)1 << 42)) return PIL_ERR_SHAPE;
    if (!(feat_dtype == PIL_F32 || feat_dtype == PIL_BF16)) return PIL_ERR_DTYPE;
    if (!(t_dtype == PIL_F32 || t_dtype == PIL_BF16 || t_dtype == PIL_U8)) return PIL_ERR_DTYPE;
    if (x_kind != PIL_X_LOGITS_SIGMOID && x_kind != PIL_X_LOGITS_TANH) return PIL_ERR_KIND;  // the tail produces logits
    int st = pil_validate_params(p);
    if (st != PIL_OK) return st;
    if (workspace_bytes < pil_tail_workspace_bytes(C) || ((uintptr_t)workspace % 8)) return PIL_ERR_WORKSPACE;
    TailFwdArgs a;
    a.feat = feat;
    a.t = t;
    a.logits = logits_out;
    a.hw = H * W;
    a.n = B * H * W;
    a.C = (int)C;
    a.bias = bias;
    a.w = weight;
    a.ticket = reinterpret_cast<unsigned int*>(workspace);
    a.partials = reinterpret_cast<double*>((char*)workspace + 256);
    a.sums = sums;
    a.p = *p;
    st = make_xchg(ex, &a.X);
    if (st != PIL_OK) return st;
    const size_t fs = feat_dtype == PIL_F32 ? 4 : 2, ts = t_dtype == PIL_F32 ? 4 : (t_dtype == PIL_BF16 ? 2 : 1);
    const bool vec = (a.hw % 4 == 0) && ((uintptr_t)feat % (4 * fs) == 0) && ((uintptr_t)t % (4 * ts) == 0) && ((uintptr_t)logits_out % 16 == 0);
    const int blocks = tail_blocks(vec ? a.n / 4 : a.n);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e;
    if (x_kind == PIL_X_LOGITS_SIGMOID)
        e = feat_dtype == PIL_F32 ? launch_tail_fwd<PIL_X_LOGITS_SIGMOID, float>(t_dtype, a, vec, blocks, s)
                                  : launch_tail_fwd<PIL_X_LOGITS_SIGMOID, __nv_bfloat16>(t_dtype, a, vec, blocks, s);
    else
        e = feat_dtype == PIL_F32 ? launch_tail_fwd<PIL_X_LOGITS_TANH, float>(t_dtype, a, vec, blocks, s)
                                  : launch_tail_fwd<PIL_X_LOGITS_TANH, __nv_bfloat16>(t_dtype, a, vec, blocks, s);
    host_state().kernels_launched.fetch_add(1, std::memory_order_relaxed);
    return (int)e;
}

int pil_tail_backward(const void* feat, int feat_dtype, const float* weight, const float* grad_logits, int64_t B, int64_t C,
                      int64_t H, int64_t W, void* grad_feat, float* grad_weight, float* grad_bias, void* workspace,
                      size_t workspace_bytes, void* stream) {
    if (!feat || !weight || !grad_logits || !grad_feat || !grad_weight || !workspace) return PIL_ERR_NULL;
    if (B < 1 || H < 1 || W < 1 || C < 1 || C > kTailMaxC) return PIL_ERR_SHAPE;
    if (!(feat_dtype == PIL_F32 || feat_dtype == PIL_BF16)) return PIL_ERR_DTYPE;
    if (workspace_bytes < pil_tail_workspace_bytes(C) || ((uintptr_t)workspace % 8)) return PIL_ERR_WORKSPACE;
    TailBwdArgs a;
    a.feat = feat;
    a.g = grad_logits;
    a.dfeat = grad_feat;
    a.hw = H * W;
    a.n = B * H * W;
    a.C = (int)C;
    a.w = weight;
    a.ticket = reinterpret_cast<unsigned int*>(workspace);
    a.partials = reinterpret_cast<double*>((char*)workspace + 256);
    a.dweight = grad_weight;
    a.dbias = grad_bias;
    const size_t fs = feat_dtype == PIL_F32 ? 4 : 2;
    const bool vec = (a.hw % 4 == 0) && ((uintptr_t)feat % (4 * fs) == 0) && ((uintptr_t)grad_feat % (4 * fs) == 0) &&
                     ((uintptr_t)grad_logits % 16 == 0);
    const int blocks = tail_blocks(vec ? a.n / 4 : a.n);
    cudaStream_t s = (cudaStream_t)stream;
    constexpr int CB = 32;
    if (feat_dtype == PIL_F32) {
        if (vec) pil_tail_bwd_kernel<float, true, CB><<<blocks, kTailThreads, 0, s>>>(a);
        else pil_tail_bwd_kernel<float, false, CB><<<blocks, kTailThreads, 0, s>>>(a);
    } else {
        if (vec) pil_tail_bwd_kernel<__nv_bfloat16, true, CB><<<blocks, kTailThreads, 0, s>>>(a);
        else pil_tail_bwd_kernel<__nv_bfloat16, false, CB><<<blocks, kTailThreads, 0, s>>>(a);
    }
    host_state().kernels_launched.fetch_add(1, std::memory_order_relaxed);
    return (int)cudaGetLastError();
}

}  // extern "C"
