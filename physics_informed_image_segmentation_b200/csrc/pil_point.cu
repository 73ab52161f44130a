// pil_point.cu -- K1L / K1M, the pointwise forward kernels of the training path, and their launchers.
// Compiled once per input kind (-DPIL_KIND=0|1|2); without PIL_KIND all three kinds are instantiated here.
#include <type_traits>

#include "pil_fwdrow.cuh"

namespace pil {
// ------------------------------------------------------------------------------------------------
// K1L: pointwise forward ("light"): only the sums that do NOT need neighbours -- I, P, T, the BCE
// terms and the double well -- as a flat, fully coalesced stream over the shard (no halo lanes, no
// row structure, 8 B/px read once).  Used by the training path, where the backward kernel visits
// every stencil anyway and accumulates sum r^2 and sum |grad u|^2 there (pil_backward_accumulate),
// so the 5-point stencils are evaluated once per step instead of twice.
// ------------------------------------------------------------------------------------------------


template <int KIND, typename XT, typename TT, bool ALIGNED>
__global__ void __launch_bounds__(kPointThreads, 4) pil_point_kernel(const PointArgs A) {
    // PDL: x and t may have been written by the kernel just before this one -> wait first; then let the
    // backward kernel's blocks take the SM slots this grid frees as its blocks retire.
    TL_STAMP(0, 0);
    pdl_wait();
    pdl_launch_dependents();
    TL_STAMP(0, 1);
    FwdRow<KIND, true> fr;
#pragma unroll
    for (int k = 0; k < 8; ++k) fr.acc[k] = 0.f;
    fr.init_packed();
    const long long tid = (long long)blockIdx.x * kPointThreads + threadIdx.x;
    const long long stride = (long long)gridDim.x * kPointThreads;
    const XT* x = reinterpret_cast<const XT*>(A.x);
    const TT* t = reinterpret_cast<const TT*>(A.t);
    if constexpr (ALIGNED) {
        const long long n4 = A.n >> 2;  // n % 4 == 0 on this path
        long long i = tid;
        // L2 policy (fp32 maps).  The backward kernel walks the shard from its END, so the last few MB of x and t
        // this front-to-back stream reads are what it needs first: those loads carry an evict_last hint.  All
        // the others are evict_first: a 537 MB stream has no business displacing the gradient lines the
        // previous backward left dirty in L2 (their write-back then interleaves with these reads).
        unsigned long long pol_keep = 0, pol_stream = 0;
        constexpr bool kHint = std::is_same<XT, float>::value && std::is_same<TT, float>::value;
        if constexpr (kHint) {
            asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
            asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
        }
        auto ldh = [&](const float* p, unsigned long long pol) -> float4 {
            float4 r;
            asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                         : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
            return r;
        };
        for (; i + (kPointUnroll - 1) * stride < n4; i += kPointUnroll * stride) {
            float4 xv[kPointUnroll], tv[kPointUnroll];
#pragma unroll
            for (int q = 0; q < kPointUnroll; ++q) {
                if constexpr (kHint) {
                    if (A.keep_from4 < n4) {
                        const unsigned long long pol = (i + q * stride >= A.keep_from4) ? pol_keep : pol_stream;
                        xv[q] = ldh(reinterpret_cast<const float*>(x) + 4 * (i + q * stride), pol);
                        tv[q] = ldh(reinterpret_cast<const float*>(t) + 4 * (i + q * stride), pol);
                        continue;
                    }
                }
                xv[q] = ld4<XT>(x + 4 * (i + q * stride));
                tv[q] = ld4<TT>(t + 4 * (i + q * stride));
            }
#pragma unroll
            for (int q = 0; q < kPointUnroll; ++q) fr.point4(xv[q], tv[q], true);
        }
        {   // last, partial batch: still issue every load before the first use
            float4 xv[kPointUnroll], tv[kPointUnroll];
#pragma unroll
            for (int q = 0; q < kPointUnroll; ++q) {
                if (i + q * stride < n4) {
                    xv[q] = ld4<XT>(x + 4 * (i + q * stride));
                    tv[q] = ld4<TT>(t + 4 * (i + q * stride));
                }
            }
#pragma unroll
            for (int q = 0; q < kPointUnroll; ++q)
                if (i + q * stride < n4) fr.point4(xv[q], tv[q], true);
        }
        fr.fold_packed();
    } else {
        // scalar path (odd sizes / unaligned bases): pairs (2i, 2i+1); an odd last pixel is paired with
        // itself and the pair's contribution halved exactly (every accumulated term is per-pixel additive)
        const long long n2 = A.n >> 1;
        for (long long i = tid; i < n2; i += stride) {
            fr.point2(make_float2(ld1<XT>(x + 2 * i), ld1<XT>(x + 2 * i + 1)),
                      make_float2(ld1<TT>(t + 2 * i), ld1<TT>(t + 2 * i + 1)), true);
        }
        fr.fold_packed();
        if ((A.n & 1) && tid == 0) {
            FwdRow<KIND, true> one;
#pragma unroll
            for (int k = 0; k < 8; ++k) one.acc[k] = 0.f;
            one.init_packed();
            const float xl = ld1<XT>(x + A.n - 1), tl = ld1<TT>(t + A.n - 1);
            one.point2(make_float2(xl, xl), make_float2(tl, tl), true);
#pragma unroll
            for (int k = 0; k < 7; ++k) fr.acc[k] += one.pa[k].x;
            fr.acc[7] += 0.5f * one.acc[7];
        }
    }
    TL_STAMP(0, 2);
    double raw[PIL_NSUMS];
    if (!reduce_to_last_block<kPointThreads, float>(fr.acc, A.partials, A.ticket, raw)) {
        TL_STAMP(0, 3);
        return;
    }
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
    if (A.X.world > 0) {  // data parallel: hand the shard's sums to every rank over NVLink
        __syncthreads();
        xchg_push(A.X, 0, s_push);
    }
    TL_STAMP(0, 3);
}

// ------------------------------------------------------------------------------------------------
// K1M: pointwise forward + per-image thresholded counts (pil_forward_pointwise_metrics).
// The per-step accuracy metrics of the reference (per-image Dice and IoU of the map thresholded at 0.5,
// src/train.py:153-160 -> src/metrics.py:38-73, src/evaluate.py:62-97) need three counts per image:
//   sum [u > thr]*t,  sum [u > thr],  sum t.
// They ride on the pass that already reads x and t.  Unlike K1L every block walks ONE contiguous piece of
// the shard, image by image, so a block flushes its counts once per image it touches (<= 3 atomics per
// warp) instead of once per iteration.
// ------------------------------------------------------------------------------------------------

template <int KIND, typename XT, typename TT, bool ALIGNED>
__global__ void __launch_bounds__(kPointThreads, 3) pil_point_metrics_kernel(const PointMetricsArgs A) {
    pdl_wait();
    pdl_launch_dependents();
    FwdRow<KIND, true> fr;
#pragma unroll
    for (int k = 0; k < 8; ++k) fr.acc[k] = 0.f;
    fr.init_packed();
    const XT* x = reinterpret_cast<const XT*>(A.x);
    const TT* t = reinterpret_cast<const TT*>(A.t);
    const float thr = A.threshold;
    constexpr int V = ALIGNED ? 4 : 1;                      // pixels per unit of work
    const long long units = A.n / V, hwu = A.hw / V;        // ALIGNED: n % 4 == 0 and hw % 4 == 0
    const long long c0 = (units * (long long)blockIdx.x) / gridDim.x, c1 = (units * ((long long)blockIdx.x + 1)) / gridDim.x;
    const int lane = threadIdx.x & 31;
    for (long long img = c0 / hwu; img * hwu < c1; ++img) {
        const long long lo = max(c0, img * hwu), hi = min(c1, (img + 1) * hwu);
        f2 cI = make_float2(0.f, 0.f), cP = make_float2(0.f, 0.f);
        const f2 T0 = fr.pa[2];
        auto count2 = [&](f2 u, f2 tt) {
            const f2 pb = make_float2(u.x > thr ? 1.0f : 0.0f, u.y > thr ? 1.0f : 0.0f);  // src/metrics.py:58
            cP = add2(cP, pb);
            cI = fma2(pb, tt, cI);
        };
        if constexpr (ALIGNED) {
            long long i = lo + threadIdx.x;
            constexpr long long S = kPointThreads;
            constexpr bool kHint = std::is_same<XT, float>::value && std::is_same<TT, float>::value;
            unsigned long long pol_stream = 0;
            if constexpr (kHint) asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
            auto ldh = [&](const float* p) -> float4 {
                float4 r;
                asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                             : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol_stream));
                return r;
            };
            for (; i < hi; i += kPointUnroll * S) {
                float4 xv[kPointUnroll], tv[kPointUnroll];
#pragma unroll
                for (int q = 0; q < kPointUnroll; ++q) {
                    if (i + q * S < hi) {
                        if constexpr (kHint) {
                            if (A.l2_stream) {
                                xv[q] = ldh(reinterpret_cast<const float*>(x) + 4 * (i + q * S));
                                tv[q] = ldh(reinterpret_cast<const float*>(t) + 4 * (i + q * S));
                                continue;
                            }
                        }
                        xv[q] = ld4<XT>(x + 4 * (i + q * S));
                        tv[q] = ld4<TT>(t + 4 * (i + q * S));
                    }
                }
#pragma unroll
                for (int q = 0; q < kPointUnroll; ++q) {
                    if (i + q * S < hi) {
                        const float4 u4 = fr.point4(xv[q], tv[q], true);
                        count2(make_float2(u4.x, u4.y), make_float2(tv[q].x, tv[q].y));
                        count2(make_float2(u4.z, u4.w), make_float2(tv[q].z, tv[q].w));
                    }
                }
            }
        } else {
            // scalar path: every pixel is evaluated as the pair (px, px); the loss sums are halved at the end
            for (long long i = lo + threadIdx.x; i < hi; i += kPointThreads) {
                const float xs = ld1<XT>(x + i), ts = ld1<TT>(t + i);
                const f2 u = fr.point2(make_float2(xs, xs), make_float2(ts, ts), true);
                const float pb = u.x > thr ? 1.0f : 0.0f;
                cP.x += pb;
                cI.x = fmaf(pb, ts, cI.x);
            }
        }
        float vI = cI.x + cI.y, vP = cP.x + cP.y;
        float vT = (fr.pa[2].x - T0.x) + (fr.pa[2].y - T0.y);
        if constexpr (!ALIGNED) vT *= 0.5f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vI += __shfl_xor_sync(0xffffffffu, vI, o);
            vP += __shfl_xor_sync(0xffffffffu, vP, o);
            vT += __shfl_xor_sync(0xffffffffu, vT, o);
        }
        if (lane == 0) {
            atomicAdd(A.image_counts + img * 4 + 0, (double)vI);
            atomicAdd(A.image_counts + img * 4 + 1, (double)vP);
            atomicAdd(A.image_counts + img * 4 + 2, (double)vT);
        }
    }
    fr.fold_packed();
    if constexpr (!ALIGNED) {
#pragma unroll
        for (int k = 0; k < 8; ++k) fr.acc[k] *= 0.5f;
    }
    double raw[PIL_NSUMS];
    if (!reduce_to_last_block<kPointThreads, float>(fr.acc, A.partials, A.ticket, raw)) return;
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
template <int KIND, typename XT, typename TT>
static cudaError_t launch_point_a(const PointArgs& a, bool aligned, cudaStream_t s, int* blocks_out) {
    static std::atomic<int> per_sm_cache[2][kMaxDevices];
    auto go = [&](auto kernel) -> cudaError_t {
        const int per_sm = blocks_per_sm_cached(kernel, kPointThreads, 0, per_sm_cache[aligned ? 1 : 0], false);
        const long long work = aligned ? (a.n >> 2) : ((a.n + 1) >> 1);                  // thread-iterations
        long long blocks = (work + (long long)kPointThreads * kPointUnroll - 1) / ((long long)kPointThreads * kPointUnroll);
        const long long cap = (long long)sm_count() * per_sm;
        if (blocks > cap) blocks = cap;
        if (blocks > kMaxPointBlocks) blocks = kMaxPointBlocks;
        if (blocks < 1) blocks = 1;
        *blocks_out = (int)blocks;
        return launch_pdl(kernel, (int)blocks, kPointThreads, 0, s, a);
    };
    if (aligned) return go(pil_point_kernel<KIND, XT, TT, true>);
    return go(pil_point_kernel<KIND, XT, TT, false>);
}
template <int KIND, typename XT>
static cudaError_t launch_point_t(int t_dtype, const PointArgs& a, bool aligned, cudaStream_t s, int* b) {
#ifdef PIL_DEV_F32_ONLY  // development builds: fp32 maps only (6x faster to compile)
    if (t_dtype != PIL_F32) return cudaErrorNotSupported;
    return launch_point_a<KIND, XT, float>(a, aligned, s, b);
#else
    switch (t_dtype) {
        case PIL_F32: return launch_point_a<KIND, XT, float>(a, aligned, s, b);
        case PIL_BF16: return launch_point_a<KIND, XT, __nv_bfloat16>(a, aligned, s, b);
        default: return launch_point_a<KIND, XT, uint8_t>(a, aligned, s, b);
    }
#endif
}
template <int KIND>
static cudaError_t launch_point_x(int x_dtype, int t_dtype, const PointArgs& a, bool aligned, cudaStream_t s, int* b) {
    if (x_dtype == PIL_F32) return launch_point_t<KIND, float>(t_dtype, a, aligned, s, b);
#ifdef PIL_DEV_F32_ONLY
    return cudaErrorNotSupported;
#else
    return launch_point_t<KIND, __nv_bfloat16>(t_dtype, a, aligned, s, b);
#endif
}
template <int KIND, typename XT, typename TT>
static cudaError_t launch_point_metrics_a(const PointMetricsArgs& a, bool aligned, cudaStream_t s, int* blocks_out) {
    const long long units = aligned ? (a.n >> 2) : a.n;
    long long blocks = (units + (long long)kPointThreads * kPointUnroll - 1) / ((long long)kPointThreads * kPointUnroll);
    const long long cap = (long long)sm_count() * 3;
    if (blocks > cap) blocks = cap;
    if (blocks > kMaxPointBlocks) blocks = kMaxPointBlocks;
    if (blocks < 1) blocks = 1;
    *blocks_out = (int)blocks;
    if (aligned) return launch_pdl(pil_point_metrics_kernel<KIND, XT, TT, true>, (int)blocks, kPointThreads, 0, s, a);
    return launch_pdl(pil_point_metrics_kernel<KIND, XT, TT, false>, (int)blocks, kPointThreads, 0, s, a);
}
template <int KIND, typename XT>
static cudaError_t launch_point_metrics_t(int t_dtype, const PointMetricsArgs& a, bool aligned, cudaStream_t s, int* b) {
#ifdef PIL_DEV_F32_ONLY
    if (t_dtype != PIL_F32) return cudaErrorNotSupported;
    return launch_point_metrics_a<KIND, XT, float>(a, aligned, s, b);
#else
    switch (t_dtype) {
        case PIL_F32: return launch_point_metrics_a<KIND, XT, float>(a, aligned, s, b);
        case PIL_BF16: return launch_point_metrics_a<KIND, XT, __nv_bfloat16>(a, aligned, s, b);
        default: return launch_point_metrics_a<KIND, XT, uint8_t>(a, aligned, s, b);
    }
#endif
}
template <int KIND>
static cudaError_t launch_point_metrics_x(int x_dtype, int t_dtype, const PointMetricsArgs& a, bool aligned, cudaStream_t s, int* b) {
    if (x_dtype == PIL_F32) return launch_point_metrics_t<KIND, float>(t_dtype, a, aligned, s, b);
#ifdef PIL_DEV_F32_ONLY
    return cudaErrorNotSupported;
#else
    return launch_point_metrics_t<KIND, __nv_bfloat16>(t_dtype, a, aligned, s, b);
#endif
}

// exported to pil_api.cu: one entry per input kind
#define PIL_POINT_EXPORT(K, KIND)                                                                                              \
    cudaError_t launch_point_k##K(int x_dtype, int t_dtype, const PointArgs& a, bool aligned, cudaStream_t s, int* b) {         \
        return launch_point_x<KIND>(x_dtype, t_dtype, a, aligned, s, b);                                                        \
    }                                                                                                                          \
    cudaError_t launch_point_metrics_k##K(int x_dtype, int t_dtype, const PointMetricsArgs& a, bool aligned, cudaStream_t s,   \
                                          int* b) {                                                                            \
        return launch_point_metrics_x<KIND>(x_dtype, t_dtype, a, aligned, s, b);                                                \
    }
#if !defined(PIL_KIND) || PIL_KIND == 0
PIL_POINT_EXPORT(0, PIL_X_PROB)
#endif
#if !defined(PIL_KIND) || PIL_KIND == 1
PIL_POINT_EXPORT(1, PIL_X_LOGITS_SIGMOID)
#endif
#if !defined(PIL_KIND) || PIL_KIND == 2
PIL_POINT_EXPORT(2, PIL_X_LOGITS_TANH)
#endif

}  // namespace pil
