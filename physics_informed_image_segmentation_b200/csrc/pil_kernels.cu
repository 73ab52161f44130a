// pil_kernels.cu -- fused sm_100a kernels of the physics-prior loss and the C ABI of include/pil.h.
//
// Hot path being replaced (reference file:line, relative to the reference checkout):
//   src/unet.py:208-214   output activation (sigmoid | (tanh+1)/2)
//   src/loss.py:114-162   DiceBCEPDELoss.forward (batch-global Dice, nn.BCELoss, weight gates)
//   src/pde.py:49-145     reflect-pad 5-point Laplacian, cubic reaction, residual, mean(r^2)
//   src/pde.py:147-212    central-difference |grad u|^2, double well, phase-field mean
//   autograd of all of the above
//
// Design: the path is an HBM-bound stencil + reduction, so no tensor cores.
//   * Each thread owns 4 adjacent columns (one 128-bit load per row per map) and marches DOWN a
//     segment of rows, keeping the rows it still needs in a register ring: vertical neighbours cost
//     nothing, horizontal neighbours are two warp shuffles.  A warp therefore covers a 128-column
//     strip of which lanes 1..30 (120 columns) produce output and lanes 0/31 only supply the halo.
//   * The reflect boundary is resolved at LOAD time (mirrored row index, mirrored halo column), so
//     the stencil arithmetic in the loop has no boundary predicates at all.
//   * forward: per-thread fp32 partial sums -> warp shuffles -> block -> per-block double partials
//     -> the last block to finish adds them in a fixed order (deterministic) and finalises the loss.
//   * backward: gather-free "scatter in registers": when residual row k is formed it is pushed into
//     the gradient accumulators of rows k-1, k, k+1 that the thread holds, so r is computed once per
//     pixel and the transpose of the (non-symmetric) reflect-Laplacian falls out of two row/column
//     factors.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pil.h"

namespace pil {

constexpr int kVec = 4;                     // columns per thread
constexpr int kOutLanes = 30;               // lanes of a warp that own output columns
constexpr int kStripCols = kOutLanes * kVec;  // 120 output columns per warp
constexpr int kWarpsPerBlock = 4;
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr int kFwdMinBlocks = 6;  // 24 warps / SM, <= 85 registers
constexpr int kBwdMinBlocks = 5;  // 20 warps / SM, <= 102 registers
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kLogClampLog2 = -100.0f * kLog2e;  // nn.BCELoss clamps ln() at -100

static_assert(PIL_NSUMS == 8, "sums layout");

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
// The three MUFU approximations are the flush-to-zero forms: one instruction each instead of the
// 4-5 the denormal-preserving forms expand to.  Consequence (documented in DESIGN.md): logits below
// -87.3 give u == 0 exactly (the reference reaches u == 0 at -88.7) and denormal probabilities are
// treated as 0 by the BCE logarithm; nothing else changes.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// u from x: src/unet.py:208-214.  (tanh(z)+1)/2 == sigmoid(2z).
template <int KIND>
__device__ __forceinline__ float activate(float x) {
    if constexpr (KIND == PIL_X_PROB) {
        return x;
    } else {
        const float s = (KIND == PIL_X_LOGITS_TANH) ? -2.0f * kLog2e : -kLog2e;
        return rcp_approx(1.0f + ex2_approx(x * s));
    }
}

__device__ __forceinline__ int mirror_clamp(int k, int n) {
    // torch reflect pad 1: -1 -> 1, n -> n-2 (src/pde.py:67); clamp keeps never-used slots in range
    k = (k < 0) ? -k : k;
    k = (k >= n) ? 2 * n - 2 - k : k;
    return min(max(k, 0), n - 1);
}

template <typename T>
__device__ __forceinline__ float ld1(const T* p);
template <>
__device__ __forceinline__ float ld1<float>(const float* p) {
    return __ldg(p);
}
template <>
__device__ __forceinline__ float ld1<__nv_bfloat16>(const __nv_bfloat16* p) {
    return __bfloat162float(*p);
}
template <>
__device__ __forceinline__ float ld1<uint8_t>(const uint8_t* p) {
    return (float)__ldg(p);
}

template <typename T>
__device__ __forceinline__ float4 ld4(const T* p);
template <>
__device__ __forceinline__ float4 ld4<float>(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}
template <>
__device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
    float4 r;
    r.x = __uint_as_float(raw.x << 16);
    r.y = __uint_as_float(raw.x & 0xffff0000u);
    r.z = __uint_as_float(raw.y << 16);
    r.w = __uint_as_float(raw.y & 0xffff0000u);
    return r;
}
template <>
__device__ __forceinline__ float4 ld4<uint8_t>(const uint8_t* p) {
    const uint32_t raw = __ldg(reinterpret_cast<const uint32_t*>(p));
    return make_float4((float)(raw & 0xff), (float)((raw >> 8) & 0xff), (float)((raw >> 16) & 0xff),
                       (float)(raw >> 24));
}

template <typename T>
__device__ __forceinline__ void st4(T* p, float4 v);
template <>
__device__ __forceinline__ void st4<float>(float* p, float4 v) {
    __stcs(reinterpret_cast<float4*>(p), v);
}
template <>
__device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 raw;
    raw.x = *reinterpret_cast<uint32_t*>(&lo);
    raw.y = *reinterpret_cast<uint32_t*>(&hi);
    __stcs(reinterpret_cast<uint2*>(p), raw);
}
template <typename T>
__device__ __forceinline__ void st1(T* p, float v);
template <>
__device__ __forceinline__ void st1<float>(float* p, float v) {
    *p = v;
}
template <>
__device__ __forceinline__ void st1<__nv_bfloat16>(__nv_bfloat16* p, float v) {
    *p = __float2bfloat16_rn(v);
}

// How one thread reads its 4 columns of a row.  ALIGNED (W % 4 == 0, 16-byte aligned bases): every
// lane issues one 128-bit load at its column clamped into the image, so the load is branch-free;
// the two halo lanes that hang over the image edge then move the mirrored column into the slot
// their neighbour reads (col -1 := col 1, col W := col W-2; src/pde.py:67).  Lanes further out hold
// finite, never-used data.  Otherwise four scalar loads at mirrored/clamped columns.
template <bool ALIGNED>
struct Cols {
    int col0;
    int colc;     // ALIGNED: col0 clamped to [0, W-4]
    int mode;     // ALIGNED: 0 in image, 1 left-edge halo, 2 right-edge halo, 3 outside
    int idx[4];   // !ALIGNED: mirrored+clamped column of each slot
    __device__ __forceinline__ void init(int c0, int W) {
        col0 = c0;
        colc = min(max(c0, 0), W - kVec);
        if constexpr (ALIGNED) {
            mode = (c0 >= 0 && c0 < W) ? 0 : (c0 == -kVec ? 1 : (c0 == W ? 2 : 3));
        } else {
            mode = 0;
#pragma unroll
            for (int p = 0; p < 4; ++p) idx[p] = mirror_clamp(c0 + p, W);
        }
    }
    // `row` already points at this thread's (clamped) column for ALIGNED, at column 0 otherwise
    template <typename T>
    __device__ __forceinline__ float4 load(const T* row) const {
        if constexpr (ALIGNED) {
            float4 r = ld4<T>(row);
            if (mode == 1) r.w = r.y;
            if (mode == 2) r.x = r.z;
            return r;
        } else {
            return make_float4(ld1<T>(row + idx[0]), ld1<T>(row + idx[1]), ld1<T>(row + idx[2]),
                               ld1<T>(row + idx[3]));
        }
    }
    template <typename T>
    __device__ __forceinline__ float4 load_plain(const T* row) const {  // targets: no halo needed
        if constexpr (ALIGNED) {
            return ld4<T>(row);
        } else {
            return make_float4(ld1<T>(row + idx[0]), ld1<T>(row + idx[1]), ld1<T>(row + idx[2]),
                               ld1<T>(row + idx[3]));
        }
    }
};

template <bool V>
struct BoolC {
    static constexpr bool value = V;
};
template <int KIND>
__device__ __forceinline__ float4 act4(float4 v) {
    return make_float4(activate<KIND>(v.x), activate<KIND>(v.y), activate<KIND>(v.z), activate<KIND>(v.w));
}

// ------------------------------------------------------------------------------------------------
// geometry shared by host and device
// ------------------------------------------------------------------------------------------------
struct Geo {
    int B, H, W;
    int strips;         // warps per row band = ceil(W / 120)
    int segs;           // row segments per image
    int rps;            // rows per segment
    long long tasks;    // B * segs * strips warp-tasks
};

struct FwdArgs {
    const void* x;
    const void* t;
    Geo g;
    float D, a;
    double* partials;        // [blocks][PIL_NSUMS]
    unsigned int* ticket;    // zero on entry, zero on exit
    double* sums;            // [PIL_NSUMS]
    float* loss_out;         // may be null
    PilParams p;
};

struct BwdArgs {
    const void* x;
    const void* t;
    void* grad;
    Geo g;
    const double* gsums;
    const float* upstream;
    float grad_scale;
    long long n_global;
    PilParams p;
};

__device__ __forceinline__ void finalize_device(const double* s, double n, const PilParams& p, float* out) {
    // src/loss.py:134-160
    const double I = s[0], P = s[1], T = s[2];
    const double dice_loss = 1.0 - (2.0 * I + p.smooth) / (P + T + p.smooth);
    const double bce = s[3] / n, rd = s[4] / n, pf = s[5] / n;
    double total = p.dice_weight * dice_loss + p.bce_weight * bce;
    if (p.pde_weight > 0.0) total += p.pde_weight * rd;
    if (p.phase_field_weight > 0.0) total += p.phase_field_weight * pf;
    out[0] = (float)total;
    out[1] = (float)dice_loss;
    out[2] = (float)bce;
    out[3] = (float)rd;
    out[4] = (float)pf;
    out[5] = (float)s[6];
    out[6] = 0.f;
    out[7] = 0.f;
}

// ------------------------------------------------------------------------------------------------
// K1: fused forward
// ------------------------------------------------------------------------------------------------
template <int KIND, bool ALIGNED>
struct FwdRow {
    // accumulators: 0 I, 1 P, 2 T, 3 sum(t*max(lg2 u,c) + (1-t)*max(lg2(1-u),c)), 4 sum r^2,
    //               5 sum dx^2+dy^2 (raw differences), 6 sum (u(1-u))^2, 7 #invalid
    float acc[8];
    float D, a;
    float m[4];  // !ALIGNED: 1 for slots that are real output pixels of this thread

    __device__ __forceinline__ void row(const float4& um, const float4& uc, const float4& up, const float4& tt) {
        const float L = __shfl_up_sync(0xffffffffu, uc.w, 1);
        const float R = __shfl_down_sync(0xffffffffu, uc.x, 1);
        const float e[6] = {L, uc.x, uc.y, uc.z, uc.w, R};
        const float vm[4] = {um.x, um.y, um.z, um.w};
        const float vp[4] = {up.x, up.y, up.z, up.w};
        const float vt[4] = {tt.x, tt.y, tt.z, tt.w};
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const float u = e[p + 1], t = vt[p];
            const float v = 1.0f - u;
            const float uv = u * v;
            const float lap = (e[p] + e[p + 2]) + (vm[p] + vp[p]) - 4.0f * u;  // src/pde.py:73-77
            const float r = fmaf(D, lap, uv * (u - a));                         // src/pde.py:99,:120
            const float dx = e[p + 2] - e[p], dy = vp[p] - vm[p];               // 2*gx, 2*gy (src/pde.py:172-173)
            const float lu = fmaxf(lg2_approx(u), kLogClampLog2);
            const float lv = fmaxf(lg2_approx(v), kLogClampLog2);
            const float b = fmaf(t, lu - lv, lv);  // t*lu + (1-t)*lv
            if constexpr (ALIGNED) {
                acc[0] = fmaf(u, t, acc[0]);
                acc[1] += u;
                acc[2] += t;
                acc[3] += b;
                acc[4] = fmaf(r, r, acc[4]);
                acc[5] = fmaf(dx, dx, acc[5]);
                acc[5] = fmaf(dy, dy, acc[5]);
                acc[6] = fmaf(uv, uv, acc[6]);
                if constexpr (KIND == PIL_X_PROB) acc[7] += (u >= 0.0f && u <= 1.0f) ? 0.0f : 1.0f;
            } else {
                const float w = m[p];
                acc[0] = fmaf(u * t, w, acc[0]);
                acc[1] = fmaf(u, w, acc[1]);
                acc[2] = fmaf(t, w, acc[2]);
                acc[3] = fmaf(b, w, acc[3]);
                acc[4] = fmaf(r * r, w, acc[4]);
                acc[5] = fmaf(dx * dx + dy * dy, w, acc[5]);
                acc[6] = fmaf(uv * uv, w, acc[6]);
                if constexpr (KIND == PIL_X_PROB) acc[7] += (u >= 0.0f && u <= 1.0f) ? 0.0f : w;
            }
        }
    }
};

template <int KIND, typename XT, typename TT, bool ALIGNED>
__global__ void __launch_bounds__(kThreads, kFwdMinBlocks) pil_fwd_kernel(const FwdArgs A) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Geo& g = A.g;
    const long long task = (long long)blockIdx.x * kWarpsPerBlock + warp;

    FwdRow<KIND, ALIGNED> fr;
#pragma unroll
    for (int k = 0; k < 8; ++k) fr.acc[k] = 0.f;
    fr.D = A.D;
    fr.a = A.a;

    if (task < g.tasks) {
        const int strip = (int)(task % g.strips);
        const long long tmp = task / g.strips;
        const int seg = (int)(tmp % g.segs);
        const int b = (int)(tmp / g.segs);
        const int r0 = seg * g.rps;
        const int r1 = min(r0 + g.rps, g.H);
        const int H = g.H, W = g.W;
        const int col0 = strip * kStripCols + (lane - 1) * kVec;
        const bool out_lane = (lane >= 1) && (lane <= kOutLanes);
        const bool counted = out_lane && col0 >= 0 && col0 < W;  // ALIGNED: all 4 slots are real pixels

        Cols<ALIGNED> cx;
        cx.init(col0, W);
        if constexpr (!ALIGNED) {
#pragma unroll
            for (int p = 0; p < 4; ++p) fr.m[p] = (out_lane && col0 + p >= 0 && col0 + p < W) ? 1.0f : 0.0f;
        }

        // per-image base pointers at this thread's column; rows are addressed with 32-bit offsets
        const int coff = ALIGNED ? cx.colc : 0;
        const XT* xb = reinterpret_cast<const XT*>(A.x) + (long long)b * H * W + coff;
        const TT* tb = reinterpret_cast<const TT*>(A.t) + (long long)b * H * W + coff;
        auto xrow = [&](int k) -> const XT* { return xb + (unsigned)(mirror_clamp(k, H) * W); };
        auto trow = [&](int k) -> const TT* { return tb + (unsigned)(min(k, H - 1) * W); };

        // prologue: rows r0-1, r0 become u; rows r0+1, r0+2 and targets r0, r0+1 are in flight
        const float4 x0 = cx.template load<XT>(xrow(r0 - 1)), x1 = cx.template load<XT>(xrow(r0));
        float4 xq0 = cx.template load<XT>(xrow(r0 + 1));
        float4 xq1 = cx.template load<XT>(xrow(min(r0 + 2, r1)));
        float4 tq0 = cx.template load_plain<TT>(trow(r0));
        float4 tq1 = cx.template load_plain<TT>(trow(r0 + 1));
        float4 um = act4<KIND>(x0), uc = act4<KIND>(x1);
        const XT* px = xb + (unsigned)((r0 + 3) * W);  // next map row to fetch (row i+3)
        const TT* pt = tb + (unsigned)((r0 + 2) * W);  // next target row to fetch (row i+2)

        // CHECK=false: steady state, every fetched row is inside the segment and the image.
        auto step = [&](int i, auto check) {
            constexpr bool CHECK = decltype(check)::value;
            const float4 xn = xq0, tn = tq0;
            xq0 = xq1;
            tq0 = tq1;
            if constexpr (CHECK) {
                if (i + 3 <= r1) xq1 = cx.template load<XT>((i + 3 == H) ? px - 2 * W : px);  // row H := row H-2
                if (i + 2 < r1) tq1 = cx.template load_plain<TT>(pt);
            } else {
                xq1 = cx.template load<XT>(px);
                tq1 = cx.template load_plain<TT>(pt);
            }
            px += W;
            pt += W;
            const float4 up = act4<KIND>(xn);
            fr.row(um, uc, up, tn);
            um = uc;
            uc = up;
        };
        int i = r0;
#pragma unroll 2
        for (; i < r1 - 3; ++i) step(i, BoolC<false>{});
        for (; i < r1; ++i) step(i, BoolC<true>{});

        if constexpr (ALIGNED) {
            if (!counted) {
#pragma unroll
                for (int k = 0; k < 8; ++k) fr.acc[k] = 0.f;
            }
        }
    }

    // ---- block reduction: warp shuffles, then 4 warps through shared memory --------------------
    __shared__ double s_part[kWarpsPerBlock][PIL_NSUMS];
    __shared__ bool s_last;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float v = fr.acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_part[warp][k] = (double)v;
    }
    __syncthreads();
    if (threadIdx.x < PIL_NSUMS) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kWarpsPerBlock; ++w) v += s_part[w][threadIdx.x];
        A.partials[(long long)blockIdx.x * PIL_NSUMS + threadIdx.x] = v;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(A.ticket, 1u);
        s_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;

    // ---- last block: fixed-order sum of all per-block partials (deterministic) -----------------
    __threadfence();
    __shared__ double s_red[kThreads];
    {
        const int c = threadIdx.x & 7, j = threadIdx.x >> 3;  // 16 row-groups x 8 components
        double v = 0.0;
        for (long long blk = j; blk < gridDim.x; blk += kThreads / 8)
            v += __ldcg(A.partials + blk * PIL_NSUMS + c);
        s_red[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x < PIL_NSUMS) {
        double v = 0.0;
        for (int j = 0; j < kThreads / 8; ++j) v += s_red[j * 8 + threadIdx.x];
        s_red[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double eps = A.p.epsilon;
        double s[PIL_NSUMS];
        s[0] = s_red[0];
        s[1] = s_red[1];
        s[2] = s_red[2];
        s[3] = -(double)kLn2 * s_red[3];                       // back from log2 units, BCE sign
        s[4] = s_red[4];
        s[5] = (eps / 8.0) * s_red[5] + s_red[6] / eps;        // (eps/2)*(dx/2)^2 ... + W/eps
        s[6] = s_red[7];
        s[7] = (double)g.B * (double)g.H * (double)g.W;
#pragma unroll
        for (int k = 0; k < PIL_NSUMS; ++k) A.sums[k] = s[k];
        if (A.loss_out != nullptr) finalize_device(s, s[7], A.p, A.loss_out);
        *A.ticket = 0u;
    }
}

// ------------------------------------------------------------------------------------------------
// K2: fused backward
// ------------------------------------------------------------------------------------------------
struct BwdCoef {
    float alpha, beta;   // dice: d/du = alpha*t + beta                    (du space)
    float cb;            // bce : cb*(u-t)/max(uv,1e-12)                   (du space)
    float cA;            // rd  : cA * (L^T r)      cA = scale*lrd*2/N*D
    float cF;            // rd  : cF * f'(u) * r    cF = scale*lrd*2/N
    float cG;            // pf  : cG * (dx[p-1]-dx[p+1] + dy[i-1]-dy[i+1]),  cG = scale*lpf/N*eps/4
    float cW;            // pf  : cW * uv*(1-2u),   cW = scale*lpf/N*2/eps
    float D, a;
    float fa2, fa;       // f'(u) = -3u^2 + fa2*u - fa,  fa2 = 2(1+a)
};

template <int KIND, typename XT, typename TT, bool ALIGNED>
__global__ void __launch_bounds__(kThreads, kBwdMinBlocks) pil_bwd_kernel(const BwdArgs A) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Geo& g = A.g;
    const long long task = (long long)blockIdx.x * kWarpsPerBlock + warp;
    if (task >= g.tasks) return;

    // ---- coefficients from the global sums (double once per thread, then fp32) -----------------
    BwdCoef c;
    {
        const double I = A.gsums[0], P = A.gsums[1], T = A.gsums[2];
        const double s = A.p.smooth, den = P + T + s;
        double scale = (double)A.grad_scale * (A.upstream ? (double)__ldg(A.upstream) : 1.0);
        if (KIND == PIL_X_LOGITS_TANH) scale *= 2.0;  // d/dz sigmoid(2z) = 2 u (1-u)
        const double invN = 1.0 / (A.n_global > 0 ? (double)A.n_global : A.gsums[7]);
        const bool use_rd = A.p.pde_weight > 0.0, use_pf = A.p.phase_field_weight > 0.0;
        c.alpha = (float)(scale * A.p.dice_weight * (-2.0 / den));
        c.beta = (float)(scale * A.p.dice_weight * (2.0 * I + s) / (den * den));
        c.cb = (float)(scale * A.p.bce_weight * invN);
        const double crd = use_rd ? scale * A.p.pde_weight * 2.0 * invN : 0.0;
        c.cA = (float)(crd * A.p.diffusion_coeff);
        c.cF = (float)crd;
        c.cG = use_pf ? (float)(scale * A.p.phase_field_weight * invN * A.p.epsilon * 0.25) : 0.f;
        c.cW = use_pf ? (float)(scale * A.p.phase_field_weight * invN * 2.0 / A.p.epsilon) : 0.f;
        c.D = (float)A.p.diffusion_coeff;
        c.a = (float)A.p.reaction_threshold;
        c.fa2 = 2.0f * (1.0f + c.a);
        c.fa = c.a;
    }

    const int strip = (int)(task % g.strips);
    const long long tmp = task / g.strips;
    const int seg = (int)(tmp % g.segs);
    const int b = (int)(tmp / g.segs);
    const int r0 = seg * g.rps;
    const int r1 = min(r0 + g.rps, g.H);
    const int H = g.H, W = g.W;
    const int col0 = strip * kStripCols + (lane - 1) * kVec;
    const bool out_lane = (lane >= 1) && (lane <= kOutLanes);
    const bool in_img = col0 >= 0 && col0 < W;  // ALIGNED: whole vector in the image
    const bool edge_warp = (strip == 0) || (strip == g.strips - 1);

    Cols<ALIGNED> cx;
    cx.init(col0, W);
    // column factors of the transposed reflect stencils:
    //   fc: 2 on the first/last image column, 0 outside the image, 1 elsewhere (for r)
    //   mc: 0 outside the image, 1 inside (for dx)
    float fc[4], mc[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int cc = col0 + p;
        const bool in = cc >= 0 && cc < W;
        mc[p] = in ? 1.0f : 0.0f;
        fc[p] = in ? ((cc == 0 || cc == W - 1) ? 2.0f : 1.0f) : 0.0f;
    }
    const bool store_vec = ALIGNED && out_lane && in_img;

    const int coff = ALIGNED ? cx.colc : 0;
    const XT* xb = reinterpret_cast<const XT*>(A.x) + (long long)b * H * W + coff;
    const TT* tb = reinterpret_cast<const TT*>(A.t) + (long long)b * H * W + coff;
    XT* gb = reinterpret_cast<XT*>(A.grad) + (long long)b * H * W;  // column added at the store
    auto xrow = [&](int k) -> const XT* { return xb + (unsigned)(mirror_clamp(k, H) * W); };
    auto trow = [&](int k) -> const TT* { return tb + (unsigned)(min(k, H - 1) * W); };

    // iteration k forms the residual row k (needs u rows k-1, k, k+1) and emits gradient row k-1.
    // k runs r0-1 .. r1; u rows r0-2 .. r1+1 are read (mirrored at the image edge).
    const int k0 = r0 - 1;
    float4 ua = act4<KIND>(cx.template load<XT>(xrow(k0 - 1)));  // row k-1
    float4 ub = act4<KIND>(cx.template load<XT>(xrow(k0)));      // row k
    float4 xq0 = cx.template load<XT>(xrow(k0 + 1));             // row k+1 (consumed by the first iteration)
    float4 xq1 = cx.template load<XT>(xrow(k0 + 2));
    float4 tq0 = cx.template load_plain<TT>(trow(r0));           // consumed when row r0 is emitted (k = r0+1)
    float4 tq1 = cx.template load_plain<TT>(trow(r0 + 1));
    const XT* px = xb + (unsigned)((k0 + 3) * W);                // next map row to fetch (row k+3)
    const TT* pt = tb + (unsigned)((r0 + 2) * W);                // next target row to fetch (row k+1 at k = r0+1)
    XT* pg = gb + (unsigned)(r0 * W) + (ALIGNED ? col0 : 0);     // next gradient row to store (row k-1)
    float gm[4] = {0.f, 0.f, 0.f, 0.f}, g0[4] = {0.f, 0.f, 0.f, 0.f}, gp[4];

    // CHECK=false: steady state -- row k strictly inside the image and the segment, all fetches valid.
    auto step = [&](int k, auto check) {
        constexpr bool CHECK = decltype(check)::value;
        const float4 xn = xq0;
        xq0 = xq1;
        if constexpr (CHECK) {
            if (k + 3 <= min(r1 + 1, H)) xq1 = cx.template load<XT>((k + 3 == H) ? px - 2 * W : px);  // row H := row H-2
        } else {
            xq1 = cx.template load<XT>(px);
        }
        px += W;
        const float4 uc4 = act4<KIND>(xn);  // row k+1

        const float va[4] = {ua.x, ua.y, ua.z, ua.w};
        const float vc[4] = {uc4.x, uc4.y, uc4.z, uc4.w};
        if (!CHECK || (k >= 0 && k < H)) {
            const float L = __shfl_up_sync(0xffffffffu, ub.w, 1);
            const float R = __shfl_down_sync(0xffffffffu, ub.x, 1);
            const float e[6] = {L, ub.x, ub.y, ub.z, ub.w, R};
            float r[4], rc[4], dx[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const float u = e[p + 1];
                const float uv = u * (1.0f - u);
                const float lap = (e[p] + e[p + 2]) + (va[p] + vc[p]) - 4.0f * u;
                r[p] = fmaf(c.D, lap, uv * (u - c.a));
                rc[p] = r[p];
                dx[p] = e[p + 2] - e[p];
            }
            if (!ALIGNED || edge_warp) {
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    rc[p] *= fc[p];
                    dx[p] *= mc[p];
                }
            }
            const float rL = __shfl_up_sync(0xffffffffu, rc[3], 1);
            const float rR = __shfl_down_sync(0xffffffffu, rc[0], 1);
            const float dL = __shfl_up_sync(0xffffffffu, dx[3], 1);
            const float dR = __shfl_down_sync(0xffffffffu, dx[0], 1);
            const float re[6] = {rL, rc[0], rc[1], rc[2], rc[3], rR};
            const float de[6] = {dL, dx[0], dx[1], dx[2], dx[3], dR};
            // row factor of the transposed vertical stencil; dy of an edge row is 0 by mirroring
            const float cAr = (CHECK && (k == 0 || k == H - 1)) ? 2.0f * c.cA : c.cA;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const float u = e[p + 1];
                const float q = cAr * r[p];
                const float ey = c.cG * (vc[p] - va[p]);
                gp[p] = q + ey;   // into row k+1
                gm[p] += q - ey;  // into row k-1
                const float fprime = fmaf(u, fmaf(-3.0f, u, c.fa2), -c.fa);
                float acc = g0[p];
                acc = fmaf(c.cA, (re[p] + re[p + 2]) - 4.0f * r[p], acc);
                acc = fmaf(c.cF * fprime, r[p], acc);
                acc = fmaf(c.cG, de[p] - de[p + 2], acc);
                g0[p] = acc;
            }
        } else {
#pragma unroll
            for (int p = 0; p < 4; ++p) gp[p] = 0.f;
        }

        if (!CHECK || k - 1 >= r0) {
            // emit gradient row k-1 : pointwise terms + accumulated stencil terms, then the chain factor
            const float4 tn = tq0;
            tq0 = tq1;
            if constexpr (CHECK) {
                if (k + 1 < r1) tq1 = cx.template load_plain<TT>(pt);
            } else {
                tq1 = cx.template load_plain<TT>(pt);
            }
            pt += W;
            const float vt[4] = {tn.x, tn.y, tn.z, tn.w};
            float o[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const float u = va[p], t = vt[p];
                const float v = 1.0f - u;
                const float uv = u * v;
                float du = gm[p] + fmaf(c.alpha, t, c.beta);
                du = fmaf(c.cW * uv, v - u, du);
                if constexpr (KIND == PIL_X_PROB) {
                    o[p] = fmaf(c.cb * (u - t), rcp_approx(fmaxf(uv, 1e-12f)), du);
                } else {
                    // (u-t)/max(uv,1e-12) * uv  ==  (u-t) * sat(uv*1e12)
                    o[p] = fmaf(du, uv, c.cb * (u - t) * __saturatef(uv * 1e12f));
                }
            }
            if constexpr (ALIGNED) {
                if (store_vec) st4<XT>(pg, make_float4(o[0], o[1], o[2], o[3]));
            } else {
                if (out_lane) {
#pragma unroll
                    for (int p = 0; p < 4; ++p)
                        if (col0 + p >= 0 && col0 + p < W) st1<XT>(pg + col0 + p, o[p]);
                }
            }
            pg += W;
        }
        ua = ub;
        ub = uc4;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            gm[p] = g0[p];
            g0[p] = gp[p];
        }
    };

    int k = k0;
    for (; k <= r1 && k <= r0; ++k) step(k, BoolC<true>{});
#pragma unroll 2
    for (; k <= r1 - 4; ++k) step(k, BoolC<false>{});
    for (; k <= r1; ++k) step(k, BoolC<true>{});
}

// ------------------------------------------------------------------------------------------------
// small kernels: finalize, workspace init, stand-alone PDERegularization operators
// ------------------------------------------------------------------------------------------------
__global__ void pil_finalize_kernel(const double* sums, long long n_global, PilParams p, float* out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s[PIL_NSUMS];
        for (int k = 0; k < PIL_NSUMS; ++k) s[k] = sums[k];
        finalize_device(s, n_global > 0 ? (double)n_global : s[7], p, out);
    }
}

// one thread per pixel; neighbours through L1/L2.  These operators are the reference's public
// PDERegularization methods (used by src/ablation.py:53-86 and for logging), not the fused hot path.
enum StencilOp { OP_LAP = 0, OP_LAP_ADJ = 1, OP_GMS = 2, OP_GMS_BWD = 3 };

template <int OP>
__global__ void __launch_bounds__(256) pil_stencil_kernel(const float* __restrict__ u, const float* __restrict__ gin,
                                                          float* __restrict__ out, int B, int H, int W) {
    const long long n = (long long)B * H * W;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(idx % W);
        const long long rowi = idx / W;
        const int i = (int)(rowi % H);
        const float* img = u + (rowi - i) * W;
        auto at = [&](const float* base, int ii, int jj) { return __ldg(base + (long long)ii * W + jj); };
        float res;
        if constexpr (OP == OP_LAP) {
            res = at(img, mirror_clamp(i - 1, H), j) + at(img, i, mirror_clamp(j - 1, W)) - 4.0f * at(img, i, j) +
                  at(img, i, mirror_clamp(j + 1, W)) + at(img, mirror_clamp(i + 1, H), j);
        } else if constexpr (OP == OP_LAP_ADJ) {
            // (L^T g)[i,j] = sum over pixels that read (i,j): doubled edge rows/cols, nothing outside
            auto fr = [&](int k, int n_) { return (k < 0 || k >= n_) ? 0.0f : ((k == 0 || k == n_ - 1) ? 2.0f : 1.0f); };
            float acc = -4.0f * at(img, i, j);
            if (i - 1 >= 0) acc += fr(i - 1, H) * at(img, i - 1, j);
            if (i + 1 < H) acc += fr(i + 1, H) * at(img, i + 1, j);
            if (j - 1 >= 0) acc += fr(j - 1, W) * at(img, i, j - 1);
            if (j + 1 < W) acc += fr(j + 1, W) * at(img, i, j + 1);
            res = acc;
        } else if constexpr (OP == OP_GMS) {
            const float gx = 0.5f * at(img, i, mirror_clamp(j + 1, W)) - 0.5f * at(img, i, mirror_clamp(j - 1, W));
            const float gy = 0.5f * at(img, mirror_clamp(i + 1, H), j) - 0.5f * at(img, mirror_clamp(i - 1, H), j);
            res = gx * gx + gy * gy;
        } else {
            // out[i,j] = sum_k g[k] * d(gx_k^2+gy_k^2)/du[i,j];  gx,gy vanish on edge columns/rows
            const float* gimg = gin + (rowi - i) * W;
            auto gxg = [&](int ii, int jj) -> float {  // g*gx at (ii,jj), 0 outside / on edge columns
                if (jj <= 0 || jj >= W - 1) return 0.0f;
                return at(gimg, ii, jj) * (0.5f * at(img, ii, jj + 1) - 0.5f * at(img, ii, jj - 1));
            };
            auto gyg = [&](int ii, int jj) -> float {
                if (ii <= 0 || ii >= H - 1) return 0.0f;
                return at(gimg, ii, jj) * (0.5f * at(img, ii + 1, jj) - 0.5f * at(img, ii - 1, jj));
            };
            res = (gxg(i, j - 1) - gxg(i, j + 1)) + (gyg(i - 1, j) - gyg(i + 1, j));
        }
        out[idx] = res;
    }
}

__global__ void __launch_bounds__(256) pil_reaction_kernel(const float* __restrict__ u, float* __restrict__ out,
                                                           long long n, float a) {
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (long long)gridDim.x * blockDim.x) {
        const float v = __ldg(u + idx);
        out[idx] = v * (1.0f - v) * (v - a);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static thread_local PilLaunchInfo t_info = {};
static long long g_kernels_launched = 0;
static int g_tune_fwd_rps = 0, g_tune_bwd_rps = 0;

static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

// Choose rows-per-segment: every warp-task costs (rps + warm) row-iterations; tasks run in
// ceil(tasks / resident_warps) rounds.  Minimise rounds * (rps + warm); ties go to longer segments.
static Geo make_geo(int64_t B, int64_t H, int64_t W, int warm_rows, int resident_warps, int forced_rps) {
    Geo g;
    g.B = (int)B;
    g.H = (int)H;
    g.W = (int)W;
    g.strips = (int)((W + kStripCols - 1) / kStripCols);
    long long best_cost = -1;
    int best_segs = 1;
    if (forced_rps > 0) {
        best_segs = (int)((H + forced_rps - 1) / forced_rps);
    } else {
        const int max_segs = (int)((H + 7) / 8);  // at least 8 rows per segment
        for (int segs = 1; segs <= max_segs; ++segs) {
            const int rps = (int)((H + segs - 1) / segs);
            const int real_segs = (int)((H + rps - 1) / rps);
            if (real_segs != segs) continue;
            const long long tasks = (long long)B * g.strips * segs;
            const long long rounds = (tasks + resident_warps - 1) / resident_warps;
            const long long cost = rounds * (rps + warm_rows);
            if (best_cost < 0 || cost < best_cost) {
                best_cost = cost;
                best_segs = segs;
            }
        }
    }
    g.rps = (int)((H + best_segs - 1) / best_segs);
    g.segs = (int)((H + g.rps - 1) / g.rps);
    g.tasks = (long long)B * g.strips * g.segs;
    return g;
}

static int check_common(const void* x, const void* t, int64_t B, int64_t H, int64_t W, int x_dtype, int t_dtype,
                        int x_kind, const PilParams* p) {
    if (!x || !t || !p) return PIL_ERR_NULL;
    if (B < 1 || H < 2 || W < 2 || B * H * W > (int64_t)1 << 40 || H > (1 << 30) || W > (1 << 30)) return PIL_ERR_SHAPE;
    if (!(x_dtype == PIL_F32 || x_dtype == PIL_BF16)) return PIL_ERR_DTYPE;
    if (!(t_dtype == PIL_F32 || t_dtype == PIL_BF16 || t_dtype == PIL_U8)) return PIL_ERR_DTYPE;
    if (x_kind < PIL_X_PROB || x_kind > PIL_X_LOGITS_TANH) return PIL_ERR_KIND;
    const uintptr_t xa = (x_dtype == PIL_F32) ? 4 : 2, ta = (t_dtype == PIL_F32) ? 4 : (t_dtype == PIL_BF16 ? 2 : 1);
    if (((uintptr_t)x % xa) || ((uintptr_t)t % ta)) return PIL_ERR_ALIGNMENT;
    return pil_validate_params(p);
}

static size_t dtype_size(int d) { return d == PIL_F32 ? 4 : (d == PIL_BF16 ? 2 : 1); }

static bool is_aligned_case(const void* x, const void* t, const void* gptr, int64_t W, int x_dtype, int t_dtype) {
    if (W % 4) return false;
    if ((uintptr_t)x % (4 * dtype_size(x_dtype))) return false;
    if ((uintptr_t)t % (4 * dtype_size(t_dtype))) return false;
    if (gptr && ((uintptr_t)gptr % (4 * dtype_size(x_dtype)))) return false;
    return true;
}

template <int KIND, typename XT, typename TT>
static cudaError_t launch_fwd_a(const FwdArgs& a, bool aligned, int blocks, cudaStream_t s) {
    if (aligned)
        pil_fwd_kernel<KIND, XT, TT, true><<<blocks, kThreads, 0, s>>>(a);
    else
        pil_fwd_kernel<KIND, XT, TT, false><<<blocks, kThreads, 0, s>>>(a);
    return cudaGetLastError();
}
template <int KIND, typename XT>
static cudaError_t launch_fwd_t(const FwdArgs& a, int t_dtype, bool aligned, int blocks, cudaStream_t s) {
    switch (t_dtype) {
        case PIL_F32: return launch_fwd_a<KIND, XT, float>(a, aligned, blocks, s);
        case PIL_BF16: return launch_fwd_a<KIND, XT, __nv_bfloat16>(a, aligned, blocks, s);
        default: return launch_fwd_a<KIND, XT, uint8_t>(a, aligned, blocks, s);
    }
}
template <int KIND>
static cudaError_t launch_fwd_x(const FwdArgs& a, int x_dtype, int t_dtype, bool aligned, int blocks, cudaStream_t s) {
    if (x_dtype == PIL_F32) return launch_fwd_t<KIND, float>(a, t_dtype, aligned, blocks, s);
    return launch_fwd_t<KIND, __nv_bfloat16>(a, t_dtype, aligned, blocks, s);
}

template <int KIND, typename XT, typename TT>
static cudaError_t launch_bwd_a(const BwdArgs& a, bool aligned, int blocks, cudaStream_t s) {
    if (aligned)
        pil_bwd_kernel<KIND, XT, TT, true><<<blocks, kThreads, 0, s>>>(a);
    else
        pil_bwd_kernel<KIND, XT, TT, false><<<blocks, kThreads, 0, s>>>(a);
    return cudaGetLastError();
}
template <int KIND, typename XT>
static cudaError_t launch_bwd_t(const BwdArgs& a, int t_dtype, bool aligned, int blocks, cudaStream_t s) {
    switch (t_dtype) {
        case PIL_F32: return launch_bwd_a<KIND, XT, float>(a, aligned, blocks, s);
        case PIL_BF16: return launch_bwd_a<KIND, XT, __nv_bfloat16>(a, aligned, blocks, s);
        default: return launch_bwd_a<KIND, XT, uint8_t>(a, aligned, blocks, s);
    }
}
template <int KIND>
static cudaError_t launch_bwd_x(const BwdArgs& a, int x_dtype, int t_dtype, bool aligned, int blocks, cudaStream_t s) {
    if (x_dtype == PIL_F32) return launch_bwd_t<KIND, float>(a, t_dtype, aligned, blocks, s);
    return launch_bwd_t<KIND, __nv_bfloat16>(a, t_dtype, aligned, blocks, s);
}

struct WorkspaceLayout {
    size_t ticket_off, partials_off, total;
};
static WorkspaceLayout workspace_layout(int64_t B, int64_t H, int64_t W) {
    // worst case number of forward blocks: 8-row segments
    const long long strips = (W + kStripCols - 1) / kStripCols;
    const long long segs = (H + 7) / 8;
    const long long blocks = (B * strips * segs + kWarpsPerBlock - 1) / kWarpsPerBlock;
    WorkspaceLayout l;
    l.ticket_off = 0;
    l.partials_off = 256;
    l.total = l.partials_off + (size_t)blocks * PIL_NSUMS * sizeof(double);
    return l;
}

}  // namespace pil

using namespace pil;

extern "C" {

int pil_version(void) { return PIL_VERSION; }

const char* pil_status_string(int status) {
    switch (status) {
        case PIL_OK: return "ok";
        case PIL_ERR_NULL: return "a required pointer is NULL";
        case PIL_ERR_SHAPE: return "bad shape: need B >= 1, H >= 2, W >= 2 (reflect padding)";
        case PIL_ERR_DTYPE: return "unsupported dtype";
        case PIL_ERR_KIND: return "unknown input kind";
        case PIL_ERR_WORKSPACE: return "workspace too small or misaligned";
        case PIL_ERR_DIFFUSION: return "diffusion_coeff must be positive";
        case PIL_ERR_THRESHOLD: return "reaction_threshold must be in (0,1)";
        case PIL_ERR_EPSILON: return "epsilon must be positive";
        case PIL_ERR_ALIGNMENT: return "pointer not aligned to its element size";
        case PIL_ERR_SESSION: return "session misuse";
        default: break;
    }
    if (status > 0) return cudaGetErrorString((cudaError_t)status);
    return "unknown pil status";
}

int pil_validate_params(const PilParams* p) {
    if (!p) return PIL_ERR_NULL;
    if (!(p->diffusion_coeff > 0.0)) return PIL_ERR_DIFFUSION;                              // src/pde.py:14-15
    if (!(p->reaction_threshold > 0.0 && p->reaction_threshold < 1.0)) return PIL_ERR_THRESHOLD;  // src/pde.py:16-17
    if (p->phase_field_weight > 0.0 && !(p->epsilon > 0.0)) return PIL_ERR_EPSILON;         // src/pde.py:199-200 via src/loss.py:155
    return PIL_OK;
}

size_t pil_workspace_bytes(int64_t B, int64_t H, int64_t W) {
    if (B < 1 || H < 2 || W < 2) return 0;
    return workspace_layout(B, H, W).total;
}

int pil_workspace_init(void* workspace, size_t workspace_bytes, void* stream) {
    if (!workspace) return PIL_ERR_NULL;
    if (workspace_bytes < 256) return PIL_ERR_WORKSPACE;
    return (int)cudaMemsetAsync(workspace, 0, 256, (cudaStream_t)stream);
}

int pil_forward(const void* x, const void* t, int64_t B, int64_t H, int64_t W, int x_dtype, int t_dtype, int x_kind,
                const PilParams* p, double* sums, float* loss_out, void* workspace, size_t workspace_bytes,
                void* stream) {
    int st = check_common(x, t, B, H, W, x_dtype, t_dtype, x_kind, p);
    if (st != PIL_OK) return st;
    if (!sums || !workspace) return PIL_ERR_NULL;
    const WorkspaceLayout wl = workspace_layout(B, H, W);
    if (workspace_bytes < wl.total || ((uintptr_t)workspace % 8)) return PIL_ERR_WORKSPACE;

    const int resident = sm_count() * kFwdMinBlocks * kWarpsPerBlock;
    FwdArgs a;
    a.x = x;
    a.t = t;
    a.g = make_geo(B, H, W, /*warm_rows=*/2, resident, g_tune_fwd_rps);
    a.D = (float)p->diffusion_coeff;
    a.a = (float)p->reaction_threshold;
    a.ticket = reinterpret_cast<unsigned int*>((char*)workspace + wl.ticket_off);
    a.partials = reinterpret_cast<double*>((char*)workspace + wl.partials_off);
    a.sums = sums;
    a.loss_out = loss_out;
    a.p = *p;
    const int blocks = (int)((a.g.tasks + kWarpsPerBlock - 1) / kWarpsPerBlock);
    if (wl.partials_off + (size_t)blocks * PIL_NSUMS * sizeof(double) > workspace_bytes) return PIL_ERR_WORKSPACE;
    const bool aligned = is_aligned_case(x, t, nullptr, W, x_dtype, t_dtype);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e;
    switch (x_kind) {
        case PIL_X_PROB: e = launch_fwd_x<PIL_X_PROB>(a, x_dtype, t_dtype, aligned, blocks, s); break;
        case PIL_X_LOGITS_SIGMOID: e = launch_fwd_x<PIL_X_LOGITS_SIGMOID>(a, x_dtype, t_dtype, aligned, blocks, s); break;
        default: e = launch_fwd_x<PIL_X_LOGITS_TANH>(a, x_dtype, t_dtype, aligned, blocks, s); break;
    }
    t_info.fwd_blocks = blocks;
    t_info.fwd_threads = kThreads;
    t_info.fwd_rows_per_segment = a.g.rps;
    t_info.fwd_aligned = aligned ? 1 : 0;
    ++g_kernels_launched;
    return (int)e;
}

int pil_finalize(const double* sums, int64_t n_global, const PilParams* p, float* loss_out, void* stream) {
    if (!sums || !p || !loss_out) return PIL_ERR_NULL;
    pil_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, (long long)n_global, *p, loss_out);
    ++g_kernels_launched;
    return (int)cudaGetLastError();
}

int pil_backward(const void* x, const void* t, void* grad, int64_t B, int64_t H, int64_t W, int x_dtype, int t_dtype,
                 int x_kind, const PilParams* p, const double* global_sums, int64_t n_global, const float* upstream,
                 float grad_scale, void* stream) {
    int st = check_common(x, t, B, H, W, x_dtype, t_dtype, x_kind, p);
    if (st != PIL_OK) return st;
    if (!grad || !global_sums) return PIL_ERR_NULL;
    if ((uintptr_t)grad % dtype_size(x_dtype)) return PIL_ERR_ALIGNMENT;

    const int resident = sm_count() * kBwdMinBlocks * kWarpsPerBlock;
    BwdArgs a;
    a.x = x;
    a.t = t;
    a.grad = grad;
    a.g = make_geo(B, H, W, /*warm_rows=*/4, resident, g_tune_bwd_rps);
    a.gsums = global_sums;
    a.upstream = upstream;
    a.grad_scale = grad_scale;
    a.n_global = (long long)n_global;
    a.p = *p;
    const int blocks = (int)((a.g.tasks + kWarpsPerBlock - 1) / kWarpsPerBlock);
    const bool aligned = is_aligned_case(x, t, grad, W, x_dtype, t_dtype);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e;
    switch (x_kind) {
        case PIL_X_PROB: e = launch_bwd_x<PIL_X_PROB>(a, x_dtype, t_dtype, aligned, blocks, s); break;
        case PIL_X_LOGITS_SIGMOID: e = launch_bwd_x<PIL_X_LOGITS_SIGMOID>(a, x_dtype, t_dtype, aligned, blocks, s); break;
        default: e = launch_bwd_x<PIL_X_LOGITS_TANH>(a, x_dtype, t_dtype, aligned, blocks, s); break;
    }
    t_info.bwd_blocks = blocks;
    t_info.bwd_threads = kThreads;
    t_info.bwd_rows_per_segment = a.g.rps;
    t_info.bwd_aligned = aligned ? 1 : 0;
    ++g_kernels_launched;
    return (int)e;
}

static int stencil_launch(int op, const float* u, const float* g, float* out, int64_t B, int64_t H, int64_t W,
                          void* stream) {
    if (!u || !out || (op == OP_GMS_BWD && !g)) return PIL_ERR_NULL;
    if (B < 1 || H < 2 || W < 2) return PIL_ERR_SHAPE;
    const long long n = (long long)B * H * W;
    const int blocks = (int)((n + 255) / 256 < (long long)sm_count() * 16 ? (n + 255) / 256 : (long long)sm_count() * 16);
    cudaStream_t s = (cudaStream_t)stream;
    switch (op) {
        case OP_LAP: pil_stencil_kernel<OP_LAP><<<blocks, 256, 0, s>>>(u, nullptr, out, (int)B, (int)H, (int)W); break;
        case OP_LAP_ADJ: pil_stencil_kernel<OP_LAP_ADJ><<<blocks, 256, 0, s>>>(u, nullptr, out, (int)B, (int)H, (int)W); break;
        case OP_GMS: pil_stencil_kernel<OP_GMS><<<blocks, 256, 0, s>>>(u, nullptr, out, (int)B, (int)H, (int)W); break;
        default: pil_stencil_kernel<OP_GMS_BWD><<<blocks, 256, 0, s>>>(u, g, out, (int)B, (int)H, (int)W); break;
    }
    ++g_kernels_launched;
    return (int)cudaGetLastError();
}

int pil_laplacian(const float* u, float* out, int64_t B, int64_t H, int64_t W, void* stream) {
    return stencil_launch(OP_LAP, u, nullptr, out, B, H, W, stream);
}
int pil_laplacian_adjoint(const float* g, float* out, int64_t B, int64_t H, int64_t W, void* stream) {
    return stencil_launch(OP_LAP_ADJ, g, nullptr, out, B, H, W, stream);
}
int pil_grad_mag_sq(const float* u, float* out, int64_t B, int64_t H, int64_t W, void* stream) {
    return stencil_launch(OP_GMS, u, nullptr, out, B, H, W, stream);
}
int pil_grad_mag_sq_backward(const float* u, const float* g, float* out, int64_t B, int64_t H, int64_t W,
                             void* stream) {
    return stencil_launch(OP_GMS_BWD, u, g, out, B, H, W, stream);
}
int pil_reaction(const float* u, float* out, int64_t n, double reaction_threshold, void* stream) {
    if (!u || !out) return PIL_ERR_NULL;
    if (n < 1) return PIL_ERR_SHAPE;
    const int blocks = (int)((n + 255) / 256 < (long long)sm_count() * 16 ? (n + 255) / 256 : (long long)sm_count() * 16);
    pil_reaction_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(u, out, (long long)n, (float)reaction_threshold);
    ++g_kernels_launched;
    return (int)cudaGetLastError();
}

int pil_last_launch_info(PilLaunchInfo* out) {
    if (!out) return PIL_ERR_NULL;
    *out = t_info;
    out->kernels_launched = g_kernels_launched;
    return PIL_OK;
}

int pil_set_tuning(int fwd_rows_per_segment, int bwd_rows_per_segment) {
    g_tune_fwd_rps = fwd_rows_per_segment > 0 ? fwd_rows_per_segment : 0;
    g_tune_bwd_rps = bwd_rows_per_segment > 0 ? bwd_rows_per_segment : 0;
    return PIL_OK;
}

}  // extern "C"
