// pil_common.cuh -- device helpers, argument structs and launch plumbing shared by the translation units of
// libpil.so: the fused sm_100a kernels of the physics-prior loss behind the C ABI of include/pil.h
// (pil_fwd.cu, pil_point.cu, pil_bwd.cu, pil_tail.cu, pil_boundary.cu, pil_api.cu, pil_session.cu).
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
#ifndef PIL_COMMON_CUH_
#define PIL_COMMON_CUH_
#include <cuda.h>  // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "pil.h"

#include <atomic>
namespace pil {

constexpr int kVec = 4;                     // columns per thread
constexpr int kOutLanes = 30;               // lanes of a warp that own output columns
constexpr int kStripCols = kOutLanes * kVec;  // 120 output columns per warp
constexpr int kWarpsPerBlock = 4;
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr int kFwdMinBlocks = 7;  // 28 warps / SM, <= 73 registers
constexpr int kBwdMinBlocks = 4;  // 16 warps / SM, <= 128 registers.  Measured at 64x1024^2: 145 us; 3 blocks (158 registers,
                                  // 675-instruction loop) 155 us; 5 blocks (96 registers, spills, 747-instruction loop) 180 us
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kLogClampLog2 = -100.0f * kLog2e;  // nn.BCELoss clamps ln() at -100

static_assert(PIL_NSUMS == 8, "sums layout");

// Development build -DPIL_BOUNDS (tools/bounds_check.py; compute-sanitizer is not available on the GPU
// pool): every global load / cp.async source / store of the fused kernels is checked against the extents of
// the tensors of the current call; violations are counted, not trapped.
#ifdef PIL_BOUNDS
#ifdef PIL_KIND
#error "PIL_BOUNDS is a single-translation-unit (unity) development build: compile csrc/pil_unity.cu"
#endif
__device__ const char* g_brd[4] = {nullptr, nullptr, nullptr, nullptr};  // x0, x1, t0, t1 (byte extents)
__device__ const char* g_bwr[2] = {nullptr, nullptr};                    // grad0, grad1
__device__ unsigned long long g_berr[4] = {0, 0, 0, 0};                  // bad reads, bad writes, first bad address, -
__device__ __forceinline__ void chk_rd(const void* p, int bytes) {
    const char* c = reinterpret_cast<const char*>(p);
    const bool ok = (c >= g_brd[0] && c + bytes <= g_brd[1]) || (c >= g_brd[2] && c + bytes <= g_brd[3]);
    if (!ok && atomicAdd(&g_berr[0], 1ull) == 0) g_berr[2] = (unsigned long long)c;
}
__device__ __forceinline__ void chk_wr(const void* p, int bytes) {
    const char* c = reinterpret_cast<const char*>(p);
    if (!(c >= g_bwr[0] && c + bytes <= g_bwr[1]) && atomicAdd(&g_berr[1], 1ull) == 0) g_berr[2] = (unsigned long long)c;
}
#define PIL_CHK_RD(p, n) chk_rd(p, n)
#define PIL_CHK_WR(p, n) chk_wr(p, n)
#else
#define PIL_CHK_RD(p, n)
#define PIL_CHK_WR(p, n)
#endif

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
    PIL_CHK_RD(p, 4);
    return __ldg(p);
}
template <>
__device__ __forceinline__ float ld1<__nv_bfloat16>(const __nv_bfloat16* p) {
    PIL_CHK_RD(p, 2);
    return __bfloat162float(*p);
}
template <>
__device__ __forceinline__ float ld1<uint8_t>(const uint8_t* p) {
    PIL_CHK_RD(p, 1);
    return (float)__ldg(p);
}

template <typename T>
__device__ __forceinline__ float4 ld4(const T* p);
template <>
__device__ __forceinline__ float4 ld4<float>(const float* p) {
    PIL_CHK_RD(p, 16);
    return __ldg(reinterpret_cast<const float4*>(p));
}
template <>
__device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
    PIL_CHK_RD(p, 8);
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
    PIL_CHK_RD(p, 4);
    const uint32_t raw = __ldg(reinterpret_cast<const uint32_t*>(p));
    return make_float4((float)(raw & 0xff), (float)((raw >> 8) & 0xff), (float)((raw >> 16) & 0xff),
                       (float)(raw >> 24));
}

template <typename T>
__device__ __forceinline__ void st4(T* p, float4 v);
template <>
__device__ __forceinline__ void st4<float>(float* p, float4 v) {
    PIL_CHK_WR(p, 16);
    __stcs(reinterpret_cast<float4*>(p), v);
}
template <>
__device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
    PIL_CHK_WR(p, 8);
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
    PIL_CHK_WR(p, 4);
    *p = v;
}
template <>
__device__ __forceinline__ void st1<__nv_bfloat16>(__nv_bfloat16* p, float v) {
    PIL_CHK_WR(p, 2);
    *p = __float2bfloat16_rn(v);
}

// ------------------------------------------------------------------------------------------------
// cp.async stage ring (ALIGNED path): every lane copies its own 16-byte (4-column) piece of the rows
// it will need kStages-1 iterations ahead into a private shared-memory slot (LDGSTS, no registers
// held while the load is in flight) and reads it back with one LDS when the row is consumed.  A lane
// only ever reads what it copied itself, so cp.async.wait_group is the only synchronisation needed.
// ------------------------------------------------------------------------------------------------
constexpr int kStages = 6;                       // == unroll factor of the steady-state loops
constexpr int kStageBytes = 2 * 32 * 16;         // one map row piece + one target row piece per lane
constexpr int kSmemPerWarp = kStages * kStageBytes;
constexpr int kSmemPerBlock = kWarpsPerBlock * kSmemPerWarp;

template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t dst_shared, const void* src) {
    PIL_CHK_RD(src, BYTES);
    if constexpr (BYTES == 16) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_shared), "l"(src) : "memory");
    } else {
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst_shared), "l"(src), "n"(BYTES) : "memory");
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <typename T>
__device__ __forceinline__ float4 lds4(const unsigned char* p);
template <>
__device__ __forceinline__ float4 lds4<float>(const unsigned char* p) {
    return *reinterpret_cast<const float4*>(p);
}
template <>
__device__ __forceinline__ float4 lds4<__nv_bfloat16>(const unsigned char* p) {
    const uint2 raw = *reinterpret_cast<const uint2*>(p);
    float4 r;
    r.x = __uint_as_float(raw.x << 16);
    r.y = __uint_as_float(raw.x & 0xffff0000u);
    r.z = __uint_as_float(raw.y << 16);
    r.w = __uint_as_float(raw.y & 0xffff0000u);
    return r;
}
template <>
__device__ __forceinline__ float4 lds4<uint8_t>(const unsigned char* p) {
    const uint32_t raw = *reinterpret_cast<const uint32_t*>(p);
    return make_float4((float)(raw & 0xff), (float)((raw >> 8) & 0xff), (float)((raw >> 16) & 0xff),
                       (float)(raw >> 24));
}

// the same three readers on a 32-bit shared-space address (TMA ring: no generic-pointer arithmetic in the loop);
// volatile keeps them behind the mbarrier wait that precedes them
template <typename T>
__device__ __forceinline__ float4 lds4s(uint32_t a);
template <>
__device__ __forceinline__ float4 lds4s<float>(uint32_t a) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(a));
    return r;
}
template <>
__device__ __forceinline__ float4 lds4s<__nv_bfloat16>(uint32_t a) {
    uint2 raw;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(raw.x), "=r"(raw.y) : "r"(a));
    float4 r;
    r.x = __uint_as_float(raw.x << 16);
    r.y = __uint_as_float(raw.x & 0xffff0000u);
    r.z = __uint_as_float(raw.y << 16);
    r.w = __uint_as_float(raw.y & 0xffff0000u);
    return r;
}
template <>
__device__ __forceinline__ float4 lds4s<uint8_t>(uint32_t a) {
    uint32_t raw;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(raw) : "r"(a));
    return make_float4((float)(raw & 0xff), (float)((raw >> 8) & 0xff), (float)((raw >> 16) & 0xff), (float)(raw >> 24));
}

template <typename XT, typename TT>
struct StageRing {
    unsigned char* my;   // generic pointer to this lane's slot of stage 0 (map piece); target piece at +512
    uint32_t my_s;       // same, shared-space address for cp.async
    __device__ __forceinline__ void init(unsigned char* smem, int warp, int lane) {
        my = smem + warp * kSmemPerWarp + lane * 16;
        my_s = (uint32_t)__cvta_generic_to_shared(my);
    }
    __device__ __forceinline__ void issue_x(int stage, const XT* src) const {
        cp_async<4 * (int)sizeof(XT)>(my_s + stage * kStageBytes, src);
    }
    __device__ __forceinline__ void issue_t(int stage, const TT* src) const {
        cp_async<4 * (int)sizeof(TT)>(my_s + stage * kStageBytes + 512, src);
    }
    __device__ __forceinline__ float4 read_x(int stage) const { return lds4<XT>(my + stage * kStageBytes); }
    __device__ __forceinline__ float4 read_t(int stage) const { return lds4<TT>(my + stage * kStageBytes + 512); }
};

// ------------------------------------------------------------------------------------------------
// TMA stage ring (ALIGNED path, Blackwell-native alternative to the cp.async ring above): a warp's rows arrive
// as 2-D tensor-map BOXES of kBoxRows rows x 128 columns -- one cp.async.bulk.tensor for the map and one for
// the targets per box, issued by ONE elected lane and completed on an mbarrier (complete_tx::bytes) -- instead
// of one 16-byte cp.async per lane per row per map.  Two stages per warp; the consumer waits on the stage's
// mbarrier phase and reads with the same LDS.128 as before.  Per 6 rows this replaces 12 LDGSTS + 6 commit +
// 6 wait + the per-row pointer arithmetic with 2 UTMALDG + 1 arrive.expect_tx + 1 try_wait.
// ------------------------------------------------------------------------------------------------
// The TMA unit wants the first byte of a box 16-byte aligned in the innermost dimension.  A warp's window
// starts at column 120*strip - 4, a multiple of 4 elements: 16 bytes for fp32, but only 8 / 4 bytes for bf16 / u8
// maps.  Those boxes start at the window's column rounded DOWN to 16 bytes and are correspondingly wider.
template <typename T>
struct TmaBox {
    static constexpr int kAlign = 16 / (int)sizeof(T);                                      // elements per 16 bytes
    static constexpr int kCols = ((32 * kVec + kAlign - kVec) + kAlign - 1) / kAlign * kAlign;  // 128 | 136 | 144
    static constexpr int kRowBytes = kCols * (int)sizeof(T);
    __device__ __forceinline__ static int first_col(int c0) { return c0 - (((c0 % kAlign) + kAlign) % kAlign); }
};
template <typename XT, typename TT, int ROWS = 6>
struct TmaRing {
    static constexpr int kBoxRows = ROWS;                  // rows per box == unroll factor of the steady-state loop that consumes it
    static constexpr int kXRowBytes = TmaBox<XT>::kRowBytes, kTRowBytes = TmaBox<TT>::kRowBytes;
    static constexpr int kXBoxBytes = kBoxRows * kXRowBytes, kTBoxBytes = kBoxRows * kTRowBytes;  // what the TMA unit delivers
    static constexpr int kXSlotBytes = (kXBoxBytes + 127) / 128 * 128, kTSlotBytes = (kTBoxBytes + 127) / 128 * 128;
    static constexpr int kStageBytes = kXSlotBytes + kTSlotBytes;  // every box starts 128-byte aligned
    static constexpr int kBarOffset = kWarpsPerBlock * 2 * kStageBytes;   // layout A: all stages, then all mbarriers (forward kernel)
    static constexpr int kSmemBytes = kBarOffset + kWarpsPerBlock * 2 * 8;
    // layout B (backward kernel): per warp {stage 0, stage 1, 2 mbarriers, pad to 128}: one base register reaches everything
    static constexpr int kWarpBarOffset = 2 * kStageBytes;
    static constexpr int kWarpBytes = 2 * kStageBytes + 128;
    static constexpr int kSmemBytesB = kWarpsPerBlock * kWarpBytes;
};
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "PIL_MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra PIL_MBAR_DONE_%=;\n"
        "bra PIL_MBAR_WAIT_%=;\n"
        "PIL_MBAR_DONE_%=:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// one 2-D box: tensor-map coordinates {column, row} (may be negative / past the end: zero-filled)
__device__ __forceinline__ void tma_load_2d(uint32_t dst_shared, const CUtensorMap* map, int col, int row, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst_shared),
                 "l"(map), "r"(col), "r"(row), "r"(bar) : "memory");
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
    // `row` already points at this thread's (clamped) column for ALIGNED, at column 0 otherwise.
    // The halo fix-up is a separate step applied when the row is CONSUMED, not when it is fetched:
    // touching the loaded registers right after the load would stall the warp on the load and
    // defeat the two-row prefetch (measured: 25% of all stall samples sat on those two selects).
    __device__ __forceinline__ float4 fix(float4 r) const {
        if constexpr (ALIGNED) {
            if (mode == 1) r.w = r.y;
            if (mode == 2) r.x = r.z;
        }
        return r;
    }
    template <typename T>
    __device__ __forceinline__ float4 load(const T* row) const {
        if constexpr (ALIGNED) {
            return ld4<T>(row);
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
// Packed fp32x2 arithmetic (sm_100 FFMA2 / FADD2 / FMUL2): one issue slot does two pixels.  Both
// kernels are issue-limited in scalar form (ncu: ~60% issue-active with the FMA pipe at 40-50%), so
// every element-wise operation of the ALIGNED path works on (slot0,slot1) / (slot2,slot3) pairs.
// Scalar constants are broadcast by the instruction itself (R.F32 operand form), no register pairs.
using f2 = float2;
__device__ __forceinline__ f2 bc(float s) { return make_float2(s, s); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }

template <int KIND>
__device__ __forceinline__ float4 act4(float4 v) {
    if constexpr (KIND == PIL_X_PROB) {
        return v;
    } else {
        const float s = (KIND == PIL_X_LOGITS_TANH) ? -2.0f * kLog2e : -kLog2e;
        const f2 a = mul2(make_float2(v.x, v.y), bc(s)), b = mul2(make_float2(v.z, v.w), bc(s));
        const f2 da = add2(make_float2(ex2_approx(a.x), ex2_approx(a.y)), bc(1.0f));
        const f2 db = add2(make_float2(ex2_approx(b.x), ex2_approx(b.y)), bc(1.0f));
        return make_float4(rcp_approx(da.x), rcp_approx(da.y), rcp_approx(db.x), rcp_approx(db.y));
    }
}

// ------------------------------------------------------------------------------------------------
// geometry shared by host and device
// ------------------------------------------------------------------------------------------------
// The B*H image rows of a shard are cut into `groups` equal ranges (to +-1 row; a range may straddle
// an image boundary and is then processed as two segments).  Group g is processed by `strips` warps,
// one per 120-column strip, so the warps of a group sweep full rows together (DRAM page locality).
// groups*strips is sized to the number of warps the grid keeps resident, so every SM gets the same
// number of blocks and all warps finish together (the kernels are not purely HBM-bound, so an SM with
// one block more than its neighbour would otherwise be the critical path).
struct Geo {
    int B, H, W;
    int strips;            // warps per row band = ceil(W / 120)
    long long groups;      // row ranges
    long long total_rows;  // B * H
    long long tasks;       // (groups + tail_groups) * strips warp-tasks
    // Optional tail phase (backward, dynamic claiming): rows [0, tail_rows) are cut into shorter ranges of tail_range rows
    // -- task indices [0, tail_groups * strips) -- and only rows [tail_rows, total_rows) into the `groups` long ones.  The
    // shard is walked from its END, so the short ranges are the LAST to be claimed: the grid finishes within a short
    // range's length of each other instead of a long one's.  tail_rows == 0: no tail phase.
    long long tail_rows, tail_groups;
    int tail_range;
};

// peer-memory exchange descriptor handed to the kernels (see the exchange helpers below)
constexpr int kSlotBytes = 128;  // 16 words of {32-bit payload half, 32-bit step tag}
constexpr int kXchgStatusOffset = 2 * 2 * PIL_MAX_RANKS * kSlotBytes;  // int status word after the slots
constexpr int kXchgEpochOffset = kXchgStatusOffset + 8;                 // device-resident step counter (PIL_XCHG_DEVICE_EPOCH)
struct XchgDev {
    int rank, world;             // world == 0: exchange disabled
    int parity;
    int defer;                   // PIL_XCHG_DEFER_FINALIZE: the backward only pushes phase 1
    int device_epoch;            // PIL_XCHG_DEVICE_EPOCH: tag and parity come from the counter in the own mailbox
    unsigned long long want;     // flag value of this step (epoch + 1)
    unsigned long long timeout_ns;
    unsigned char* box[PIL_MAX_RANKS];
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
    // dynamic work distribution: after its first, statically assigned range a warp claims further (range, strip) tasks
    // from this counter (null: one static task per warp); tasks [0, first_dynamic) are the static ones
    unsigned int* task_counter;
    long long first_dynamic;
    XchgDev X;               // MOMENTS, world > 0: the last block pushes the 16 moment sums to every rank (phases 0 and 1)
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
    int reverse;  // walk the shard back to front (L2 reuse after the pointwise forward)
    int skip_unit_upstream;  // pil_backward_if_scaled: nothing to do when *upstream == 1 (grad already holds that gradient)
    // accumulate mode (pil_backward_accumulate): the stencil sums the pointwise forward left out
    int accumulate;
    double* partials;
    unsigned int* ticket;
    double* stencil_sums;  // out: {0,0,0,0, sum r^2, (eps/8) sum(dx^2+dy^2), 0, 0} of this shard
    float* loss_out;       // optional: finalize(gsums + stencil_sums) as if this shard were the batch
    double* total_sums;    // optional: gsums + stencil_sums (may alias gsums: written by the last block only)
    // dynamic work distribution (accumulate mode): after its first, statically assigned range a warp
    // claims further (range, strip) tasks from this counter; tasks [0, first_dynamic) are the static ones
    unsigned int* task_counter;
    long long first_dynamic;
    XchgDev X;             // world > 0: global sums come from the mailbox (phase 0); the last block exchanges
                           // the stencil sums (phase 1) and finalises the GLOBAL loss
};

// The loss from the sums, src/loss.py:134-160, in two steps so that the four quotients can be formed by four threads
// side by side (the backward's last block: serial time at the end of the step): term 0 dice_loss, 1 bce, 2 rd, 3 pf.
__device__ __forceinline__ double finalize_term(const double* s, double n, const PilParams& p, int k) {
    if (k == 0) return 1.0 - (2.0 * s[0] + p.smooth) / (s[1] + s[2] + p.smooth);
    return s[2 + k] / n;
}
__device__ __forceinline__ void finalize_from_terms(const double* term, const double* s, const PilParams& p, float* out) {
    const double dice_loss = term[0], bce = term[1], rd = term[2], pf = term[3];
    double total = p.dice_weight * dice_loss + p.bce_weight * bce;
    if (p.pde_weight > 0.0) total += p.pde_weight * rd;
    if (p.phase_field_weight > 0.0) total += p.phase_field_weight * pf;
    if (s[6] > 0.0) total = __longlong_as_double(0x7ff8000000000000LL);  // see finalize_device
    out[0] = (float)total;
    out[1] = (float)dice_loss;
    out[2] = (float)bce;
    out[3] = (float)rd;
    out[4] = (float)pf;
    out[5] = (float)s[6];
    out[6] = 0.f;
    out[7] = 0.f;
}
__device__ __forceinline__ void finalize_device(const double* s, double n, const PilParams& p, float* out) {
    // src/loss.py:134-160
    const double I = s[0], P = s[1], T = s[2];
    const double dice_loss = 1.0 - (2.0 * I + p.smooth) / (P + T + p.smooth);
    const double bce = s[3] / n, rd = s[4] / n, pf = s[5] / n;
    double total = p.dice_weight * dice_loss + p.bce_weight * bce;
    if (p.pde_weight > 0.0) total += p.pde_weight * rd;
    if (p.phase_field_weight > 0.0) total += p.phase_field_weight * pf;
    // nn.BCELoss refuses inputs outside [0,1] (the reference's step dies there, src/loss.py:141).  Stream-ordered
    // code cannot raise, so the total is poisoned instead: a NaN loss is the loudest failure that needs no host
    // sync.  The count stays in slot 5; the Python module can raise on it (strict_inputs).
    if (s[6] > 0.0) total = __longlong_as_double(0x7ff8000000000000LL);
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
// Programmatic dependent launch (PDL): every hot kernel is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so its blocks may become resident while the
// previous kernel of the stream is still draining (tail blocks, the last-block reduction, the launch
// latency itself).  pdl_wait() blocks until the previous kernel has completed and its writes are
// visible; nothing written by an earlier kernel may be touched before it.  pdl_launch_dependents()
// lets the NEXT kernel's blocks be scheduled as soon as SM resources free up.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Development instrumentation (-DPIL_TIMELINE, tools/timeline.py): per-block globaltimer stamps.
#ifdef PIL_TIMELINE
__device__ unsigned long long* g_timeline = nullptr;  // [kernel 0/1][4096 blocks][8]
__device__ __forceinline__ void tl_stamp(int kernel, int slot) {
    if (g_timeline != nullptr && threadIdx.x == 0 && blockIdx.x < 4096) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        unsigned long long* p = g_timeline + ((size_t)kernel * 4096 + blockIdx.x) * 8;
        p[slot] = t;
        if (slot == 0) {
            unsigned int smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            p[7] = smid;
        }
    }
}
#define TL_STAMP(k, s) tl_stamp(k, s)
#else
#define TL_STAMP(k, s)
#endif

// ------------------------------------------------------------------------------------------------
// Peer-memory exchange of the sums vectors (data parallel, one process per GPU; include/pil.h
// PilExchange).  Every rank owns a small mailbox in its own HBM that all peers have mapped (CUDA IPC
// over NVLink/NVSwitch).  The last block of a kernel PUSHES its shard's 8 doubles into slot
// [phase][epoch parity][my rank] of every rank's mailbox (16 remote 8-byte stores, each carrying its
// own step tag); the consumer POLLS its own, local copy and adds the R vectors in rank order, so every
// rank forms bit-identical global sums.  No NCCL call, no extra launch, no host involvement.
//   phase 0: pointwise sums  (pushed by the pointwise forward, consumed by every block of the backward)
//   phase 1: stencil sums    (pushed by the backward's last block, consumed by that same block)
// Slot reuse: a slot of parity q is rewritten two steps later; by then every rank has passed the
// phase-0 wait of the step in between, which is stream-ordered after its reads of the old value.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned char* xchg_slot(unsigned char* base, int phase, int parity, int src) {
    return base + (size_t)(((phase * 2 + parity) * PIL_MAX_RANKS + src) * kSlotBytes);
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Wire format ("LL" style, no fences): a slot is 16 words of 8 bytes, word w = {32-bit half w of the
// 8 doubles, 32-bit step tag}.  A naturally aligned 8-byte store is single-copy atomic, so a reader that
// sees the tag of this step in a word also sees that word's payload -- no release/acquire pair, no
// system-scope fence (which would cost an NVLink round trip in the kernel's tail).
constexpr int kSlotWords = 2 * PIL_NSUMS;
static_assert(kSlotWords * 8 == kSlotBytes, "slot layout");

// Step tag and slot parity.  Host epoch: both come with the descriptor.  Device epoch (PIL_XCHG_DEVICE_EPOCH):
// they derive from a counter in the rank's OWN mailbox that the backward's last block increments when the
// step is complete -- the kernel arguments are then identical from step to step, so the pair of launches
// can be captured once in a CUDA graph and replayed.  Every rank counts the same completed steps.
__device__ __forceinline__ void xchg_epoch(const XchgDev& X, unsigned long long& want, int& parity) {
    want = X.want;
    parity = X.parity;
    if (X.device_epoch) {
        unsigned long long e;
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(e) : "l"(X.box[X.rank] + kXchgEpochOffset) : "memory");
        want = (e % 0xfffffffeull) + 1ull;
        parity = (int)(e & 1ull);
    }
}
// the step is complete on this rank: called by ONE thread of the backward's last block, after its last wait
__device__ __forceinline__ void xchg_advance_epoch(const XchgDev& X) {
    if (X.device_epoch) {
        unsigned long long* p = reinterpret_cast<unsigned long long*>(X.box[X.rank] + kXchgEpochOffset);
        *p = *p + 1ull;
    }
}
// called by ALL threads of ONE block (blockDim >= 16*world): v (shared memory) -> every rank's mailbox
__device__ __forceinline__ void xchg_push(const XchgDev& X, int phase, const double* v) {
    const int i = (int)threadIdx.x;
    if (i < kSlotWords * X.world) {
        unsigned long long want;
        int parity;
        xchg_epoch(X, want, parity);
        const int r = i / kSlotWords, w = i % kSlotWords;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(v[w >> 1]);
        const unsigned long long half = (w & 1) ? (bits >> 32) : (bits & 0xffffffffull);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(xchg_slot(X.box[r], phase, parity, X.rank)) + w;
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(half | (want << 32)) : "memory");
    }
}
// called by ALL threads of ONE block (blockDim >= 16*world; contains __syncthreads): waits for the R
// vectors of `phase` in the LOCAL mailbox and adds them in rank order into out[0..7] (shared memory).
// On timeout the sums are NaN, the mailbox status word is set and the function returns false (block-uniform):
// the backward then writes a ZERO gradient, so a rank that lost its peers cannot poison the weights.
static __device__ __forceinline__ bool xchg_wait_sum(const XchgDev& X, int phase, double* out) {
    __shared__ unsigned int s_half[PIL_MAX_RANKS * kSlotWords];
    __shared__ int s_bad;
    const int i = (int)threadIdx.x;
    if (i == 0) s_bad = 0;
    __syncthreads();
    if (i < kSlotWords * X.world) {
        unsigned long long want;
        int parity;
        xchg_epoch(X, want, parity);
        const int r = i / kSlotWords, w = i % kSlotWords;
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(xchg_slot(X.box[X.rank], phase, parity, r)) + w;
        const unsigned long long t0 = globaltimer_ns();
        unsigned long long word;
        for (;;) {
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(word) : "l"(src) : "memory");
            if ((word >> 32) == want) break;
            if (globaltimer_ns() - t0 > X.timeout_ns) {
                s_bad = 1;
                break;
            }
            __nanosleep(32);
        }
        s_half[i] = (unsigned int)(word & 0xffffffffull);
    }
    __syncthreads();
    if (i < PIL_NSUMS) {
        double v = 0.0;
        for (int r = 0; r < X.world; ++r) {
            const unsigned long long lo = s_half[r * kSlotWords + 2 * i], hi = s_half[r * kSlotWords + 2 * i + 1];
            v += __longlong_as_double((long long)(lo | (hi << 32)));
        }
        out[i] = s_bad ? __longlong_as_double(0x7ff8000000000000LL) : v;
    }
    if (i == 0 && s_bad) *reinterpret_cast<volatile int*>(X.box[X.rank] + kXchgStatusOffset) = 1;
    __syncthreads();
    return s_bad == 0;
}

// ------------------------------------------------------------------------------------------------
// Deterministic two-level reduction of the per-thread accumulators:
//   warp shuffles -> per-block doubles in `partials` -> the LAST block to finish (ticket) adds all
//   per-block partials in a fixed order, so the result does not depend on block scheduling.
// Returns true in the last block only; there thread 0 holds the totals in out[].
// accumulator layout: 0 I, 1 P, 2 T, 3 bce (log2 units, un-negated), 4 r^2, 5 dx^2+dy^2, 6 (uv)^2, 7 #invalid
//
// No fence on the way.  This is serial time at the very end of a kernel (the next kernel of the stream waits for it), and
// a __threadfence() there first waits for every gradient store the thread still has in flight (1.3-1.7 us in the backward,
// tools/timeline.py).  Instead every partial travels self-validating, like the peer mailbox words above: a double is two
// 8-byte words {32 payload bits, 32-bit tag of THIS launch}, a naturally aligned 8-byte store is single-copy atomic, and
// the last block re-reads a word until it carries the tag (only stores that left just before the last ticket can still be
// in flight).  The tag comes with the ticket: the 64-bit word {reduction epoch, blocks arrived} at header words 2-3 is
// incremented by every block, so each block learns the epoch from its own atomic, consistently (no block of this launch
// can see the value the last block leaves behind: {epoch + 1, 0}).  Stale slots carry the tags of earlier epochs or the
// zero of pil_workspace_init, never the current one.
// Workspace header (unsigned int, relative to `ticket`): [0] ticket of the kernels that keep their own scheme,
// [1] task counter, [2..3] {blocks arrived, reduction epoch}; the WHOLE workspace is zero before first use.
// ------------------------------------------------------------------------------------------------
constexpr int kWsRedWord = 2;
constexpr size_t kPartialBytes = 16;  // bytes one double occupies in the partials area
__device__ __forceinline__ ulonglong2 partial_load(const ulonglong2* p) {
    ulonglong2 w;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w.x), "=l"(w.y) : "l"(p) : "memory");
    return w;
}
// second level: the block's N doubles (thread k < N holds component k in `blk`) -> `partials` -> the LAST block to
// finish adds all blocks' partials in a fixed order.  Every warp of the block has finished its work when this is called
// (the callers synchronise before forming `blk`).  Returns true in the last block only; there thread 0 holds the
// totals in out[], and the header is already set for the next launch.
template <int THREADS, int N, int TLK = -1>
__device__ __forceinline__ bool blocks_to_last(double blk, double* partials, unsigned int* ticket, double* out) {
    constexpr int kWarps = THREADS / 32;
    static_assert((N & (N - 1)) == 0 && N <= 32 && THREADS % N == 0, "component layout");
    __shared__ double s_red[kWarps * N];
    __shared__ double s_tot[N];
    __shared__ unsigned long long s_old;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long* red = reinterpret_cast<unsigned long long*>(ticket + kWsRedWord);
    if constexpr (TLK >= 0) TL_STAMP(TLK, 4);  // every warp of the block has arrived
    if (threadIdx.x == 0) s_old = atomicAdd(red, 1ull);
    __syncthreads();
    const unsigned long long old = s_old;
    const unsigned int epoch = (unsigned int)(old >> 32);
    const unsigned long long tag = (unsigned long long)(epoch % 0xfffffffeu + 1u) << 32;
    if (threadIdx.x < N) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(blk);
        ulonglong2* dst = reinterpret_cast<ulonglong2*>(partials) + (long long)blockIdx.x * N + threadIdx.x;
        asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"((bits & 0xffffffffull) | tag), "l"((bits >> 32) | tag)
                     : "memory");
    }
    if constexpr (TLK >= 0) TL_STAMP(TLK, 5);  // ticket taken, partial sent
    if ((unsigned int)old != gridDim.x - 1) return false;
    if (threadIdx.x == 0) *red = (unsigned long long)(epoch + 1u) << 32;  // every block has arrived: ready for the next launch
    constexpr int kGroups = THREADS / N;
    {
        // kGroups block-groups x N components: 16-byte L2 loads, kIlp in flight per thread (the tail batch is
        // predicated, not serialised), fixed order -> bit-reproducible; 1-2 L2 round trips for the usual grids
        constexpr int kIlp = N >= 16 ? 14 : (N >= 8 ? 10 : 8);
        const int c = threadIdx.x % N, j = threadIdx.x / N;
        const long long nb = gridDim.x;
        const ulonglong2* src = reinterpret_cast<const ulonglong2*>(partials) + c;
        double v = 0.0;
        for (long long b0 = j; b0 < nb; b0 += (long long)kIlp * kGroups) {
            ulonglong2 w[kIlp];
#pragma unroll
            for (int q = 0; q < kIlp; ++q) {
                const long long b = b0 + (long long)q * kGroups;
                w[q] = (b < nb) ? partial_load(src + b * N) : make_ulonglong2(tag, tag);
            }
#pragma unroll
            for (int q = 0; q < kIlp; ++q) {
                while ((((w[q].x ^ tag) | (w[q].y ^ tag)) >> 32) != 0ull)  // still in flight
                    w[q] = partial_load(src + (b0 + (long long)q * kGroups) * N);
                v += __longlong_as_double((long long)((w[q].x & 0xffffffffull) | (w[q].y << 32)));
            }
        }
        // threads with the same lane % N hold partial sums of the same component: butterfly over the other
        // lane bits (fixed pattern -> bit-reproducible), then one value per warp and component through shared memory
#pragma unroll
        for (int o = 16; o >= N; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane < N) s_red[warp * N + lane] = v;
    }
    __syncthreads();
    if (threadIdx.x < N) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) v += s_red[w * N + threadIdx.x];
        s_tot[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) out[k] = s_tot[k];
    }
    if constexpr (TLK >= 0) TL_STAMP(TLK, 4);  // last block only (overwrites its arrival stamp): every block's partials added
    return true;
}

// first level: per-thread accumulators -> warp shuffles -> the block's doubles; then blocks_to_last
template <int THREADS, typename AccT, int N = PIL_NSUMS, int TLK = -1>
__device__ __forceinline__ bool reduce_to_last_block(const AccT* acc, double* partials, unsigned int* ticket, double* out) {
    constexpr int kWarps = THREADS / 32;
    __shared__ double s_part[kWarps][N];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        AccT v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_part[warp][k] = (double)v;
    }
    __syncthreads();
    double blk = 0.0;
    if (threadIdx.x < N) {
#pragma unroll
        for (int w = 0; w < kWarps; ++w) blk += s_part[w][threadIdx.x];
    }
    return blocks_to_last<THREADS, N, TLK>(blk, partials, ticket, out);
}

// raw accumulator totals -> the sums vector of include/pil.h
__device__ __forceinline__ void sums_from_raw(const double* raw, double eps, double n_pixels, double* s) {
    s[0] = raw[0];
    s[1] = raw[1];
    s[2] = raw[2];
    s[3] = -(double)kLn2 * raw[3];                 // back from log2 units, BCE sign
    s[4] = raw[4];
    s[5] = (eps / 8.0) * raw[5] + raw[6] / eps;    // (eps/2)*(dx/2)^2 ... + (uv)^2/eps
    s[6] = raw[7];
    s[7] = n_pixels;
}

// argument blocks of the pointwise forward kernels (pil_point.cu)
constexpr int kPointThreads = 256;
constexpr int kPointUnroll = 4;  // float4 pairs in flight per thread
struct PointArgs {
    const void* x;
    const void* t;
    long long keep_from4;    // fp32 maps: float4 index from which the loads ask L2 to keep the lines (evict_last)
    long long n;             // pixels in the shard
    double* partials;
    unsigned int* ticket;
    double* sums;
    PilParams p;
    XchgDev X;               // world > 0: the last block pushes the shard's sums to every rank (phase 0)
};
struct PointMetricsArgs {
    const void* x;
    const void* t;
    long long n;    // pixels in the shard
    long long hw;   // pixels per image
    double* partials;
    unsigned int* ticket;
    double* sums;
    PilParams p;
    double* image_counts;  // [B][4], zero on entry
    float threshold;
    int l2_stream;         // fp32 maps: loads carry an L2 evict_first hint (see pil_point_kernel)
    XchgDev X;
};

// ------------------------------------------------------------------------------------------------
// host side shared by the translation units
// ------------------------------------------------------------------------------------------------
// Process-wide knobs and counters (pil_set_tuning, pil_set_l2_keep_mb, pil_last_launch_info).  Atomics: the
// library is re-entrant across host threads and streams.
struct HostState {
    std::atomic<long long> kernels_launched{0};
    std::atomic<int> tune_fwd_rps{0}, tune_bwd_rps{0};
    std::atomic<long long> l2_keep_mb{-1};  // < 0: PIL_L2_KEEP_MB or the default
    std::atomic<int> bwd_stage{-1};         // pil_set_bwd_staging: 0 cp.async, 1 TMA, < 0 default / PIL_BWD_STAGE
};
HostState& host_state();  // pil_api.cu
extern thread_local PilLaunchInfo t_info;  // how the last launch on this host thread was tiled

// Per-DEVICE caches: one process may drive several GPUs (and several host threads), so nothing that depends
// on the device is cached process-wide.
constexpr int kMaxDevices = 64;
inline int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
    return dev;
}
int sm_count();  // SMs of the current device (pil_api.cu)
// resident blocks per SM of `kernel` on the current device; `cache` is the caller's static [kMaxDevices] array
template <typename K>
inline int blocks_per_sm_cached(K kernel, int threads, int smem_bytes, std::atomic<int>* cache, bool max_smem_carveout) {
    std::atomic<int>& slot = cache[current_device()];
    int n = slot.load(std::memory_order_relaxed);
    if (n == 0) {
        if (max_smem_carveout) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem_bytes) != cudaSuccess || n < 1) n = 1;
        slot.store(n, std::memory_order_relaxed);
    }
    return n;
}

constexpr long long kMaxPointBlocks = 2048;  // grid cap of the pointwise forward (workspace sizing)

// Size the row-range partition: one range-group per `strips` resident warps (waves = 1), so that
// every SM holds the same number of blocks; never fewer than kMinRows rows per range.
constexpr int kMinRows = 8;
inline Geo make_geo(int64_t B, int64_t H, int64_t W, int resident_blocks, int forced_rows, int waves) {
    Geo g;
    g.B = (int)B;
    g.H = (int)H;
    g.W = (int)W;
    g.strips = (int)((W + kStripCols - 1) / kStripCols);
    g.total_rows = (long long)B * H;
    g.tail_rows = g.tail_groups = 0;
    g.tail_range = 0;
    long long groups;
    if (forced_rows > 0) {
        groups = (g.total_rows + forced_rows - 1) / forced_rows;
    } else {
        const long long warps = (long long)resident_blocks * kWarpsPerBlock * (waves > 0 ? waves : 1);
        groups = warps / g.strips;
        const long long cap = (g.total_rows + kMinRows - 1) / kMinRows;
        if (groups > cap) groups = cap;
    }
    if (groups < 1) groups = 1;
    if (groups > g.total_rows) groups = g.total_rows;
    g.groups = groups;
    g.tasks = groups * g.strips;
    return g;
}


struct LaunchOut {
    int blocks = 0, rows = 0;
    int status = PIL_OK;              // PIL_ERR_WORKSPACE when the partials do not fit
    size_t partials_avail = ~(size_t)0;  // in: bytes of the per-block partials area (backward, accumulate mode)
    int tma = 0;                         // out: rows staged by TMA boxes
    unsigned int* task_counter = nullptr;  // in (forward): the workspace's task counter, for dynamic claiming
};
// Rows per range the kernels like best (measured on B200, 64x1024^2 .. 128x2048^2): long enough to
// amortise the warm-up rows and the pipeline fill of a segment, short enough that the hardware block
// scheduler can still even out SM-to-SM speed differences with a few waves.
constexpr int kTargetRowsFwd = 256, kTargetRowsBwd = 200;
inline int tuning_waves(bool bwd, int64_t B, int64_t H, int64_t W, int resident_blocks) {
    static int forced[2] = {-1, -1};
    if (forced[bwd] < 0) {
        const char* e = getenv(bwd ? "PIL_WAVES_BWD" : "PIL_WAVES_FWD");
        forced[bwd] = (e && atoi(e) > 0) ? atoi(e) : 0;
    }
    if (forced[bwd] > 0) return forced[bwd];
    const long long strips = (W + kStripCols - 1) / kStripCols;
    const long long groups1 = (long long)resident_blocks * kWarpsPerBlock / strips;  // ranges in one wave
    if (groups1 < 1) return 1;
    const double rows1 = (double)(B * H) / (double)groups1;
    const int target = bwd ? kTargetRowsBwd : kTargetRowsFwd;
    int w = (int)(rows1 / target + 0.5);
    return w < 1 ? 1 : w;
}

// Launch with programmatic stream serialization (see pdl_wait above).  PIL_PDL=0 turns it off.
inline bool use_pdl() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PIL_PDL");
        v = (e && atoi(e) == 0) ? 0 : 1;
    }
    return v == 1;
}
template <typename K, typename... A>
inline cudaError_t launch_pdl(K kernel, int blocks, int threads, int smem, cudaStream_t s, const A&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = use_pdl() ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

template <typename T>
constexpr int dtype_code() {
    return std::is_same<T, float>::value ? PIL_F32 : (std::is_same<T, __nv_bfloat16>::value ? PIL_BF16 : PIL_U8);
}

// 2-D tensor map of a (rows x cols) row-major map with a (box_rows x box_cols) box; false when the layout does not
// qualify (TMA needs a 16-byte aligned base and row pitch) or the driver entry point is unavailable.  pil_api.cu
bool make_tensor_map_2d(CUtensorMap* out, const void* base, int dtype, long long rows, long long cols, int box_rows, int box_cols);

// PilExchange (include/pil.h) -> the descriptor the kernels take; PIL_ERR_EXCHANGE for a bad one.  pil_api.cu
int make_xchg(const PilExchange* ex, XchgDev* X);

// per-kind launchers, one translation unit each in release builds (pil_fwd.cu, pil_point.cu, pil_bwd.cu)
#define PIL_DECL_KIND(K)                                                                                                        \
    cudaError_t launch_fwd_k##K(int x_dtype, int t_dtype, FwdArgs& a, int64_t B, int64_t H, int64_t W, bool aligned,            \
                                size_t avail, cudaStream_t s, LaunchOut* out, bool moments);                                    \
    cudaError_t launch_bwd_k##K(int x_dtype, int t_dtype, BwdArgs& a, int64_t B, int64_t H, int64_t W, bool aligned,            \
                                cudaStream_t s, LaunchOut* out);                                                                \
    cudaError_t launch_point_k##K(int x_dtype, int t_dtype, const PointArgs& a, bool aligned, cudaStream_t s, int* b);          \
    cudaError_t launch_point_metrics_k##K(int x_dtype, int t_dtype, const PointMetricsArgs& a, bool aligned, cudaStream_t s,   \
                                          int* b);
PIL_DECL_KIND(0)
PIL_DECL_KIND(1)
PIL_DECL_KIND(2)
#undef PIL_DECL_KIND

}  // namespace pil
#endif  // PIL_COMMON_CUH_
