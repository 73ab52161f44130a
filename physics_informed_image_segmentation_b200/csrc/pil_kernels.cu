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
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "pil.h"

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
    long long tasks;       // groups * strips warp-tasks
};

// peer-memory exchange descriptor handed to the kernels (see the exchange helpers below)
constexpr int kSlotBytes = 128;  // 16 words of {32-bit payload half, 32-bit step tag}
constexpr int kXchgStatusOffset = 2 * 2 * PIL_MAX_RANKS * kSlotBytes;  // int status word after the slots
struct XchgDev {
    int rank, world;             // world == 0: exchange disabled
    int parity;
    int defer;                   // PIL_XCHG_DEFER_FINALIZE: the backward only pushes phase 1
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

// called by ALL threads of ONE block (blockDim >= 16*world): v (shared memory) -> every rank's mailbox
__device__ __forceinline__ void xchg_push(const XchgDev& X, int phase, const double* v) {
    const int i = (int)threadIdx.x;
    if (i < kSlotWords * X.world) {
        const int r = i / kSlotWords, w = i % kSlotWords;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(v[w >> 1]);
        const unsigned long long half = (w & 1) ? (bits >> 32) : (bits & 0xffffffffull);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(xchg_slot(X.box[r], phase, X.parity, X.rank)) + w;
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(half | (X.want << 32)) : "memory");
    }
}
// called by ALL threads of ONE block (blockDim >= 16*world; contains __syncthreads): waits for the R
// vectors of `phase` in the LOCAL mailbox and adds them in rank order into out[0..7] (shared memory).
// On timeout the sums are NaN and the mailbox status word is set.
__device__ __noinline__ void xchg_wait_sum(const XchgDev& X, int phase, double* out) {
    __shared__ unsigned int s_half[PIL_MAX_RANKS * kSlotWords];
    __shared__ int s_bad;
    const int i = (int)threadIdx.x;
    if (i == 0) s_bad = 0;
    __syncthreads();
    if (i < kSlotWords * X.world) {
        const int r = i / kSlotWords, w = i % kSlotWords;
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(xchg_slot(X.box[X.rank], phase, X.parity, r)) + w;
        const unsigned long long t0 = globaltimer_ns();
        unsigned long long word;
        for (;;) {
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(word) : "l"(src) : "memory");
            if ((word >> 32) == X.want) break;
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
}

// ------------------------------------------------------------------------------------------------
// Deterministic two-level reduction of the 8 per-thread accumulators:
//   warp shuffles -> per-block doubles in `partials` -> the LAST block to finish (ticket) adds all
//   per-block partials in a fixed order, so the result does not depend on block scheduling.
// Returns true in the last block only; there thread 0 holds the totals in out[].  The caller resets
// *ticket to 0 when it is done (so the workspace is reusable by the next launch on the stream).
// accumulator layout: 0 I, 1 P, 2 T, 3 bce (log2 units, un-negated), 4 r^2, 5 dx^2+dy^2, 6 (uv)^2, 7 #invalid
// ------------------------------------------------------------------------------------------------
template <int THREADS, typename AccT, int N = PIL_NSUMS>
__device__ __forceinline__ bool reduce_to_last_block(const AccT* acc, double* partials, unsigned int* ticket, double* out) {
    constexpr int kWarps = THREADS / 32;
    constexpr int NP = N / 2;  // component pairs (16-byte loads)
    static_assert(N % 2 == 0 && THREADS % NP == 0, "component layout");
    __shared__ double s_part[kWarps][N];
    __shared__ double s_red[2 * THREADS];
    __shared__ double s_tot[N];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        AccT v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_part[warp][k] = (double)v;
    }
    __syncthreads();
    if (threadIdx.x < N) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) v += s_part[w][threadIdx.x];
        partials[(long long)blockIdx.x * N + threadIdx.x] = v;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(ticket, 1u);
        s_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    constexpr int kGroups = THREADS / NP;
    {
        // kGroups block-groups x NP component PAIRS: 16-byte L2 loads, 8 in flight per thread (the tail
        // batch is predicated, not serialised), fixed order -> bit-reproducible.  This is serial time
        // at the very end of the kernel, so it is kept to 2-3 L2 round trips.
        constexpr int kIlp = 8;
        const int c2 = threadIdx.x % NP, j = threadIdx.x / NP;
        const long long nb = gridDim.x;
        double vx = 0.0, vy = 0.0;
        for (long long blk = j; blk < nb; blk += (long long)kIlp * kGroups) {
            double2 w[kIlp];
#pragma unroll
            for (int q = 0; q < kIlp; ++q) {
                const long long b = blk + (long long)q * kGroups;
                w[q] = (b < nb) ? __ldcg(reinterpret_cast<const double2*>(partials + b * N) + c2) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int q = 0; q < kIlp; ++q) {
                vx += w[q].x;
                vy += w[q].y;
            }
        }
        s_red[2 * threadIdx.x] = vx;       // s_red viewed as [kGroups][N]
        s_red[2 * threadIdx.x + 1] = vy;
    }
    __syncthreads();
    if (threadIdx.x < N) {
        double v = 0.0;
        for (int j = 0; j < kGroups; ++j) v += s_red[j * N + threadIdx.x];
        s_tot[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) out[k] = s_tot[k];
    }
    return true;
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

// ------------------------------------------------------------------------------------------------
// K1: fused forward
// ------------------------------------------------------------------------------------------------
// MOMENTS (pil_forward_moments, the parameter-sweep entry): instead of sum r^2 for ONE (D, a) the row
// accumulates the parameter-independent second moments of {lap, g = u(1-u), h = g*u}; r = D*lap + h - a*g,
// so sum r^2 for ANY (D, a) is a quadratic form in them (include/pil.h).
template <int KIND, bool ALIGNED, bool MOMENTS = false>
struct FwdRow {
    // accumulators: 0 I, 1 P, 2 T, 3 sum(t*max(lg2 u,c) + (1-t)*max(lg2(1-u),c)), 4 sum r^2 (MOMENTS: sum lap^2),
    //               5 sum dx^2+dy^2 (raw differences), 6 sum (u(1-u))^2, 7 #invalid
    //     MOMENTS:  8 sum lap*h, 9 sum lap*g, 10 sum h^2, 11 sum h*g, 12..15 unused
    float acc[MOMENTS ? 16 : 8];
    f2 ma[4];  // MOMENTS, packed path: 8..11
    float D, a;
    float m[4];  // !ALIGNED: 1 for slots that are real output pixels of this thread

    // ALIGNED: accumulators are kept per slot parity (a0 = even slots, a1 = odd slots) as pairs
    f2 pa[7];  // 0 I, 1 P, 2 T, 3 bce(log2 units), 4 r^2, 5 dx^2+dy^2, 6 (uv)^2
    __device__ __forceinline__ void init_packed() {
#pragma unroll
        for (int k = 0; k < 7; ++k) pa[k] = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) ma[k] = make_float2(0.f, 0.f);
    }
    __device__ __forceinline__ void fold_packed() {
#pragma unroll
        for (int k = 0; k < 7; ++k) acc[k] = pa[k].x + pa[k].y;
        if constexpr (MOMENTS) {
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[8 + k] = ma[k].x + ma[k].y;
        }
    }
    // ---- packed forward, split in two so that the BCE logarithms can reuse the sigmoid's internals ----
    // point2(): activation + every per-pixel term of a pixel pair (I, P, T, BCE, double well).
    //   logits kinds: den = 1 + 2^xs (xs = -z*log2e), u = 1/den, and with ONE more MUFU, L = lg2(den):
    //       lg2(u)   = -L
    //       lg2(1-u) = xs - L        (1-u = 2^xs / den)
    //   i.e. 3 MUFU per pixel instead of 4 (the forward kernel is bound by the XU pipe: ncu 71% with
    //   mio/math-pipe throttle as the top stalls).  The reference evaluates log(1 - fl(u)); the two agree
    //   to the reference's own fp32 rounding of u, except where fl(u) == 1 exactly: there the reference's
    //   log(0) is clamped to -100 (nn.BCELoss), which is reproduced by the v == 0 select below.
    //   probability kind: u is given, both logarithms are taken directly (2 MUFU).
    __device__ __forceinline__ f2 point2(f2 x, f2 t, bool count) {
        f2 u, lu, lv, v;
        if constexpr (KIND == PIL_X_PROB) {
            u = x;
            v = sub2(bc(1.0f), u);
            lu = make_float2(fmaxf(lg2_approx(u.x), kLogClampLog2), fmaxf(lg2_approx(u.y), kLogClampLog2));
            lv = make_float2(fmaxf(lg2_approx(v.x), kLogClampLog2), fmaxf(lg2_approx(v.y), kLogClampLog2));
        } else {
            const float sc = (KIND == PIL_X_LOGITS_TANH) ? -2.0f * kLog2e : -kLog2e;
            const f2 xs = mul2(x, bc(sc));
            const f2 den = add2(make_float2(ex2_approx(xs.x), ex2_approx(xs.y)), bc(1.0f));
            u = make_float2(rcp_approx(den.x), rcp_approx(den.y));
            const f2 L = make_float2(lg2_approx(den.x), lg2_approx(den.y));
            v = sub2(bc(1.0f), u);
            lu = make_float2(fmaxf(-L.x, kLogClampLog2), fmaxf(-L.y, kLogClampLog2));
            // u*(1-u) == 0 exactly <=> fl(u) is 0 (2^xs overflowed, L = inf) or 1 (2^xs absorbed):
            // there lg2(1-u) is 0 resp. the clamp -- i.e. clamp*u -- instead of xs - L.
            const f2 d = sub2(xs, L), sp = mul2(bc(kLogClampLog2), u), uvz = mul2(u, v);
            lv = make_float2(uvz.x == 0.0f ? sp.x : d.x, uvz.y == 0.0f ? sp.y : d.y);
        }
        if (count) {
            const f2 uv = mul2(u, v);
            pa[0] = fma2(u, t, pa[0]);
            pa[1] = add2(pa[1], u);
            pa[2] = add2(pa[2], t);
            pa[3] = add2(pa[3], fma2(t, sub2(lu, lv), lv));  // t*lu + (1-t)*lv
            pa[6] = fma2(uv, uv, pa[6]);
            if constexpr (KIND == PIL_X_PROB) {
                acc[7] += (u.x >= 0.0f && u.x <= 1.0f) ? 0.0f : 1.0f;
                acc[7] += (u.y >= 0.0f && u.y <= 1.0f) ? 0.0f : 1.0f;
            }
        }
        return u;
    }
    __device__ __forceinline__ float4 point4(const float4& x, const float4& t, bool count) {
        const f2 a2 = point2(make_float2(x.x, x.y), make_float2(t.x, t.y), count);
        const f2 b2 = point2(make_float2(x.z, x.w), make_float2(t.z, t.w), count);
        return make_float4(a2.x, a2.y, b2.x, b2.y);
    }
    // stencil2(): residual and gradient-energy terms of a pixel pair of the centre row.
    //   r = D*(s4 - 4u) + u(1-u)(u-a) as a polynomial in u: u*(u*((1+a) - u) - a - 4D) + D*s4
    float a1, c0;  // 1+a, -a-4D
    __device__ __forceinline__ void stencil2(f2 u, f2 lf, f2 rt, f2 m, f2 p) {
        const f2 s4 = add2(add2(lf, rt), add2(m, p));                                       // src/pde.py:73-77
        const f2 dx = sub2(rt, lf), dy = sub2(p, m);                                        // 2*gx, 2*gy (src/pde.py:172-173)
        if constexpr (MOMENTS) {
            const f2 lap = fma2(bc(-4.0f), u, s4);
            const f2 gq = mul2(u, sub2(bc(1.0f), u)), hq = mul2(gq, u);
            pa[4] = fma2(lap, lap, pa[4]);
            ma[0] = fma2(lap, hq, ma[0]);
            ma[1] = fma2(lap, gq, ma[1]);
            ma[2] = fma2(hq, hq, ma[2]);
            ma[3] = fma2(hq, gq, ma[3]);
        } else {
            const f2 r = fma2(u, fma2(u, sub2(bc(a1), u), bc(c0)), mul2(bc(D), s4));        // src/pde.py:99,:120
            pa[4] = fma2(r, r, pa[4]);
        }
        pa[5] = fma2(dx, dx, pa[5]);
        pa[5] = fma2(dy, dy, pa[5]);
    }
    __device__ __forceinline__ void stencil4(const float4& um, const float4& uc, const float4& up) {
        const float L = __shfl_up_sync(0xffffffffu, uc.w, 1);
        const float R = __shfl_down_sync(0xffffffffu, uc.x, 1);
        const f2 A = make_float2(L, uc.x), B = make_float2(uc.y, uc.z), C = make_float2(uc.w, R);
        stencil2(make_float2(uc.x, uc.y), A, B, make_float2(um.x, um.y), make_float2(up.x, up.y));
        stencil2(make_float2(uc.z, uc.w), B, C, make_float2(um.z, um.w), make_float2(up.z, up.w));
    }

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
            if constexpr (MOMENTS) {
                const float w = ALIGNED ? 1.0f : m[p];
                const float hq = uv * u;
                acc[8] = fmaf(lap * hq, w, acc[8]);
                acc[9] = fmaf(lap * uv, w, acc[9]);
                acc[10] = fmaf(hq * hq, w, acc[10]);
                acc[11] = fmaf(hq * uv, w, acc[11]);
            }
            const float r4 = MOMENTS ? lap : r;  // slot 4 holds sum lap^2 in MOMENTS mode
            if constexpr (ALIGNED) {
                acc[0] = fmaf(u, t, acc[0]);
                acc[1] += u;
                acc[2] += t;
                acc[3] += b;
                acc[4] = fmaf(r4, r4, acc[4]);
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
                acc[4] = fmaf(r4 * r4, w, acc[4]);
                acc[5] = fmaf(dx * dx + dy * dy, w, acc[5]);
                acc[6] = fmaf(uv * uv, w, acc[6]);
                if constexpr (KIND == PIL_X_PROB) acc[7] += (u >= 0.0f && u <= 1.0f) ? 0.0f : w;
            }
        }
    }
};

template <int KIND, typename XT, typename TT, bool ALIGNED, bool MOMENTS>
__global__ void __launch_bounds__(kThreads, MOMENTS ? kFwdMinBlocks - 1 : kFwdMinBlocks) pil_fwd_kernel(const FwdArgs A) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Geo& g = A.g;
    const long long task = (long long)blockIdx.x * kWarpsPerBlock + warp;

    constexpr int NACC = MOMENTS ? 16 : 8;
    FwdRow<KIND, ALIGNED, MOMENTS> fr;
#pragma unroll
    for (int k = 0; k < NACC; ++k) fr.acc[k] = 0.f;
    fr.init_packed();
    fr.D = A.D;
    fr.a = A.a;
    fr.a1 = 1.0f + A.a;
    fr.c0 = -A.a - 4.0f * A.D;

    if (task < g.tasks) {
        const int strip = (int)(task % g.strips);
        const long long grp = task / g.strips;
        long long pos = (g.total_rows * grp) / g.groups;              // flattened image row b*H + r
        const long long end = (g.total_rows * (grp + 1)) / g.groups;
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
#pragma unroll 1
      while (pos < end) {  // one segment per image the range touches (normally one, at most a few)
        const int b = (int)(pos / H);
        const int r0 = (int)(pos - (long long)b * H);
        const int r1 = (int)min((long long)H, (long long)r0 + (end - pos));
        pos += r1 - r0;

        // per-image base pointers at this thread's column; rows are addressed with 32-bit offsets
        const int coff = ALIGNED ? cx.colc : 0;
        const XT* xb = reinterpret_cast<const XT*>(A.x) + (long long)b * H * W + coff;
        const TT* tb = reinterpret_cast<const TT*>(A.t) + (long long)b * H * W + coff;
        auto xrow = [&](int k) -> const XT* { return xb + (unsigned)(mirror_clamp(k, H) * W); };
        auto trow = [&](int k) -> const TT* { return tb + (unsigned)(min(k, H - 1) * W); };

        if constexpr (ALIGNED) {
            // ---- staged path: iteration i consumes stage (i-r0)%6 = {map row i+1, target row i+1} ----
            // It activates row i+1 and adds that row's per-pixel terms (if the segment owns it), then adds
            // the stencil terms of row i, whose three u rows are now all in registers.
            extern __shared__ __align__(16) unsigned char smem_raw[];
            StageRing<XT, TT> ring;
            ring.init(smem_raw, warp, lane);
            const XT* px = xb + (unsigned)((r0 + 1) * W);  // map row of the next iteration to be issued
            const TT* pt = tb + (unsigned)((r0 + 1) * W);  // target row of the next iteration to be issued
            // issue(j): the copies iteration j will consume; CHECK handles the segment/image end
            auto issue = [&](int j, int stage, auto check) {
                constexpr bool CHECK = decltype(check)::value;
                if constexpr (CHECK) {
                    if (j < r1) ring.issue_x(stage, (j + 1 == H) ? px - 2 * W : px);  // row H := row H-2 (mirror)
                    if (j + 1 < r1) ring.issue_t(stage, pt);
                } else {
                    ring.issue_x(stage, px);
                    ring.issue_t(stage, pt);
                }
                px += W;
                pt += W;
                cp_async_commit();
            };
            const float4 x0 = cx.template load<XT>(xrow(r0 - 1)), x1 = cx.template load<XT>(xrow(r0));
            const float4 t1 = cx.template load_plain<TT>(trow(r0));
#pragma unroll
            for (int q = 0; q < kStages - 1; ++q) issue(r0 + q, q, BoolC<true>{});
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 U0 = fr.point4(cx.fix(x0), zero4, false);  // row r0-1: halo row, contributes no pixel terms
            float4 U1 = fr.point4(cx.fix(x1), t1, true);      // row r0
            float4 U2;

            auto step = [&](int i, int stage, auto check, const float4& um, const float4& uc, float4& up) {
                constexpr bool CHECK = decltype(check)::value;
                issue(i + kStages - 1, (stage + kStages - 1) % kStages, check);
                cp_async_wait<kStages - 1>();
                const float4 xn = ring.read_x(stage), tn = ring.read_t(stage);
                up = fr.point4(cx.fix(xn), tn, !CHECK || (i + 1 < r1));
                fr.stencil4(um, uc, up);
            };
            int i = r0;
            // steady state: the issue of iteration i+5+5 must stay clean -> i + 11 <= r1 - 2
#pragma unroll 1
            for (; i + 2 * kStages <= r1 - 1; i += kStages) {
                step(i + 0, 0, BoolC<false>{}, U0, U1, U2);
                step(i + 1, 1, BoolC<false>{}, U1, U2, U0);
                step(i + 2, 2, BoolC<false>{}, U2, U0, U1);
                step(i + 3, 3, BoolC<false>{}, U0, U1, U2);
                step(i + 4, 4, BoolC<false>{}, U1, U2, U0);
                step(i + 5, 5, BoolC<false>{}, U2, U0, U1);
            }
            int stage = 0;
#pragma unroll 1
            for (; i < r1; ++i) {  // segment tail (and short segments): dynamic stage, explicit rotation
                step(i, stage, BoolC<true>{}, U0, U1, U2);
                stage = (stage + 1 == kStages) ? 0 : stage + 1;
                U0 = U1;
                U1 = U2;
            }
            cp_async_wait<0>();
        } else {
            // prologue: rows r0-1, r0 become u; rows r0+1, r0+2 and targets r0, r0+1 are in flight.
            // Two fetch slots (A,B) alternate: iteration i consumes the slot holding map row i+1 / target
            // row i and immediately refills it with rows i+3 / i+2, so loaded registers are never moved
            // (a MOV of a loaded register would stall on the load and defeat the prefetch).
            const float4 x0 = cx.template load<XT>(xrow(r0 - 1)), x1 = cx.template load<XT>(xrow(r0));
            float4 xA = cx.template load<XT>(xrow(r0 + 1));
            float4 xB = cx.template load<XT>(xrow(min(r0 + 2, r1)));
            float4 tA = cx.template load_plain<TT>(trow(r0));
            float4 tB = cx.template load_plain<TT>(trow(r0 + 1));
            float4 U0 = act4<KIND>(cx.fix(x0)), U1 = act4<KIND>(cx.fix(x1)), U2;
            const XT* px = xb + (unsigned)((r0 + 3) * W);  // next map row to fetch (row i+3)
            const TT* pt = tb + (unsigned)((r0 + 2) * W);  // next target row to fetch (row i+2)

            // CHECK=false: steady state, every fetched row is inside the segment and the image.
            auto step = [&](int i, auto check, float4& xs, float4& ts, const float4& um, const float4& uc, float4& up) {
                constexpr bool CHECK = decltype(check)::value;
                const float4 xn = xs, tn = ts;
                if constexpr (CHECK) {
                    if (i + 3 <= r1) xs = cx.template load<XT>((i + 3 == H) ? px - 2 * W : px);  // row H := row H-2
                    if (i + 2 < r1) ts = cx.template load_plain<TT>(pt);
                } else {
                    xs = cx.template load<XT>(px);
                    ts = cx.template load_plain<TT>(pt);
                }
                px += W;
                pt += W;
                up = act4<KIND>(cx.fix(xn));
                fr.row(um, uc, up, tn);
            };
            int i = r0;
#pragma unroll 1
            for (; i + 6 <= r1 - 3; i += 6) {
                step(i + 0, BoolC<false>{}, xA, tA, U0, U1, U2);
                step(i + 1, BoolC<false>{}, xB, tB, U1, U2, U0);
                step(i + 2, BoolC<false>{}, xA, tA, U2, U0, U1);
                step(i + 3, BoolC<false>{}, xB, tB, U0, U1, U2);
                step(i + 4, BoolC<false>{}, xA, tA, U1, U2, U0);
                step(i + 5, BoolC<false>{}, xB, tB, U2, U0, U1);
            }
#pragma unroll 1
            for (; i < r1; ++i) {  // segment tail (and short segments): same step, explicit rotation
                step(i, BoolC<true>{}, xA, tA, U0, U1, U2);
                float4 sw = xA;
                xA = xB;
                xB = sw;
                sw = tA;
                tA = tB;
                tB = sw;
                U0 = U1;
                U1 = U2;
            }

        }
      }  // segments

        if constexpr (ALIGNED) {
            fr.fold_packed();
            if (!counted) {
#pragma unroll
                for (int k = 0; k < NACC; ++k) fr.acc[k] = 0.f;
            }
        }
    }

    // ---- deterministic cross-block reduction; the last block finalises --------------------------
    if constexpr (MOMENTS) {
        double raw[16];
        if (!reduce_to_last_block<kThreads, float, 16>(fr.acc, A.partials, A.ticket, raw)) return;
        if (threadIdx.x == 0) {  // layout: include/pil.h PIL_NMOMENTS
            double* mo = A.sums;
            mo[0] = raw[0];
            mo[1] = raw[1];
            mo[2] = raw[2];
            mo[3] = -(double)kLn2 * raw[3];
            mo[4] = raw[4];            // sum lap^2
            mo[5] = 0.25 * raw[5];     // sum gx^2 + gy^2
            mo[6] = raw[6];            // sum g^2 = sum u^2 (1-u)^2
            mo[7] = raw[7];            // n_invalid
            mo[8] = raw[8];            // sum lap*h
            mo[9] = raw[9];            // sum lap*g
            mo[10] = raw[10];          // sum h^2
            mo[11] = raw[11];          // sum h*g
            mo[12] = (double)g.B * (double)g.H * (double)g.W;
            mo[13] = mo[14] = mo[15] = 0.0;
            *A.ticket = 0u;
        }
        return;
    }
    double raw[PIL_NSUMS];
    if (!reduce_to_last_block<kThreads, float>(fr.acc, A.partials, A.ticket, raw)) return;
    if (threadIdx.x == 0) {
        double s[PIL_NSUMS];
        sums_from_raw(raw, A.p.epsilon, (double)g.B * (double)g.H * (double)g.W, s);
#pragma unroll
        for (int k = 0; k < PIL_NSUMS; ++k) A.sums[k] = s[k];
        if (A.loss_out != nullptr) finalize_device(s, s[7], A.p, A.loss_out);
        *A.ticket = 0u;
    }
}

// ------------------------------------------------------------------------------------------------
// K1L: pointwise forward ("light"): only the sums that do NOT need neighbours -- I, P, T, the BCE
// terms and the double well -- as a flat, fully coalesced stream over the shard (no halo lanes, no
// row structure, 8 B/px read once).  Used by the training path, where the backward kernel visits
// every stencil anyway and accumulates sum r^2 and sum |grad u|^2 there (pil_backward_accumulate),
// so the 5-point stencils are evaluated once per step instead of twice.
// ------------------------------------------------------------------------------------------------
constexpr long long kMaxPointBlocks = 2048;  // grid cap of the pointwise forward (workspace sizing)
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
        *A.ticket = 0u;
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
        *A.ticket = 0u;
    }
    if (A.X.world > 0) {
        __syncthreads();
        xchg_push(A.X, 0, s_push);
    }
}

// per-image Dice and IoU from the counts (src/metrics.py:66-70, src/evaluate.py:90-94)
__global__ void pil_image_metrics_kernel(const double* counts, long long B, double smooth, float* dice, float* iou) {
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
        const double I = counts[4 * b], P = counts[4 * b + 1], T = counts[4 * b + 2];
        if (dice) dice[b] = (float)((2.0 * I + smooth) / (P + T + smooth));
        if (iou) iou[b] = (float)((I + smooth) / (P + T - I + smooth));
    }
}

// ------------------------------------------------------------------------------------------------
// K2: fused backward
// ------------------------------------------------------------------------------------------------
struct BwdCoef {
    float alpha, beta;   // dice: d/du = alpha*t + beta                    (du space)
    float cb;            // bce : cb*(u-t)/max(uv,1e-12)                   (du space)
    float cA;            // rd  : cA * (L^T r)      cA = scale*lrd*2/N*D
    float f3, f2, f1;    // rd  : cF * f'(u) = f3 u^2 + f2 u + f1,  cF = scale*lrd*2/N, f' = -3u^2 + 2(1+a)u - a
    float cG;            // pf  : cG * (dx[p-1]-dx[p+1] + dy[i-1]-dy[i+1]),  cG = scale*lpf/N*eps/4
    float cW;            // pf  : cW * uv*(1-2u),   cW = scale*lpf/N*2/eps
    float D;
    float a1, c0;        // r = u*(u*(a1 - u) + c0) + D*(sum of 4 neighbours),  a1 = 1+a, c0 = -a - 4D
    // packed path only
    float beta_half;     // beta / 2 (travels with the horizontal transposed-stencil terms)
    float f1c;           // f1 - 4 cA  (centre tap of the transposed Laplacian folded into f')
    float cW2n;          // -2 cW       (cW uv (1-2u) = uv (cW2n u + cW))
};

// accumulate mode: reduce the stencil sums across the grid; the last block publishes them and,
// if asked, assembles the loss from (global pointwise sums + these) -- src/loss.py:144-160.
__device__ __noinline__ void bwd_epilogue(const BwdArgs& A, const double* acc, const double* gs) {
    double raw[PIL_NSUMS];
    if (!reduce_to_last_block<kThreads, double>(acc, A.partials, A.ticket, raw)) return;
    __shared__ double s_push[PIL_NSUMS];   // this shard's stencil sums
    __shared__ double s_glob[PIL_NSUMS];   // the global stencil sums
    if (threadIdx.x == 0) {
        double sb[PIL_NSUMS] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        sb[4] = raw[4];
        sb[5] = (A.p.epsilon / 8.0) * raw[5];
#pragma unroll
        for (int k = 0; k < PIL_NSUMS; ++k) {
            A.stencil_sums[k] = sb[k];
            s_push[k] = sb[k];
            s_glob[k] = sb[k];
        }
    }
    bool finalize = true;
    if (A.X.world > 0) {  // data parallel: swap the stencil sums with every rank, then finalise globally
        __syncthreads();
        xchg_push(A.X, 1, s_push);
        if (A.X.defer) {
            finalize = false;
        } else {
            xchg_wait_sum(A.X, 1, s_glob);
        }
    }
    if (threadIdx.x == 0) {
        if ((A.loss_out != nullptr || A.total_sums != nullptr) && finalize) {
            double tot[PIL_NSUMS];
#pragma unroll
            for (int k = 0; k < PIL_NSUMS; ++k) tot[k] = gs[k] + s_glob[k];
            if (A.loss_out != nullptr) finalize_device(tot, A.n_global > 0 ? (double)A.n_global : tot[7], A.p, A.loss_out);
            if (A.total_sums != nullptr) {  // every block has read gsums long before the last one gets here
#pragma unroll
                for (int k = 0; k < PIL_NSUMS; ++k) A.total_sums[k] = tot[k];
            }
        }
        if (A.task_counter != nullptr) *A.task_counter = 0u;  // every warp has made its last claim
        *A.ticket = 0u;
    }
}

template <int KIND, typename XT, typename TT, bool ALIGNED>
__global__ void __launch_bounds__(kThreads, kBwdMinBlocks) pil_bwd_kernel(const BwdArgs A) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Geo& g = A.g;
    const long long task = (long long)blockIdx.x * kWarpsPerBlock + warp;
    TL_STAMP(1, 0);

    // task id -> (strip, first row, end row): equal row ranges (+-1 row), `strips` tasks per range
    // Tasks are handed out from the END of the shard backwards: the pointwise forward is a front-to-back
    // stream, so the last ~100 MB of x and t it read are still in the 126 MB L2 when this kernel starts.
    auto decode = [&](long long tk, int& strip_o, long long& pos_o, long long& end_o) {
        if (A.reverse) tk = g.tasks - 1 - tk;
        strip_o = (int)(tk % g.strips);
        const long long grp = tk / g.strips;
        pos_o = (g.total_rows * grp) / g.groups;
        end_o = (g.total_rows * (grp + 1)) / g.groups;
    };
    // pull the rows a range starts with (2 halo rows + the pipeline depth) into L2; no architectural effect
    auto prefetch_rows = [&](int strip_p, long long pos_p) {
        if ((lane & 7) == 0 || lane == 31) {
            const int colp = min(max(strip_p * kStripCols + (lane - 1) * kVec, 0), g.W - 1);
            const char* xp = reinterpret_cast<const char*>(A.x) + (pos_p * g.W + colp) * (long long)sizeof(XT);
            const char* tp = reinterpret_cast<const char*>(A.t) + (pos_p * g.W + colp) * (long long)sizeof(TT);
#pragma unroll
            for (int q = -2; q < kStages; ++q) {
                if (pos_p + q >= 0 && pos_p + q < g.total_rows) {
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(xp + (long long)q * g.W * (long long)sizeof(XT)));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(tp + (long long)q * g.W * (long long)sizeof(TT)));
                }
            }
        }
    };

    // ---- PDL prologue: nothing an earlier kernel wrote may be READ before pdl_wait(), but the rows this
    // warp starts with can already be pulled into L2 (a prefetch has no architectural effect), so the
    // DRAM latency of the pipeline fill overlaps the tail of the previous kernel.
    if (task < g.tasks) {
        int strip_p;
        long long pos_p, end_p;
        decode(task, strip_p, pos_p, end_p);
        prefetch_rows(strip_p, pos_p);
    }
    pdl_wait();
    pdl_launch_dependents();
    TL_STAMP(1, 1);

    // ---- global sums: given, or (data parallel) collected from the peer mailbox -----------------
    __shared__ double s_gs[PIL_NSUMS];
    const double* gs = A.gsums;
    if (A.X.world > 0) {
        xchg_wait_sum(A.X, 0, s_gs);
        gs = s_gs;
    }

    if (task >= g.tasks) {
        if (A.accumulate) {  // idle warp of the last block still takes part in the block reduction
            const double zero[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            bwd_epilogue(A, zero, gs);
        }
        return;
    }

    // ---- coefficients from the global sums (double once per thread, then fp32) -----------------
    BwdCoef c;
    {
        const double I = gs[0], P = gs[1], T = gs[2];
        const double s = A.p.smooth, den = P + T + s;
        double scale = (double)A.grad_scale * (A.upstream ? (double)__ldg(A.upstream) : 1.0);
        if (KIND == PIL_X_LOGITS_TANH) scale *= 2.0;  // d/dz sigmoid(2z) = 2 u (1-u)
        const double invN = 1.0 / (A.n_global > 0 ? (double)A.n_global : gs[7]);
        const bool use_rd = A.p.pde_weight > 0.0, use_pf = A.p.phase_field_weight > 0.0;
        c.alpha = (float)(scale * A.p.dice_weight * (-2.0 / den));
        c.beta = (float)(scale * A.p.dice_weight * (2.0 * I + s) / (den * den));
        c.cb = (float)(scale * A.p.bce_weight * invN);
        const double crd = use_rd ? scale * A.p.pde_weight * 2.0 * invN : 0.0;
        c.cA = (float)(crd * A.p.diffusion_coeff);
        c.f3 = (float)(-3.0 * crd);
        c.f2 = (float)(2.0 * (1.0 + A.p.reaction_threshold) * crd);
        c.f1 = (float)(-A.p.reaction_threshold * crd);
        c.cG = use_pf ? (float)(scale * A.p.phase_field_weight * invN * A.p.epsilon * 0.25) : 0.f;
        c.cW = use_pf ? (float)(scale * A.p.phase_field_weight * invN * 2.0 / A.p.epsilon) : 0.f;
        c.D = (float)A.p.diffusion_coeff;
        c.a1 = (float)(1.0 + A.p.reaction_threshold);
        c.c0 = (float)(-A.p.reaction_threshold - 4.0 * A.p.diffusion_coeff);
        c.beta_half = 0.5f * c.beta;
        c.f1c = (float)(-A.p.reaction_threshold * crd - 4.0 * crd * A.p.diffusion_coeff);
        c.cW2n = -2.0f * c.cW;
    }

    // ---- task loop.  Equal static row ranges finish far apart (measured: 83..145 us per block at
    // 64x1024^2 -- SMs do not get equal shares of the memory system), and with one resident wave nothing
    // evens that out.  So the ranges are made short and, after its first one, every warp claims the next
    // unprocessed (range, strip) task from a global counter until none is left.
    const int H = g.H, W = g.W;
    const bool out_lane = (lane >= 1) && (lane <= kOutLanes);
    double tot_r2 = 0.0, tot_g2 = 0.0;  // stencil sums over all tasks of this thread (task partials are fp32)
    long long task_cur = task;
#pragma unroll 1
  for (;;) {
    int strip;
    long long pos, end;  // flattened image rows b*H + r of this task
    decode(task_cur, strip, pos, end);
    // claim the NEXT task now and pull its first rows into L2, so that the pipeline fill of the next
    // range costs an L2 round trip instead of a DRAM one (a range is only a few tens of microseconds)
    const int col0 = strip * kStripCols + (lane - 1) * kVec;
    const bool in_img = col0 >= 0 && col0 < W;  // ALIGNED: whole vector in the image

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
    // packed path: the same factors folded into per-slot coefficient pairs
    const f2 cAf2[2] = {make_float2(c.cA * fc[0], c.cA * fc[1]), make_float2(c.cA * fc[2], c.cA * fc[3])};
    const f2 cGm2[2] = {make_float2(c.cG * mc[0], c.cG * mc[1]), make_float2(c.cG * mc[2], c.cG * mc[3])};
    const bool store_vec = ALIGNED && out_lane && in_img;
    // stencil sums of the rows this warp owns (accumulate mode): sum r^2 and sum dx^2+dy^2
    f2 sr2 = make_float2(0.f, 0.f), sg2 = make_float2(0.f, 0.f);
    float sr2s = 0.f, sg2s = 0.f;  // scalar path

#pragma unroll 1
  while (pos < end) {  // one segment per image the range touches (normally one, at most a few)
    const int b = (int)(pos / H);
    const int r0 = (int)(pos - (long long)b * H);
    const int r1 = (int)min((long long)H, (long long)r0 + (end - pos));
    pos += r1 - r0;
    const int coff = ALIGNED ? cx.colc : 0;
    const XT* xb = reinterpret_cast<const XT*>(A.x) + (long long)b * H * W + coff;
    const TT* tb = reinterpret_cast<const TT*>(A.t) + (long long)b * H * W + coff;
    XT* gb = reinterpret_cast<XT*>(A.grad) + (long long)b * H * W;  // column added at the store
    auto xrow = [&](int k) -> const XT* { return xb + (unsigned)(mirror_clamp(k, H) * W); };
    auto trow = [&](int k) -> const TT* { return tb + (unsigned)(min(k, H - 1) * W); };

    // iteration k forms the residual row k (needs u rows k-1, k, k+1) and emits gradient row k-1.
    // k runs r0-1 .. r1; u rows r0-2 .. r1+1 are read (mirrored at the image edge).
    // The u rows and the gradient accumulators live in rings of three that are renamed, not moved, in
    // the 6x unrolled steady state.  Rows are fetched 5 iterations ahead: through the cp.async stage
    // ring on the ALIGNED path, through two alternating register slots otherwise.
    const int k0 = r0 - 1;
    const float4 xa = cx.template load<XT>(xrow(k0 - 1)), xbq = cx.template load<XT>(xrow(k0));
    XT* pg = gb + (unsigned)(r0 * W) + (ALIGNED ? col0 : 0);  // next gradient row to store (row k-1)
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 U0, U1, U2;                  // u rows k-1, k, k+1 (the step produces k+1)
    float4 G0 = zero4, G1 = zero4, G2;  // gradient accumulators of rows k-1, k, k+1

    // compute part of one iteration, given the freshly fetched map row k+1 (xn) and, when a row is
    // emitted, target row k-1 (tn).  CHECK=false: steady state -- row k strictly inside the image.
    auto compute_scalar = [&](int k, auto check, const float4& xn, const float4& tn, const float4& ua, const float4& ub,
                       float4& uc4, float4& gm4, float4& g04, float4& gp4) {
        constexpr bool CHECK = decltype(check)::value;
        uc4 = act4<KIND>(cx.fix(xn));  // row k+1
        const float va[4] = {ua.x, ua.y, ua.z, ua.w};
        const float vc[4] = {uc4.x, uc4.y, uc4.z, uc4.w};
        float gm[4] = {gm4.x, gm4.y, gm4.z, gm4.w};
        float g0[4] = {g04.x, g04.y, g04.z, g04.w};
        float gp[4];
        if (!CHECK || (k >= 0 && k < H)) {
            const float L = __shfl_up_sync(0xffffffffu, ub.w, 1);
            const float R = __shfl_down_sync(0xffffffffu, ub.x, 1);
            const float e[6] = {L, ub.x, ub.y, ub.z, ub.w, R};
            float r[4], rc[4], dx[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const float u = e[p + 1];
                const float s4 = (e[p] + e[p + 2]) + (va[p] + vc[p]);
                // r = D*(s4 - 4u) + u(1-u)(u-a), as a polynomial in u   (src/pde.py:73-77,:99,:120)
                r[p] = fmaf(u, fmaf(u, c.a1 - u, c.c0), c.D * s4);
                rc[p] = r[p];
                dx[p] = e[p + 2] - e[p];
            }
            if (k >= r0 && k < r1) {  // rows this segment owns: the loss terms the light forward skipped
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float w = out_lane ? mc[p] : 0.0f, dy = vc[p] - va[p];
                    sr2s = fmaf(r[p] * r[p], w, sr2s);
                    sg2s = fmaf(dx[p] * dx[p] + dy * dy, w, sg2s);
                }
            }
            if constexpr (ALIGNED) {
                // only slots 0 and 3 can be an image-edge column or feed a neighbour lane; four
                // unconditional multiplies by per-thread constants (1 everywhere but at the edges)
                // beat a predicated block, which ptxas expands to 8 issue slots per row in every warp.
                rc[0] = r[0] * fc[0];
                rc[3] = r[3] * fc[3];
                dx[0] *= mc[0];
                dx[3] *= mc[3];
            } else {
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
                const float ey = c.cG * (vc[p] - va[p]);
                gp[p] = fmaf(cAr, r[p], ey);           // into row k+1
                gm[p] = fmaf(cAr, r[p], gm[p] - ey);   // into row k-1
                const float fpr = fmaf(u, fmaf(c.f3, u, c.f2), c.f1);  // cF * f'(u)
                float acc = g0[p];
                acc = fmaf(c.cA, (re[p] + re[p + 2]) - 4.0f * r[p], acc);
                acc = fmaf(fpr, r[p], acc);
                acc = fmaf(c.cG, de[p] - de[p + 2], acc);
                g0[p] = acc;
            }
        } else {
#pragma unroll
            for (int p = 0; p < 4; ++p) gp[p] = 0.f;
        }
        g04 = make_float4(g0[0], g0[1], g0[2], g0[3]);
        gp4 = make_float4(gp[0], gp[1], gp[2], gp[3]);

        if (!CHECK || k - 1 >= r0) {
            // emit gradient row k-1 : pointwise terms + accumulated stencil terms, then the chain factor
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
    };

    // Packed (fp32x2) form of the same iteration for the ALIGNED path.  Pairs are (slot0,slot1) and
    // (slot2,slot3).  The kernel is bound by FMA-pipe cycles (a packed op holds the pipe for two), so the
    // arithmetic is arranged to minimise them:
    //   * everything that combines HORIZONTAL neighbours is done with scalar adds on the six values
    //     {L, s0..s3, R}: a scalar add costs the pipe what half a packed add does, and it avoids the
    //     register moves that forming misaligned pairs (L,s0) (s1,s2) (s3,R) would need;
    //   * the transposed horizontal stencils travel as two combined quantities instead of four:
    //         Ah = cA*fc*r + cG*mc*dx   goes to the RIGHT neighbour,   Bh = cA*fc*r - cG*mc*dx   to the LEFT
    //     (fc/mc: image-edge column factors, folded into per-slot constants), 2 shuffles instead of 4;
    //   * the Dice constant beta rides along: Ah and Bh each carry beta/2 and every pixel receives
    //     exactly one of each, so the emission needs no separate "+ beta";
    //   * -4*cA*r joins the reaction derivative: (cF f'(u) - 4 cA) * r, one FMA.
    auto compute_packed = [&](int k, auto check, const float4& xn, const float4& tn, const float4& ua, const float4& ub,
                              float4& uc4, float4& gm4, float4& g04, float4& gp4) {
        constexpr bool CHECK = decltype(check)::value;
        uc4 = act4<KIND>(cx.fix(xn));  // row k+1
        const f2 va[2] = {make_float2(ua.x, ua.y), make_float2(ua.z, ua.w)};
        const f2 vc[2] = {make_float2(uc4.x, uc4.y), make_float2(uc4.z, uc4.w)};
        f2 gm[2] = {make_float2(gm4.x, gm4.y), make_float2(gm4.z, gm4.w)};
        f2 g0[2] = {make_float2(g04.x, g04.y), make_float2(g04.z, g04.w)};
        f2 gp[2];
        if (!CHECK || (k >= 0 && k < H)) {
            const float L = __shfl_up_sync(0xffffffffu, ub.w, 1);
            const float R = __shfl_down_sync(0xffffffffu, ub.x, 1);
            const f2 u[2] = {make_float2(ub.x, ub.y), make_float2(ub.z, ub.w)};
            // horizontal neighbour sums and differences, scalar (src/pde.py:73-77, :172)
            const f2 hs[2] = {make_float2(L + ub.y, ub.x + ub.z), make_float2(ub.y + ub.w, ub.z + R)};
            const f2 dx[2] = {make_float2(ub.y - L, ub.z - ub.x), make_float2(ub.w - ub.y, R - ub.z)};
            f2 r[2], dy[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const f2 s4 = add2(hs[h], add2(va[h], vc[h]));
                // r = D*(s4 - 4u) + u(1-u)(u-a), as a polynomial in u   (src/pde.py:73-77,:99,:120)
                r[h] = fma2(u[h], fma2(u[h], sub2(bc(c.a1), u[h]), bc(c.c0)), mul2(bc(c.D), s4));
                dy[h] = sub2(vc[h], va[h]);
            }
            if (!CHECK || (k >= r0 && k < r1)) {  // rows this segment owns: loss terms the light forward skipped
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    sr2 = fma2(r[h], r[h], sr2);
                    sg2 = fma2(dx[h], dx[h], sg2);
                    sg2 = fma2(dy[h], dy[h], sg2);
                }
            }
            // row factor of the transposed vertical stencil; dy of an edge row is 0 by mirroring
            const float cAr = (CHECK && (k == 0 || k == H - 1)) ? 2.0f * c.cA : c.cA;
            f2 Ah[2], Bh[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const f2 qv = mul2(bc(cAr), r[h]);
                gp[h] = fma2(bc(c.cG), dy[h], qv);                     // into row k+1:  cA r + cG dy
                gm[h] = add2(gm[h], fma2(bc(-c.cG), dy[h], qv));       // into row k-1:  cA r - cG dy
                const f2 qf = fma2(cAf2[h], r[h], bc(c.beta_half));
                Ah[h] = fma2(cGm2[h], dx[h], qf);
                Bh[h] = fma2(make_float2(-cGm2[h].x, -cGm2[h].y), dx[h], qf);
            }
            const float AL = __shfl_up_sync(0xffffffffu, Ah[1].y, 1);    // from the pixel left of slot 0
            const float BR = __shfl_down_sync(0xffffffffu, Bh[0].x, 1);  // from the pixel right of slot 3
            const f2 hg[2] = {make_float2(AL + Bh[0].y, Ah[0].x + Bh[1].x), make_float2(Ah[0].y + Bh[1].y, Ah[1].x + BR)};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const f2 fpr = fma2(u[h], fma2(bc(c.f3), u[h], bc(c.f2)), bc(c.f1c));  // cF f'(u) - 4 cA
                g0[h] = fma2(fpr, r[h], add2(g0[h], hg[h]));
            }
        } else {
            gp[0] = gp[1] = make_float2(0.f, 0.f);
        }
        g04 = make_float4(g0[0].x, g0[0].y, g0[1].x, g0[1].y);
        gp4 = make_float4(gp[0].x, gp[0].y, gp[1].x, gp[1].y);

        if (!CHECK || k - 1 >= r0) {
            // emit gradient row k-1 : pointwise terms + accumulated stencil terms (beta included), chain factor
            const f2 vt[2] = {make_float2(tn.x, tn.y), make_float2(tn.z, tn.w)};
            f2 o[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const f2 u = va[h], t = vt[h];
                const f2 v = sub2(bc(1.0f), u);
                const f2 uv = mul2(u, v);
                // du = gm + alpha t + cW uv (1-2u)
                const f2 du = fma2(uv, fma2(bc(c.cW2n), u, bc(c.cW)), fma2(bc(c.alpha), t, gm[h]));
                const f2 w = mul2(bc(c.cb), sub2(u, t));
                if constexpr (KIND == PIL_X_PROB) {
                    const f2 inv = make_float2(rcp_approx(fmaxf(uv.x, 1e-12f)), rcp_approx(fmaxf(uv.y, 1e-12f)));
                    o[h] = fma2(w, inv, du);
                } else {
                    // (u-t)/max(uv,1e-12) * uv  ==  (u-t) * sat(uv*1e12)
                    const f2 m = make_float2(__saturatef(uv.x * 1e12f), __saturatef(uv.y * 1e12f));
                    o[h] = fma2(du, uv, mul2(w, m));
                }
            }
            if (store_vec) st4<XT>(pg, make_float4(o[0].x, o[0].y, o[1].x, o[1].y));
            pg += W;
        }
    };
    auto compute = [&](int k, auto check, const float4& xn, const float4& tn, const float4& ua, const float4& ub,
                       float4& uc4, float4& gm4, float4& g04, float4& gp4) {
        if constexpr (ALIGNED) {
            compute_packed(k, check, xn, tn, ua, ub, uc4, gm4, g04, gp4);
        } else {
            compute_scalar(k, check, xn, tn, ua, ub, uc4, gm4, g04, gp4);
        }
    };

    if constexpr (ALIGNED) {
        // ---- staged path: iteration k consumes stage (k-k0)%6 = {map row k+1, target row k-1} ----
        extern __shared__ __align__(16) unsigned char smem_raw[];
        StageRing<XT, TT> ring;
        ring.init(smem_raw, warp, lane);
        const XT* px = xb + (long long)(k0 + 1) * W;  // map row of the next iteration to be issued
        const TT* pt = tb + (long long)(k0 - 1) * W;  // target row of the next iteration to be issued
        auto issue = [&](int j, int stage, auto check) {
            constexpr bool CHECK = decltype(check)::value;
            if constexpr (CHECK) {
                if (j + 1 <= min(r1 + 1, H)) ring.issue_x(stage, (j + 1 == H) ? px - 2 * W : px);  // row H := row H-2
                if (j - 1 >= r0 && j - 1 < r1) ring.issue_t(stage, pt);
            } else {
                ring.issue_x(stage, px);
                ring.issue_t(stage, pt);
            }
            px += W;
            pt += W;
            cp_async_commit();
        };
#pragma unroll
        for (int q = 0; q < kStages - 1; ++q) issue(k0 + q, q, BoolC<true>{});
        U0 = act4<KIND>(cx.fix(xa));   // row k0-1
        U1 = act4<KIND>(cx.fix(xbq));  // row k0

        auto step = [&](int k, int stage, auto check, const float4& ua, const float4& ub, float4& uc4, float4& gm4,
                        float4& g04, float4& gp4) {
            issue(k + kStages - 1, (stage + kStages - 1) % kStages, check);
            cp_async_wait<kStages - 1>();
            const float4 xn = ring.read_x(stage), tn = ring.read_t(stage);
            compute(k, check, xn, tn, ua, ub, uc4, gm4, g04, gp4);
        };
        int k = k0, stage = 0;
        auto rot_step = [&](int kk) {
            step(kk, stage, BoolC<true>{}, U0, U1, U2, G0, G1, G2);
            stage = (stage + 1 == kStages) ? 0 : stage + 1;
            U0 = U1;
            U1 = U2;
            G0 = G1;
            G1 = G2;
        };
        rot_step(k++);  // k = r0-1: forms r[r0-1], emits nothing
        rot_step(k++);  // k = r0  : forms r[r0],   emits nothing
        // steady state: compute clean for k in [r0+1, r1-4]; issue (k+5) clean for k+5 <= r1-2
#pragma unroll 1
        for (; k + 5 <= r1 - 7; k += kStages) {
            step(k + 0, 2, BoolC<false>{}, U0, U1, U2, G0, G1, G2);
            step(k + 1, 3, BoolC<false>{}, U1, U2, U0, G1, G2, G0);
            step(k + 2, 4, BoolC<false>{}, U2, U0, U1, G2, G0, G1);
            step(k + 3, 5, BoolC<false>{}, U0, U1, U2, G0, G1, G2);
            step(k + 4, 0, BoolC<false>{}, U1, U2, U0, G1, G2, G0);
            step(k + 5, 1, BoolC<false>{}, U2, U0, U1, G2, G0, G1);
        }
#pragma unroll 1
        for (; k <= r1; ++k) rot_step(k);
        cp_async_wait<0>();
    } else {
        // ---- register path (scalar loads): two alternating fetch slots, two rows ahead ----
        float4 xA = cx.template load<XT>(xrow(k0 + 1));   // row k+1 of the first iteration
        float4 xB = cx.template load<XT>(xrow(k0 + 2));
        float4 tA = cx.template load_plain<TT>(trow(r0));  // consumed when row r0 is emitted (k = r0+1)
        float4 tB = cx.template load_plain<TT>(trow(r0 + 1));
        U0 = act4<KIND>(cx.fix(xa));
        U1 = act4<KIND>(cx.fix(xbq));
        const XT* px = xb + (unsigned)((k0 + 3) * W);  // next map row to fetch (row k+3)
        const TT* pt = tb + (unsigned)((r0 + 2) * W);  // next target row to fetch
#pragma unroll 1
        for (int k = k0; k <= r1; ++k) {
            const float4 xn = xA, tn = tA;
            if (k + 3 <= min(r1 + 1, H)) xA = cx.template load<XT>((k + 3 == H) ? px - 2 * W : px);
            px += W;
            const bool emits = k - 1 >= r0;
            if (emits) {
                if (k + 1 < r1) tA = cx.template load_plain<TT>(pt);
                pt += W;
            }
            compute(k, BoolC<true>{}, xn, tn, U0, U1, U2, G0, G1, G2);
            float4 sw = xA;
            xA = xB;
            xB = sw;
            if (emits) {
                sw = tA;
                tA = tB;
                tB = sw;
            }
            U0 = U1;
            U1 = U2;
            G0 = G1;
            G1 = G2;
        }
    }
  }  // segments

    if constexpr (ALIGNED) {
        if (store_vec) {
            tot_r2 += (double)(sr2.x + sr2.y);
            tot_g2 += (double)(sg2.x + sg2.y);
        }
    } else {
        tot_r2 += (double)sr2s;
        tot_g2 += (double)sg2s;
    }
    // claim the next unprocessed task.  (Claiming earlier -- to prefetch the next range -- was measured
    // to lose more than it gains: a task held in reserve is not available to a warp that runs dry.)
    if (A.task_counter == nullptr) break;
    unsigned int claimed = 0u;
    if (lane == 0) claimed = atomicAdd(A.task_counter, 1u);
    claimed = __shfl_sync(0xffffffffu, claimed, 0);
    task_cur = A.first_dynamic + (long long)claimed;
    if (task_cur >= g.tasks) break;
  }  // tasks

    TL_STAMP(1, 2);
    if (!A.accumulate) return;  // uniform: plain pil_backward
    double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    acc[4] = tot_r2;
    acc[5] = tot_g2;
    bwd_epilogue(A, acc, gs);
    TL_STAMP(1, 3);
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

// Losses of up to kSweepChunk parameter settings from one moments vector (pil_forward_moments):
//   sum r^2 = D^2 <lap,lap> + 2D <lap,h> - 2aD <lap,g> + <h,h> - 2a <h,g> + a^2 <g,g>      (r = D lap + h - a g)
//   sum pf  = (eps/2) sum |grad u|^2 + <g,g> / eps
constexpr int kSweepChunk = 32;
struct SweepParams {
    int n;
    PilParams p[kSweepChunk];
};
__global__ void pil_sweep_finalize_kernel(const double* mo, long long n_global, SweepParams sp, float* out) {
    const int k = threadIdx.x;
    if (blockIdx.x != 0 || k >= sp.n) return;
    const PilParams& p = sp.p[k];
    const double D = p.diffusion_coeff, a = p.reaction_threshold, eps = p.epsilon;
    double s[PIL_NSUMS];
    s[0] = mo[0];
    s[1] = mo[1];
    s[2] = mo[2];
    s[3] = mo[3];
    s[4] = D * D * mo[4] + 2.0 * D * mo[8] - 2.0 * a * D * mo[9] + mo[10] - 2.0 * a * mo[11] + a * a * mo[6];
    s[5] = (p.phase_field_weight > 0.0) ? 0.5 * eps * mo[5] + mo[6] / eps : 0.0;
    s[6] = mo[7];
    s[7] = mo[12];
    finalize_device(s, n_global > 0 ? (double)n_global : s[7], p, out + (size_t)k * PIL_NOUT);
}

// deferred finalisation of a data-parallel step: both exchanged vectors -> the global loss report
__global__ void __launch_bounds__(kThreads) pil_xchg_finalize_kernel(XchgDev X, long long n_global, PilParams p, float* out, double* total_sums) {
    __shared__ double s_a[PIL_NSUMS], s_b[PIL_NSUMS];
    xchg_wait_sum(X, 0, s_a);
    xchg_wait_sum(X, 1, s_b);
    if (threadIdx.x == 0) {
        double a[PIL_NSUMS];
#pragma unroll
        for (int k = 0; k < PIL_NSUMS; ++k) a[k] = s_a[k] + s_b[k];
        if (out != nullptr) finalize_device(a, n_global > 0 ? (double)n_global : a[7], p, out);
        if (total_sums != nullptr) {
#pragma unroll
            for (int k = 0; k < PIL_NSUMS; ++k) total_sums[k] = a[k];
        }
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
static size_t dtype_size(int d);
#ifdef PIL_BOUNDS
static void set_bounds(const void* x, const void* t, const void* grad, int64_t n, int x_dtype, int t_dtype) {
    cudaDeviceSynchronize();  // development build: serialise, the extents are globals
    const char* rd[4] = {(const char*)x, (const char*)x + n * dtype_size(x_dtype), (const char*)t, (const char*)t + n * dtype_size(t_dtype)};
    const char* wr[2] = {(const char*)grad, grad ? (const char*)grad + n * dtype_size(x_dtype) : (const char*)grad};
    cudaMemcpyToSymbol(g_brd, rd, sizeof(rd));
    cudaMemcpyToSymbol(g_bwr, wr, sizeof(wr));
}
#define PIL_SET_BOUNDS(x, t, g, n, xd, td) set_bounds(x, t, g, n, xd, td)
#else
#define PIL_SET_BOUNDS(x, t, g, n, xd, td)
#endif

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

static thread_local PilLaunchInfo t_info = {};
static long long g_kernels_launched = 0;
static int g_tune_fwd_rps = 0, g_tune_bwd_rps = 0;
static long long g_l2_keep_mb = -1;  // pil_set_l2_keep_mb; < 0: PIL_L2_KEEP_MB or the default

static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

// Size the row-range partition: one range-group per `strips` resident warps (waves = 1), so that
// every SM holds the same number of blocks; never fewer than kMinRows rows per range.
constexpr int kMinRows = 8;
static Geo make_geo(int64_t B, int64_t H, int64_t W, int resident_blocks, int forced_rows, int waves) {
    Geo g;
    g.B = (int)B;
    g.H = (int)H;
    g.W = (int)W;
    g.strips = (int)((W + kStripCols - 1) / kStripCols);
    g.total_rows = (long long)B * H;
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

template <typename K>
static int blocks_per_sm(K kernel, int smem_bytes) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kThreads, smem_bytes) != cudaSuccess || n < 1) n = 1;
    return n;
}

struct LaunchOut {
    int blocks = 0, rows = 0;
    int status = PIL_OK;  // PIL_ERR_WORKSPACE when the partials do not fit
};
// Rows per range the kernels like best (measured on B200, 64x1024^2 .. 128x2048^2): long enough to
// amortise the warm-up rows and the pipeline fill of a segment, short enough that the hardware block
// scheduler can still even out SM-to-SM speed differences with a few waves.
constexpr int kTargetRowsFwd = 256, kTargetRowsBwd = 200;
static int tuning_waves(bool bwd, int64_t B, int64_t H, int64_t W, int resident_blocks) {
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
static bool use_pdl() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("PIL_PDL");
        v = (e && atoi(e) == 0) ? 0 : 1;
    }
    return v == 1;
}
template <typename K, typename A>
static cudaError_t launch_pdl(K kernel, int blocks, int threads, int smem, cudaStream_t s, const A& args) {
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
    return cudaLaunchKernelEx(&cfg, kernel, args);
}

// Rows per dynamically claimed range of the backward kernel.  PIL_BWD_ROWS forces a value (0 = static
// partition).  Automatic: 64 rows -- long enough to amortise the 4 halo rows and the pipeline fill of a
// range, short enough to balance.  Measured on B200 at 64x1024^2 fp32: static 157.8 us; dynamic 24 rows
// 169.9, 32: 157.5, 48: 155.6, 61: 149-153, 63: 150.5, 64: 145-147, 67: 152, 96: 149.5, 128: 155.7 (the
// power of two wins over its neighbours: range boundaries then tile the 2 MB pages).  Shrinking ranges
// (guided self-scheduling), a small-range tail phase and claiming one range ahead to prefetch it were all
// measured slower.  Problems too small for ~2.5 waves of such ranges keep the static one-wave partition.
static int bwd_dynamic_rows(long long total_rows, long long strips, long long resident_warps) {
    static int forced = -2;
    if (forced == -2) {
        const char* e = getenv("PIL_BWD_ROWS");
        forced = e ? atoi(e) : -1;
    }
    if (forced >= 0) return (forced > 0 && forced < kMinRows) ? kMinRows : forced;
    const int rows = 64;
    const double waves = (double)(((total_rows + rows - 1) / rows) * strips) / (double)resident_warps;
    return waves < 2.5 ? 0 : rows;
}

static unsigned long long xchg_timeout_ns() {
    static unsigned long long v = 0;
    if (v == 0) {
        const char* e = getenv("PIL_XCHG_TIMEOUT_MS");
        const long long ms = (e && atoll(e) > 0) ? atoll(e) : 20000;
        v = (unsigned long long)ms * 1000000ull;
    }
    return v;
}
static int make_xchg(const PilExchange* ex, XchgDev* X) {
    *X = XchgDev{};
    if (!ex) return PIL_OK;
    if (ex->world < 1 || ex->world > PIL_MAX_RANKS || ex->rank < 0 || ex->rank >= ex->world) return PIL_ERR_EXCHANGE;
    X->rank = ex->rank;
    X->world = ex->world;
    X->parity = (int)(ex->epoch & 1ull);
    X->defer = (ex->flags & PIL_XCHG_DEFER_FINALIZE) ? 1 : 0;
    X->want = (ex->epoch % 0xfffffffeull) + 1ull;  // 32-bit step tag, never 0 (mailboxes start zeroed)
    X->timeout_ns = xchg_timeout_ns();
    for (int r = 0; r < ex->world; ++r) {
        if (!ex->mailbox[r]) return PIL_ERR_EXCHANGE;
        X->box[r] = reinterpret_cast<unsigned char*>(ex->mailbox[r]);
    }
    return PIL_OK;
}

template <int KIND, typename XT, typename TT>
static cudaError_t launch_fwd_a(FwdArgs& a, int64_t B, int64_t H, int64_t W, bool aligned, size_t partial_bytes_avail,
                                cudaStream_t s, LaunchOut* out, bool moments) {
    static int per_sm_cache[4] = {0, 0, 0, 0};  // per template instantiation x {scalar, aligned} x {sums, moments}
    auto go = [&](auto kernel, int smem) -> cudaError_t {
        int& per_sm = per_sm_cache[(aligned ? 1 : 0) + (moments ? 2 : 0)];
        if (per_sm == 0) {
            cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
            per_sm = blocks_per_sm(kernel, smem);
        }
        a.g = make_geo(B, H, W, sm_count() * per_sm, g_tune_fwd_rps, tuning_waves(false, B, H, W, sm_count() * per_sm));
        out->blocks = (int)((a.g.tasks + kWarpsPerBlock - 1) / kWarpsPerBlock);
        out->rows = (int)((a.g.total_rows + a.g.groups - 1) / a.g.groups);
        if ((size_t)out->blocks * (moments ? 16 : PIL_NSUMS) * sizeof(double) > partial_bytes_avail) {
            out->status = PIL_ERR_WORKSPACE;
            return cudaSuccess;
        }
        kernel<<<out->blocks, kThreads, smem, s>>>(a);
        return cudaGetLastError();
    };
    if (moments) {
        if (aligned) return go(pil_fwd_kernel<KIND, XT, TT, true, true>, kSmemPerBlock);
        return go(pil_fwd_kernel<KIND, XT, TT, false, true>, 0);
    }
    if (aligned) return go(pil_fwd_kernel<KIND, XT, TT, true, false>, kSmemPerBlock);
    return go(pil_fwd_kernel<KIND, XT, TT, false, false>, 0);
}
#define PIL_FWD_ARGS FwdArgs &a, int64_t B, int64_t H, int64_t W, bool aligned, size_t avail, cudaStream_t s, LaunchOut *out, bool moments
#define PIL_FWD_PASS a, B, H, W, aligned, avail, s, out, moments
template <int KIND, typename XT>
static cudaError_t launch_fwd_t(int t_dtype, PIL_FWD_ARGS) {
#ifdef PIL_DEV_F32_ONLY  // development builds: fp32 maps only (6x faster to compile)
    if (t_dtype != PIL_F32) return cudaErrorNotSupported;
    return launch_fwd_a<KIND, XT, float>(PIL_FWD_PASS);
#else
    switch (t_dtype) {
        case PIL_F32: return launch_fwd_a<KIND, XT, float>(PIL_FWD_PASS);
        case PIL_BF16: return launch_fwd_a<KIND, XT, __nv_bfloat16>(PIL_FWD_PASS);
        default: return launch_fwd_a<KIND, XT, uint8_t>(PIL_FWD_PASS);
    }
#endif
}
template <int KIND>
static cudaError_t launch_fwd_x(int x_dtype, int t_dtype, PIL_FWD_ARGS) {
    if (x_dtype == PIL_F32) return launch_fwd_t<KIND, float>(t_dtype, PIL_FWD_PASS);
#ifdef PIL_DEV_F32_ONLY
    return cudaErrorNotSupported;
#else
    return launch_fwd_t<KIND, __nv_bfloat16>(t_dtype, PIL_FWD_PASS);
#endif
}

template <int KIND, typename XT, typename TT>
static cudaError_t launch_bwd_a(BwdArgs& a, int64_t B, int64_t H, int64_t W, bool aligned, cudaStream_t s, LaunchOut* out) {
    static int per_sm_cache[2] = {0, 0};  // per template instantiation x {scalar, aligned} kernel
    auto go = [&](auto kernel, int smem) -> cudaError_t {
        int& per_sm = per_sm_cache[aligned ? 1 : 0];
        if (per_sm == 0) {
            cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
            per_sm = blocks_per_sm(kernel, smem);
        }
        const int resident = sm_count() * per_sm;
        const long long strips_ = (W + kStripCols - 1) / kStripCols;
        const int dyn_rows = a.task_counter != nullptr
                                 ? (g_tune_bwd_rps > 0 ? g_tune_bwd_rps : bwd_dynamic_rows(B * H, strips_, (long long)resident * kWarpsPerBlock))
                                 : 0;
        if (dyn_rows > 0) {
            // persistent grid of the resident blocks; short ranges claimed dynamically (see the kernel)
            a.g = make_geo(B, H, W, resident, dyn_rows, 1);
            const long long need = (a.g.tasks + kWarpsPerBlock - 1) / kWarpsPerBlock;
            out->blocks = (int)(need < resident ? need : resident);
            a.first_dynamic = (long long)out->blocks * kWarpsPerBlock;
        } else {
            a.task_counter = nullptr;
            a.first_dynamic = 0;
            a.g = make_geo(B, H, W, resident, g_tune_bwd_rps, tuning_waves(true, B, H, W, resident));
            out->blocks = (int)((a.g.tasks + kWarpsPerBlock - 1) / kWarpsPerBlock);
        }
        out->rows = (int)((a.g.total_rows + a.g.groups - 1) / a.g.groups);
        return launch_pdl(kernel, out->blocks, kThreads, smem, s, a);
    };
    if (aligned) return go(pil_bwd_kernel<KIND, XT, TT, true>, kSmemPerBlock);
    return go(pil_bwd_kernel<KIND, XT, TT, false>, 0);
}
#define PIL_BWD_ARGS BwdArgs &a, int64_t B, int64_t H, int64_t W, bool aligned, cudaStream_t s, LaunchOut *out
#define PIL_BWD_PASS a, B, H, W, aligned, s, out
template <int KIND, typename XT>
static cudaError_t launch_bwd_t(int t_dtype, PIL_BWD_ARGS) {
#ifdef PIL_DEV_F32_ONLY  // development builds: fp32 maps only (6x faster to compile)
    if (t_dtype != PIL_F32) return cudaErrorNotSupported;
    return launch_bwd_a<KIND, XT, float>(PIL_BWD_PASS);
#else
    switch (t_dtype) {
        case PIL_F32: return launch_bwd_a<KIND, XT, float>(PIL_BWD_PASS);
        case PIL_BF16: return launch_bwd_a<KIND, XT, __nv_bfloat16>(PIL_BWD_PASS);
        default: return launch_bwd_a<KIND, XT, uint8_t>(PIL_BWD_PASS);
    }
#endif
}
template <int KIND>
static cudaError_t launch_bwd_x(int x_dtype, int t_dtype, PIL_BWD_ARGS) {
    if (x_dtype == PIL_F32) return launch_bwd_t<KIND, float>(t_dtype, PIL_BWD_PASS);
#ifdef PIL_DEV_F32_ONLY
    return cudaErrorNotSupported;
#else
    return launch_bwd_t<KIND, __nv_bfloat16>(t_dtype, PIL_BWD_PASS);
#endif
}

template <int KIND, typename XT, typename TT>
static cudaError_t launch_point_a(const PointArgs& a, bool aligned, cudaStream_t s, int* blocks_out) {
    static int per_sm_cache[2] = {0, 0};
    auto go = [&](auto kernel) -> cudaError_t {
        int& per_sm = per_sm_cache[aligned ? 1 : 0];
        if (per_sm == 0) {
            int n = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, kPointThreads, 0) != cudaSuccess || n < 1) n = 1;
            per_sm = n;
        }
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

__global__ void __launch_bounds__(256) pil_scale_kernel(float* __restrict__ g, long long n4, const float* __restrict__ up) {
    const float s = __ldg(up);
    if (s == 1.0f) return;  // loss.backward() on the loss itself: nothing to do
    float4* g4 = reinterpret_cast<float4*>(g);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = g4[i];
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        g4[i] = v;
    }
}
__global__ void __launch_bounds__(256) pil_scale_kernel_generic(void* __restrict__ g, int is_bf16, long long n, const float* __restrict__ up) {
    const float s = __ldg(up);
    if (s == 1.0f) return;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (is_bf16) {
            __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(g) + i;
            *p = __float2bfloat16_rn(__bfloat162float(*p) * s);
        } else {
            reinterpret_cast<float*>(g)[i] *= s;
        }
    }
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

struct WorkspaceLayout {
    size_t ticket_off, scratch_off, partials_off, total;
};
static WorkspaceLayout workspace_layout(int64_t B, int64_t H, int64_t W) {
    // worst case number of forward blocks: 8-row segments
    const long long strips = (W + kStripCols - 1) / kStripCols;
    const long long segs = (H + 7) / 8;
    const long long blocks = (B * strips * segs + kWarpsPerBlock - 1) / kWarpsPerBlock;
    WorkspaceLayout l;
    l.ticket_off = 0;
    l.scratch_off = 64;   // PIL_NSUMS doubles of scratch (pil_loss_fwd_bwd)
    l.partials_off = 256;
    l.total = l.partials_off + (size_t)(blocks > kMaxPointBlocks ? blocks : kMaxPointBlocks) * PIL_NMOMENTS * sizeof(double);
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
        case PIL_ERR_EXCHANGE: return "bad PilExchange (rank/world out of range or a mailbox pointer is NULL)";
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

static int forward_impl(const void* x, const void* t, int64_t B, int64_t H, int64_t W, int x_dtype, int t_dtype, int x_kind,
                        const PilParams* p, double* sums, float* loss_out, void* workspace, size_t workspace_bytes,
                        void* stream, bool moments) {
    int st = check_common(x, t, B, H, W, x_dtype, t_dtype, x_kind, p);
    if (st != PIL_OK) return st;
    if (!sums || !workspace) return PIL_ERR_NULL;
    const WorkspaceLayout wl = workspace_layout(B, H, W);
    if (workspace_bytes < wl.total || ((uintptr_t)workspace % 8)) return PIL_ERR_WORKSPACE;

    FwdArgs a;
    a.x = x;
    a.t = t;
    a.D = (float)p->diffusion_coeff;
    a.a = (float)p->reaction_threshold;
    a.ticket = reinterpret_cast<unsigned int*>((char*)workspace + wl.ticket_off);
    a.partials = reinterpret_cast<double*>((char*)workspace + wl.partials_off);
    a.sums = sums;
    a.loss_out = loss_out;
    a.p = *p;
    PIL_SET_BOUNDS(x, t, nullptr, B * H * W, x_dtype, t_dtype);
    const bool aligned = is_aligned_case(x, t, nullptr, W, x_dtype, t_dtype);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t avail = workspace_bytes - wl.partials_off;
    LaunchOut lo;
    cudaError_t e;
    switch (x_kind) {
        case PIL_X_PROB: e = launch_fwd_x<PIL_X_PROB>(x_dtype, t_dtype, a, B, H, W, aligned, avail, s, &lo, moments); break;
        case PIL_X_LOGITS_SIGMOID: e = launch_fwd_x<PIL_X_LOGITS_SIGMOID>(x_dtype, t_dtype, a, B, H, W, aligned, avail, s, &lo, moments); break;
        default: e = launch_fwd_x<PIL_X_LOGITS_TANH>(x_dtype, t_dtype, a, B, H, W, aligned, avail, s, &lo, moments); break;
    }
    if (lo.status != PIL_OK) return lo.status;
    const int blocks = lo.blocks;
    t_info.fwd_blocks = blocks;
    t_info.fwd_threads = kThreads;
    t_info.fwd_rows_per_segment = lo.rows;
    t_info.fwd_aligned = aligned ? 1 : 0;
    ++g_kernels_launched;
    return (int)e;
}

int pil_forward(const void* x, const void* t, int64_t B, int64_t H, int64_t W, int x_dtype, int t_dtype, int x_kind,
                const PilParams* p, double* sums, float* loss_out, void* workspace, size_t workspace_bytes,
                void* stream) {
    return forward_impl(x, t, B, H, W, x_dtype, t_dtype, x_kind, p, sums, loss_out, workspace, workspace_bytes, stream, false);
}

// ---- parameter sweeps: one pass over the maps serves any number of (D, a, eps, weights) settings ----
int pil_forward_moments(const void* x, const void* t, int64_t B, int64_t H, int64_t W, int x_dtype, int t_dtype, int x_kind,
                        double* moments, void* workspace, size_t workspace_bytes, void* stream) {
    PilParams neutral = {0.5, 0.5, 0.0, 0.0, 1.0, 0.5, 1.0, 1e-6};  // the moments do not depend on any knob
    return forward_impl(x, t, B, H, W, x_dtype, t_dtype, x_kind, &neutral, moments, nullptr, workspace, workspace_bytes, stream, true);
}

int pil_sweep_finalize(const double* moments, int64_t n_global, const PilParams* params, int n_params, float* loss_out,
                       void* stream) {
    if (!moments || !params || !loss_out) return PIL_ERR_NULL;
    if (n_params < 1) return PIL_ERR_SHAPE;
    for (int k = 0; k < n_params; ++k) {
        const int st = pil_validate_params(params + k);
        if (st != PIL_OK) return st;
    }
    for (int k0 = 0; k0 < n_params; k0 += kSweepChunk) {
        SweepParams sp;
        sp.n = n_params - k0 < kSweepChunk ? n_params - k0 : kSweepChunk;
        for (int k = 0; k < sp.n; ++k) sp.p[k] = params[k0 + k];
        pil_sweep_finalize_kernel<<<1, kSweepChunk, 0, (cudaStream_t)stream>>>(moments, (long long)n_global, sp, loss_out + (size_t)k0 * PIL_NOUT);
        ++g_kernels_launched;
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    return PIL_OK;
}

int pil_finalize(const double* sums, int64_t n_global, const PilParams* p, float* loss_out, void* stream) {
    if (!sums || !p || !loss_out) return PIL_ERR_NULL;
    pil_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, (long long)n_global, *p, loss_out);
    ++g_kernels_launched;
    return (int)cudaGetLastError();
}

static int backward_impl(const void* x, const void* t, void* grad, int64_t B, int64_t H, int64_t W, int x_dtype, int t_dtype,
                         int x_kind, const PilParams* p, const double* global_sums, int64_t n_global, const float* upstream,
                         float grad_scale, double* stencil_sums, float* loss_out, double* total_sums, void* acc_ws,
                         void* stream, const PilExchange* ex = nullptr) {
    int st = check_common(x, t, B, H, W, x_dtype, t_dtype, x_kind, p);
    if (st != PIL_OK) return st;
    if (!grad || (!global_sums && !ex)) return PIL_ERR_NULL;
    if ((uintptr_t)grad % dtype_size(x_dtype)) return PIL_ERR_ALIGNMENT;

    BwdArgs a;
    a.x = x;
    a.t = t;
    a.grad = grad;
    a.gsums = global_sums;
    a.upstream = upstream;
    a.grad_scale = grad_scale;
    a.n_global = (long long)n_global;
    a.p = *p;
    a.accumulate = acc_ws != nullptr ? 1 : 0;
    {
        static int rev = -1;
        if (rev < 0) {
            const char* e = getenv("PIL_BWD_REVERSE");
            rev = e ? atoi(e) : 1;
        }
        a.reverse = rev;
    }
    a.partials = nullptr;
    a.ticket = nullptr;
    a.task_counter = nullptr;
    a.first_dynamic = 0;
    a.stencil_sums = stencil_sums;
    a.loss_out = loss_out;
    a.total_sums = total_sums;
    st = make_xchg(ex, &a.X);
    if (st != PIL_OK) return st;
    if (a.accumulate) {
        const WorkspaceLayout wl = workspace_layout(B, H, W);
        a.ticket = reinterpret_cast<unsigned int*>((char*)acc_ws + wl.ticket_off);
        a.task_counter = a.ticket + 1;
        a.partials = reinterpret_cast<double*>((char*)acc_ws + wl.partials_off);
    }
    PIL_SET_BOUNDS(x, t, grad, B * H * W, x_dtype, t_dtype);
    const bool aligned = is_aligned_case(x, t, grad, W, x_dtype, t_dtype);
    cudaStream_t s = (cudaStream_t)stream;
    LaunchOut lo;
    cudaError_t e;
    switch (x_kind) {
        case PIL_X_PROB: e = launch_bwd_x<PIL_X_PROB>(x_dtype, t_dtype, a, B, H, W, aligned, s, &lo); break;
        case PIL_X_LOGITS_SIGMOID: e = launch_bwd_x<PIL_X_LOGITS_SIGMOID>(x_dtype, t_dtype, a, B, H, W, aligned, s, &lo); break;
        default: e = launch_bwd_x<PIL_X_LOGITS_TANH>(x_dtype, t_dtype, a, B, H, W, aligned, s, &lo); break;
    }
    const int blocks = lo.blocks;
    t_info.bwd_blocks = blocks;
    t_info.bwd_threads = kThreads;
    t_info.bwd_rows_per_segment = lo.rows;
    t_info.bwd_aligned = aligned ? 1 : 0;
    ++g_kernels_launched;
    return (int)e;
}

int pil_backward(const void* x, const void* t, void* grad, int64_t B, int64_t H, int64_t W, int x_dtype, int t_dtype,
                 int x_kind, const PilParams* p, const double* global_sums, int64_t n_global, const float* upstream,
                 float grad_scale, void* stream) {
    return backward_impl(x, t, grad, B, H, W, x_dtype, t_dtype, x_kind, p, global_sums, n_global, upstream, grad_scale,
                         nullptr, nullptr, nullptr, nullptr, stream);
}

int pil_backward_accumulate(const void* x, const void* t, void* grad, int64_t B, int64_t H, int64_t W, int x_dtype,
                            int t_dtype, int x_kind, const PilParams* p, const double* global_sums, int64_t n_global,
                            const float* upstream, float grad_scale, double* stencil_sums, float* loss_out,
                            void* workspace, size_t workspace_bytes, void* stream) {
    if (!stencil_sums || !workspace) return PIL_ERR_NULL;
    if (B >= 1 && H >= 2 && W >= 2) {
        const WorkspaceLayout wl = workspace_layout(B, H, W);
        if (workspace_bytes < wl.total || ((uintptr_t)workspace % 8)) return PIL_ERR_WORKSPACE;
    }
    return backward_impl(x, t, grad, B, H, W, x_dtype, t_dtype, x_kind, p, global_sums, n_global, upstream, grad_scale,
                         stencil_sums, loss_out, nullptr, workspace, stream);
}

static int pointwise_impl(const void* x, const void* t, int64_t B, int64_t H, int64_t W, int x_dtype, int t_dtype,
                          int x_kind, const PilParams* p, double* sums, void* workspace, size_t workspace_bytes,
                          void* stream, const PilExchange* ex) {
    int st = check_common(x, t, B, H, W, x_dtype, t_dtype, x_kind, p);
    if (st != PIL_OK) return st;
    if (!sums || !workspace) return PIL_ERR_NULL;
    const WorkspaceLayout wl = workspace_layout(B, H, W);
    if (workspace_bytes < wl.total || ((uintptr_t)workspace % 8)) return PIL_ERR_WORKSPACE;
    PointArgs a;
    a.x = x;
    a.t = t;
    a.n = (long long)B * H * W;
    {
        long long keep_mb = g_l2_keep_mb;
        if (keep_mb < 0) {
            static long long env_mb = -2;
            if (env_mb == -2) {
                const char* e = getenv("PIL_L2_KEEP_MB");
                env_mb = e ? atoll(e) : 12;  // interleaved A/B at 64x1024^2: 8-16 MB per map 2-3% faster per step than 0, 40 no better
            }
            keep_mb = env_mb;
        }
        const long long keep4 = (keep_mb << 20) / 16;  // float4s of EACH map to keep
        a.keep_from4 = (keep_mb > 0 && (a.n >> 2) > 2 * keep4) ? (a.n >> 2) - keep4 : (a.n >> 2);
    }
    a.ticket = reinterpret_cast<unsigned int*>((char*)workspace + wl.ticket_off);
    a.partials = reinterpret_cast<double*>((char*)workspace + wl.partials_off);
    a.sums = sums;
    a.p = *p;
    st = make_xchg(ex, &a.X);
    if (st != PIL_OK) return st;
    // flat stream: only total size and base alignment matter
    PIL_SET_BOUNDS(x, t, nullptr, B * H * W, x_dtype, t_dtype);
    const bool aligned = (a.n % 4 == 0) && is_aligned_case(x, t, nullptr, 4, x_dtype, t_dtype);
    cudaStream_t s = (cudaStream_t)stream;
    int blocks = 0;
    cudaError_t e;
    switch (x_kind) {
        case PIL_X_PROB: e = launch_point_x<PIL_X_PROB>(x_dtype, t_dtype, a, aligned, s, &blocks); break;
        case PIL_X_LOGITS_SIGMOID: e = launch_point_x<PIL_X_LOGITS_SIGMOID>(x_dtype, t_dtype, a, aligned, s, &blocks); break;
        default: e = launch_point_x<PIL_X_LOGITS_TANH>(x_dtype, t_dtype, a, aligned, s, &blocks); break;
    }
    t_info.fwd_blocks = blocks;
    t_info.fwd_threads = kPointThreads;
    t_info.fwd_rows_per_segment = 0;
    t_info.fwd_aligned = aligned ? 1 : 0;
    ++g_kernels_launched;
    return (int)e;
}

int pil_forward_pointwise(const void* x, const void* t, int64_t B, int64_t H, int64_t W, int x_dtype, int t_dtype,
                          int x_kind, const PilParams* p, double* sums, void* workspace, size_t workspace_bytes,
                          void* stream) {
    return pointwise_impl(x, t, B, H, W, x_dtype, t_dtype, x_kind, p, sums, workspace, workspace_bytes, stream, nullptr);
}

// ---- data-parallel training step over the peer-memory exchange (no NCCL call, 2 launches) --------
size_t pil_exchange_bytes(void) { return (size_t)kXchgStatusOffset + 128; }

int pil_exchange_alloc(void** mailbox, void* ipc_handle_out) {
    if (!mailbox) return PIL_ERR_NULL;
    static_assert(sizeof(cudaIpcMemHandle_t) == PIL_IPC_HANDLE_BYTES, "ipc handle size");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, pil_exchange_bytes());
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(p, 0, pil_exchange_bytes());
    if (e == cudaSuccess && ipc_handle_out) e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(ipc_handle_out), p);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(p);
        return (int)e;
    }
    *mailbox = p;
    return PIL_OK;
}

int pil_exchange_open(const void* ipc_handle, void** peer_mailbox) {
    if (!ipc_handle || !peer_mailbox) return PIL_ERR_NULL;
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    return (int)cudaIpcOpenMemHandle(peer_mailbox, h, cudaIpcMemLazyEnablePeerAccess);
}

int pil_exchange_close(void* peer_mailbox) {
    if (!peer_mailbox) return PIL_ERR_NULL;
    return (int)cudaIpcCloseMemHandle(peer_mailbox);
}

int pil_exchange_free(void* mailbox) {
    if (!mailbox) return PIL_ERR_NULL;
    return (int)cudaFree(mailbox);
}

int pil_exchange_status(const void* mailbox, int* status_out, void* stream) {
    if (!mailbox || !status_out) return PIL_ERR_NULL;
    cudaError_t e = cudaMemcpyAsync(status_out, (const char*)mailbox + kXchgStatusOffset, sizeof(int), cudaMemcpyDeviceToHost,
                                    (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    return (int)e;
}

int pil_forward_pointwise_xchg(const void* x, const void* t, int64_t B, int64_t H, int64_t W, int x_dtype, int t_dtype,
                               int x_kind, const PilParams* p, double* sums, void* workspace, size_t workspace_bytes,
                               const PilExchange* ex, void* stream) {
    if (!ex) return PIL_ERR_NULL;
    return pointwise_impl(x, t, B, H, W, x_dtype, t_dtype, x_kind, p, sums, workspace, workspace_bytes, stream, ex);
}

int pil_backward_accumulate_xchg(const void* x, const void* t, void* grad, int64_t B, int64_t H, int64_t W, int x_dtype,
                                 int t_dtype, int x_kind, const PilParams* p, const PilExchange* ex, int64_t n_global,
                                 const float* upstream, float grad_scale, double* stencil_sums, float* loss_out,
                                 double* total_sums, void* workspace, size_t workspace_bytes, void* stream) {
    if (!ex || !stencil_sums || !workspace) return PIL_ERR_NULL;
    if (B >= 1 && H >= 2 && W >= 2) {
        const WorkspaceLayout wl = workspace_layout(B, H, W);
        if (workspace_bytes < wl.total || ((uintptr_t)workspace % 8)) return PIL_ERR_WORKSPACE;
    }
    return backward_impl(x, t, grad, B, H, W, x_dtype, t_dtype, x_kind, p, nullptr, n_global, upstream, grad_scale,
                         stencil_sums, loss_out, total_sums, workspace, stream, ex);
}

int pil_forward_pointwise_metrics(const void* x, const void* t, int64_t B, int64_t H, int64_t W, int x_dtype, int t_dtype,
                                  int x_kind, const PilParams* p, double* sums, double* image_counts, float threshold,
                                  void* workspace, size_t workspace_bytes, const PilExchange* ex, void* stream) {
    int st = check_common(x, t, B, H, W, x_dtype, t_dtype, x_kind, p);
    if (st != PIL_OK) return st;
    if (!sums || !workspace || !image_counts) return PIL_ERR_NULL;
    const WorkspaceLayout wl = workspace_layout(B, H, W);
    if (workspace_bytes < wl.total || ((uintptr_t)workspace % 8) || ((uintptr_t)image_counts % 8)) return PIL_ERR_WORKSPACE;
    PointMetricsArgs a;
    a.x = x;
    a.t = t;
    a.n = (long long)B * H * W;
    a.hw = (long long)H * W;
    a.ticket = reinterpret_cast<unsigned int*>((char*)workspace + wl.ticket_off);
    a.partials = reinterpret_cast<double*>((char*)workspace + wl.partials_off);
    a.sums = sums;
    a.p = *p;
    a.image_counts = image_counts;
    a.threshold = threshold;
    a.l2_stream = g_l2_keep_mb != 0 ? 1 : 0;
    st = make_xchg(ex, &a.X);
    if (st != PIL_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(image_counts, 0, (size_t)B * 4 * sizeof(double), s);
    if (e != cudaSuccess) return (int)e;
    PIL_SET_BOUNDS(x, t, nullptr, B * H * W, x_dtype, t_dtype);
    const bool aligned = (a.hw % 4 == 0) && is_aligned_case(x, t, nullptr, 4, x_dtype, t_dtype);
    int blocks = 0;
    switch (x_kind) {
        case PIL_X_PROB: e = launch_point_metrics_x<PIL_X_PROB>(x_dtype, t_dtype, a, aligned, s, &blocks); break;
        case PIL_X_LOGITS_SIGMOID: e = launch_point_metrics_x<PIL_X_LOGITS_SIGMOID>(x_dtype, t_dtype, a, aligned, s, &blocks); break;
        default: e = launch_point_metrics_x<PIL_X_LOGITS_TANH>(x_dtype, t_dtype, a, aligned, s, &blocks); break;
    }
    t_info.fwd_blocks = blocks;
    t_info.fwd_threads = kPointThreads;
    t_info.fwd_rows_per_segment = 0;
    t_info.fwd_aligned = aligned ? 1 : 0;
    ++g_kernels_launched;
    return (int)e;
}

int pil_image_metrics(const double* image_counts, int64_t B, double smooth, float* dice_out, float* iou_out, void* stream) {
    if (!image_counts || (!dice_out && !iou_out)) return PIL_ERR_NULL;
    if (B < 1) return PIL_ERR_SHAPE;
    const int blocks = (int)((B + 127) / 128 < 64 ? (B + 127) / 128 : 64);
    pil_image_metrics_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(image_counts, (long long)B, smooth, dice_out, iou_out);
    ++g_kernels_launched;
    return (int)cudaGetLastError();
}

int pil_exchange_finalize(const PilExchange* ex, int64_t n_global, const PilParams* p, float* loss_out, double* total_sums,
                          void* stream) {
    if (!ex || !p || (!loss_out && !total_sums)) return PIL_ERR_NULL;
    XchgDev X;
    int st = make_xchg(ex, &X);
    if (st != PIL_OK) return st;
    pil_xchg_finalize_kernel<<<1, kThreads, 0, (cudaStream_t)stream>>>(X, (long long)n_global, *p, loss_out, total_sums);
    ++g_kernels_launched;
    return (int)cudaGetLastError();
}

int pil_loss_fwd_bwd(const void* x, const void* t, void* grad, int64_t B, int64_t H, int64_t W, int x_dtype, int t_dtype,
                     int x_kind, const PilParams* p, double* sums, float* loss_out, void* workspace,
                     size_t workspace_bytes, void* stream) {
    if (!loss_out) return PIL_ERR_NULL;
    int st = pil_forward_pointwise(x, t, B, H, W, x_dtype, t_dtype, x_kind, p, sums, workspace, workspace_bytes, stream);
    if (st != PIL_OK) return st;
    const WorkspaceLayout wl = workspace_layout(B, H, W);
    double* scratch = reinterpret_cast<double*>((char*)workspace + wl.scratch_off);
    if (B >= 1 && H >= 2 && W >= 2 && (workspace_bytes < wl.total || ((uintptr_t)workspace % 8))) return PIL_ERR_WORKSPACE;
    // the backward's last block also writes sums := sums + stencil sums, so the caller ends up with the
    // same vector pil_forward would have produced -- no extra launch
    return backward_impl(x, t, grad, B, H, W, x_dtype, t_dtype, x_kind, p, sums, B * H * W, nullptr, 1.0f, scratch, loss_out,
                         sums, workspace, stream);
}

int pil_scale_gradient(void* grad, int dtype, int64_t n, const float* upstream, void* stream) {
    if (!grad || !upstream) return PIL_ERR_NULL;
    if (n < 1) return PIL_ERR_SHAPE;
    if (!(dtype == PIL_F32 || dtype == PIL_BF16)) return PIL_ERR_DTYPE;
    cudaStream_t s = (cudaStream_t)stream;
    const int blocks = sm_count() * 8;
    if (dtype == PIL_F32 && n % 4 == 0 && (uintptr_t)grad % 16 == 0)
        pil_scale_kernel<<<blocks, 256, 0, s>>>((float*)grad, (long long)(n >> 2), upstream);
    else
        pil_scale_kernel_generic<<<blocks, 256, 0, s>>>(grad, dtype == PIL_BF16, (long long)n, upstream);
    ++g_kernels_launched;
    return (int)cudaGetLastError();
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

#ifdef PIL_BOUNDS
int pil_debug_bounds(unsigned long long* out4) {  // {bad reads, bad writes, first bad address, 0}; resets the counters
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpyFromSymbol(out4, pil::g_berr, sizeof(unsigned long long) * 4);
    unsigned long long z[4] = {0, 0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(pil::g_berr, z, sizeof(z));
    return (int)e;
}
#endif

#ifdef PIL_TIMELINE
int pil_debug_timeline(void* buf) { return (int)cudaMemcpyToSymbol(pil::g_timeline, &buf, sizeof(buf)); }
#endif

int pil_set_l2_keep_mb(int mb) {
    g_l2_keep_mb = mb;
    return PIL_OK;
}

int pil_set_tuning(int fwd_rows_per_segment, int bwd_rows_per_segment) {
    g_tune_fwd_rps = fwd_rows_per_segment > 0 ? fwd_rows_per_segment : 0;
    g_tune_bwd_rps = bwd_rows_per_segment > 0 ? bwd_rows_per_segment : 0;
    return PIL_OK;
}

}  // extern "C"
