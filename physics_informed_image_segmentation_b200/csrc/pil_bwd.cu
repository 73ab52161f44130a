// pil_bwd.cu -- K2, the fused backward kernel (+ stencil sums in accumulate mode), and its launcher.
// Compiled once per input kind (-DPIL_KIND=0|1|2); without PIL_KIND all three kinds are instantiated here.
#include <stdio.h>

#include <type_traits>

#include "pil_common.cuh"

namespace pil {

constexpr int kBwdTailRange = 16, kBwdTailPercent = 10;  // tail phase of the dynamic partition (PIL_BWD_TAIL=rows,percent; 0,0 = off)
constexpr int kBwdTmaDefault = 1;  // staging of the aligned backward when nothing is forced: TMA boxes (A/B in DESIGN.md)

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
    // packed path works on rho = r / D = (sum of 4 neighbours) + u*(u*(a1/D - u/D) + c0/D): one multiply less per
    // pixel; every coefficient that meets r carries the factor D instead, and sum r^2 = D^2 sum rho^2
    float nDi, a1D, c0D;  // -1/D, a1/D, c0/D
    float cAD;            // cA * D
    float f3D, f2D, f1cD; // (cF f'(u) - 4 cA) * D as a polynomial in u
};

// accumulate mode: reduce the stencil sums across the grid; the last block publishes them and,
// if asked, assembles the loss from (global pointwise sums + these) -- src/loss.py:144-160.
// acc2 = {sum r^2, sum dx^2+dy^2} of this thread: the only two sums the backward produces, so the cross-block
// reduction moves 16 bytes per block instead of 64 (it is serial time at the very end of the kernel).
static __device__ __forceinline__ void bwd_epilogue(const BwdArgs& A, const double* acc2, const double* gs) {
    // the global pointwise sums the loss is assembled from: requested now (threads 0..7), used after the reduction's
    // L2 round trips instead of adding one more to the serial tail of the kernel
    const bool wants_total = A.loss_out != nullptr || A.total_sums != nullptr;
    double g_k = 0.0;
    if (threadIdx.x < PIL_NSUMS && wants_total) g_k = gs[threadIdx.x];
    double raw2[2];
    if (!reduce_to_last_block<kThreads, double, 2, 1>(acc2, A.partials, A.ticket, raw2)) return;
    __shared__ double s_push[PIL_NSUMS];   // this shard's stencil sums
    __shared__ double s_glob[PIL_NSUMS];   // the global stencil sums
    __shared__ double s_tot[PIL_NSUMS];    // global pointwise + stencil sums
    __shared__ double s_term[PIL_NSUMS];   // loss terms (finalize_term)
    if (threadIdx.x == 0) {
        double sb[PIL_NSUMS] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        sb[4] = raw2[0];
        sb[5] = (A.p.epsilon / 8.0) * raw2[1];
#pragma unroll
        for (int k = 0; k < PIL_NSUMS; ++k) {
            A.stencil_sums[k] = sb[k];
            s_push[k] = sb[k];
            s_glob[k] = sb[k];
        }
        if (A.task_counter != nullptr) *A.task_counter = 0u;  // every warp has made its last claim
    }
    __syncthreads();
    bool finalize = true;
    if (A.X.world > 0) {  // data parallel: swap the stencil sums with every rank, then finalise globally
        xchg_push(A.X, 1, s_push);
        if (A.X.defer) {
            finalize = false;
        } else {
            xchg_wait_sum(A.X, 1, s_glob);
        }
    }
    // the loss from (global pointwise sums + global stencil sums), src/loss.py:134-160: the four quotients are formed
    // by four threads side by side (same operations, same roundings as finalize_device)
    if (wants_total && finalize && threadIdx.x < 32) {
        if (threadIdx.x < PIL_NSUMS) {
            const double tk = g_k + s_glob[threadIdx.x];
            s_tot[threadIdx.x] = tk;
            if (A.total_sums != nullptr) A.total_sums[threadIdx.x] = tk;  // every block has read gsums long before the last one gets here
        }
        __syncwarp();
        if (A.loss_out != nullptr) {
            const double n = A.n_global > 0 ? (double)A.n_global : s_tot[7];
            if (threadIdx.x < 4) s_term[threadIdx.x] = finalize_term(s_tot, n, A.p, (int)threadIdx.x);
            __syncwarp();
            if (threadIdx.x == 0) finalize_from_terms(s_term, s_tot, A.p, A.loss_out);
        }
    }
    TL_STAMP(1, 6);  // last block only: loss assembled
    if (threadIdx.x == 0 && A.X.world > 0) xchg_advance_epoch(A.X);  // device-epoch mode: this rank has completed the step
}

// TMA (ALIGNED only): the rows travel as 2-D tensor-map boxes (cp.async.bulk.tensor, one elected lane per warp, an
// mbarrier per stage) instead of one 16-byte cp.async per lane per row -- see the staged paths below.
template <int KIND, typename XT, typename TT, bool ALIGNED, bool TMA>
__device__ __forceinline__ void bwd_body(const BwdArgs& A, [[maybe_unused]] const CUtensorMap* tmx, [[maybe_unused]] const CUtensorMap* tmt) {
    static_assert(ALIGNED || !TMA, "the TMA stage ring needs the aligned layout");
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform by construction: TMA operands live in uniform registers
    [[maybe_unused]] unsigned int tma_q = 0;  // TMA: boxes this warp has consumed so far (stage = q & 1, phase = (q >> 1) & 1)
    if constexpr (TMA) {
        extern __shared__ __align__(128) unsigned char smem_tma[];
        if (threadIdx.x == 0) {
            using Ring0 = TmaRing<XT, TT>;
            const uint32_t base0 = (uint32_t)__cvta_generic_to_shared(smem_tma);
#pragma unroll
            for (int q = 0; q < 2 * kWarpsPerBlock; ++q) mbar_init(base0 + (q >> 1) * Ring0::kWarpBytes + Ring0::kWarpBarOffset + 8 * (q & 1), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    const Geo& g = A.g;
    const long long task = (long long)blockIdx.x * kWarpsPerBlock + warp;
    TL_STAMP(1, 0);

    // task id -> (strip, first row, end row): equal row ranges (+-1 row), `strips` tasks per range
    // Tasks are handed out from the END of the shard backwards: the pointwise forward is a front-to-back
    // stream, so the last ~100 MB of x and t it read are still in the 126 MB L2 when this kernel starts.
    auto decode = [&](long long tk, int& strip_o, long long& pos_o, long long& end_o) {
        if (A.reverse) tk = g.tasks - 1 - tk;
        const long long tail_tasks = g.tail_groups * g.strips;
        if (tk < tail_tasks) {  // tail phase: short ranges over the first rows of the shard, claimed last
            strip_o = (int)(tk % g.strips);
            const long long grp = tk / g.strips;
            pos_o = grp * g.tail_range;
            end_o = min(pos_o + g.tail_range, g.tail_rows);
            return;
        }
        tk -= tail_tasks;
        strip_o = (int)(tk % g.strips);
        const long long grp = tk / g.strips;
        const long long body = g.total_rows - g.tail_rows;
        pos_o = g.tail_rows + (body * grp) / g.groups;
        end_o = g.tail_rows + (body * (grp + 1)) / g.groups;
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
    if (A.skip_unit_upstream && A.upstream != nullptr && __ldg(A.upstream) == 1.0f) return;  // uniform: plain loss.backward()

    // ---- global sums: given, or (data parallel) collected from the peer mailbox -----------------
    // Every block waits here for all ranks' pointwise sums.  The rows this block starts with were already
    // requested into L2 before pdl_wait(), so their DRAM latency overlaps the wait.  Waiting any later does not
    // pay: the Dice term beta rides in the transposed horizontal stencil from the first residual row on, and
    // the first gradient row leaves two rows (~1 us) after a range starts.  A wait that timed out (a peer died
    // or is out of lock step) zeroes every coefficient: the rank writes a ZERO gradient and a NaN loss, and
    // sets the mailbox status word, instead of poisoning the weights.
    __shared__ double s_gs[PIL_NSUMS];
    const double* gs = A.gsums;
    bool xchg_ok = true;
    if (A.X.world > 0) {
        xchg_ok = xchg_wait_sum(A.X, 0, s_gs);
        gs = s_gs;
    }

    if (task >= g.tasks) {
        if (A.accumulate) {  // idle warp of the last block still takes part in the block reduction
            const double zero[2] = {0.0, 0.0};
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
        if (!xchg_ok) scale = 0.0;
        const double invN = xchg_ok ? 1.0 / (A.n_global > 0 ? (double)A.n_global : gs[7]) : 0.0;
        const double dice_a = xchg_ok ? -2.0 / den : 0.0, dice_b = xchg_ok ? (2.0 * I + s) / (den * den) : 0.0;
        const bool use_rd = A.p.pde_weight > 0.0, use_pf = A.p.phase_field_weight > 0.0;
        c.alpha = (float)(scale * A.p.dice_weight * dice_a);
        c.beta = (float)(scale * A.p.dice_weight * dice_b);
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
        const double Dd = A.p.diffusion_coeff;
        c.nDi = (float)(-1.0 / Dd);
        c.a1D = (float)((1.0 + A.p.reaction_threshold) / Dd);
        c.c0D = (float)((-A.p.reaction_threshold - 4.0 * Dd) / Dd);
        c.cAD = (float)(crd * Dd * Dd);
        c.f3D = (float)(-3.0 * crd * Dd);
        c.f2D = (float)(2.0 * (1.0 + A.p.reaction_threshold) * crd * Dd);
        c.f1cD = (float)((-A.p.reaction_threshold * crd - 4.0 * crd * Dd) * Dd);
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
    const f2 cAf2[2] = {make_float2(c.cAD * fc[0], c.cAD * fc[1]), make_float2(c.cAD * fc[2], c.cAD * fc[3])};  // meet rho = r / D
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
        if (CHECK && k + 1 == H) uc4 = ua;  // row H := row H-2 (src/pde.py:67), which is row k-1
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
        if (CHECK && k + 1 == H) uc4 = ua;  // row H := row H-2 (src/pde.py:67), which is row k-1
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
                // rho = r / D with r = D*(s4 - 4u) + u(1-u)(u-a) as a polynomial in u   (src/pde.py:73-77,:99,:120)
                r[h] = fma2(u[h], fma2(u[h], fma2(bc(c.nDi), u[h], bc(c.a1D)), bc(c.c0D)), s4);
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
            const float cAr = (CHECK && (k == 0 || k == H - 1)) ? 2.0f * c.cAD : c.cAD;
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
                const f2 fpr = fma2(u[h], fma2(bc(c.f3D), u[h], bc(c.f2D)), bc(c.f1cD));  // (cF f'(u) - 4 cA) * D
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
                const f2 uv = fma2(make_float2(-u.x, -u.y), u, u);  // u(1-u) = u - u^2, rounded once
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

    if constexpr (ALIGNED && TMA) {
        // ---- TMA staged path.  Rows travel in BOXES of 6 rows x 128 columns, one box of the map and one of the
        // targets per stage, two stages per warp.  Box n of a segment holds map rows r0+2+6n.. and target rows
        // r0+6n.., i.e. exactly what iterations k = r0+1+6n .. r0+6+6n consume (map row k+1, target row k-1); the
        // two head iterations (k = r0-1, r0) take map rows r0, r0+1 from direct loads issued together with the
        // two halo rows.  One elected lane arms the stage's mbarrier with the byte count and issues the two
        // cp.async.bulk.tensor copies; all lanes wait on the mbarrier's phase and read their 16 bytes per row
        // with LDS.128.  Out-of-tensor parts of a box (column -4 of the first strip, columns >= W, rows past
        // the shard) are zero-filled by the TMA unit and never used; the mirror fix-up stays at consume time.
        using Ring = TmaRing<XT, TT>;
        extern __shared__ __align__(128) unsigned char smem_tma[];
        const uint32_t wbase_s = (uint32_t)__cvta_generic_to_shared(smem_tma) + warp * Ring::kWarpBytes;  // this warp's two stages ...
        const uint32_t bar_s = wbase_s + Ring::kWarpBarOffset;                                                // ... and their two mbarriers
        const int c0 = strip * kStripCols - kVec;  // first column of the warp's 128-column window
        const int cbx = TmaBox<XT>::first_col(c0), cbt = TmaBox<TT>::first_col(c0);  // first columns of the (16-byte aligned) boxes
        const int xoff = (cx.colc - cbx) * (int)sizeof(XT), toff = (cx.colc - cbt) * (int)sizeof(TT) + Ring::kXSlotBytes;
        const int nb = (r1 - r0 - 1) / Ring::kBoxRows + 1;  // boxes this segment consumes
        const int rowt = (int)((long long)b * H) + r0;         // flattened row of target box 0; map boxes start 2 rows later
        const unsigned int q0 = tma_q;
        auto issue_box = [&](int n) {
            if (lane == 0) {
                const uint32_t st = (q0 + (unsigned)n) & 1u;
                const uint32_t dst = wbase_s + st * Ring::kStageBytes, bar = bar_s + 8 * st;
                mbar_expect_tx(bar, Ring::kXBoxBytes + Ring::kTBoxBytes);
                tma_load_2d(dst, tmx, cbx, rowt + 2 + Ring::kBoxRows * n, bar);
                tma_load_2d(dst + Ring::kXSlotBytes, tmt, cbt, rowt + Ring::kBoxRows * n, bar);
            }
        };
        const float4 xh0 = cx.template load<XT>(xrow(r0)), xh1 = cx.template load<XT>(xrow(r0 + 1));
        __syncwarp();  // every lane is done with the previous segment's stages
        issue_box(0);
        if (nb > 1) issue_box(1);
        U0 = act4<KIND>(cx.fix(xa));   // row k0-1
        U1 = act4<KIND>(cx.fix(xbq));  // row k0
        auto rotate = [&]() {
            U0 = U1;
            U1 = U2;
            G0 = G1;
            G1 = G2;
        };
        compute(k0, BoolC<true>{}, xh0, zero4, U0, U1, U2, G0, G1, G2);  // k = r0-1: forms r[r0-1], emits nothing
        rotate();
        compute(k0 + 1, BoolC<true>{}, xh1, zero4, U0, U1, U2, G0, G1, G2);  // k = r0: forms r[r0], emits nothing
        rotate();
        int k = r0 + 1, n = 0;
        const int kclean = (r1 == H) ? r1 - 2 : r1 - 1;  // last iteration that needs no row / segment checks
        // this lane's 4 columns of box row `row` of the stage at shared address sp
        auto rd_x = [&](uint32_t sp, int row) { return lds4s<XT>(sp + xoff + row * Ring::kXRowBytes); };
        auto rd_t = [&](uint32_t sp, int row) { return lds4s<TT>(sp + toff + row * Ring::kTRowBytes); };
#pragma unroll 1
        for (; k + Ring::kBoxRows - 1 <= kclean; k += Ring::kBoxRows, ++n) {
            const unsigned int g_ = q0 + (unsigned)n;
            mbar_wait(bar_s + 8 * (g_ & 1u), (g_ >> 1) & 1u);
            const uint32_t sp = wbase_s + (g_ & 1u) * Ring::kStageBytes;
            compute(k + 0, BoolC<false>{}, rd_x(sp, 0), rd_t(sp, 0), U0, U1, U2, G0, G1, G2);
            compute(k + 1, BoolC<false>{}, rd_x(sp, 1), rd_t(sp, 1), U1, U2, U0, G1, G2, G0);
            compute(k + 2, BoolC<false>{}, rd_x(sp, 2), rd_t(sp, 2), U2, U0, U1, G2, G0, G1);
            compute(k + 3, BoolC<false>{}, rd_x(sp, 3), rd_t(sp, 3), U0, U1, U2, G0, G1, G2);
            compute(k + 4, BoolC<false>{}, rd_x(sp, 4), rd_t(sp, 4), U1, U2, U0, G1, G2, G0);
            compute(k + 5, BoolC<false>{}, rd_x(sp, 5), rd_t(sp, 5), U2, U0, U1, G2, G0, G1);
            __syncwarp();  // all lanes have read the stage: it can be refilled
            if (n + 2 < nb) issue_box(n + 2);
        }
        {
            int row = 0;  // the tail starts on a box boundary; boxes n, n+1 are already in flight
#pragma unroll 1
            for (; k <= r1; ++k) {
                const unsigned int g_ = q0 + (unsigned)n;
                if (row == 0) mbar_wait(bar_s + 8 * (g_ & 1u), (g_ >> 1) & 1u);
                const uint32_t sp = wbase_s + (g_ & 1u) * Ring::kStageBytes;
                compute(k, BoolC<true>{}, rd_x(sp, row), rd_t(sp, row), U0, U1, U2, G0, G1, G2);
                rotate();
                if (++row == Ring::kBoxRows) {
                    row = 0;
                    ++n;
                }
            }
        }
        tma_q = q0 + (unsigned)nb;
    } else if constexpr (ALIGNED) {
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
            tot_r2 += (double)(sr2.x + sr2.y) * (A.p.diffusion_coeff * A.p.diffusion_coeff);  // sum r^2 = D^2 sum rho^2
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
    const double acc[2] = {tot_r2, tot_g2};
    bwd_epilogue(A, acc, gs);
    TL_STAMP(1, 3);
}

template <int KIND, typename XT, typename TT, bool ALIGNED>
__global__ void __launch_bounds__(kThreads, kBwdMinBlocks) pil_bwd_kernel(const BwdArgs A) {
    bwd_body<KIND, XT, TT, ALIGNED, false>(A, nullptr, nullptr);
}
// same kernel with the rows staged by TMA boxes; the two tensor maps (x and t as (B*H) x W matrices) are kernel
// parameters in constant space, where the TMA unit reads them
template <int KIND, typename XT, typename TT>
__global__ void __launch_bounds__(kThreads, kBwdMinBlocks) pil_bwd_kernel_tma(const BwdArgs A, const __grid_constant__ CUtensorMap tmx,
                                                                              const __grid_constant__ CUtensorMap tmt) {
    bwd_body<KIND, XT, TT, true, true>(A, &tmx, &tmt);
}
// Rows per dynamically claimed range of the backward kernel.  PIL_BWD_ROWS forces a value (0 = static
// partition).  Automatic: long enough to amortise the 4 halo rows and the pipeline fill of a range, short enough to
// balance.  Measured on B200, 64x1024^2 / 128x2048^2, with the 10% tail phase below where it applies:
//   fp32 maps: 24 rows 136.0 us / 973 us, 32: 135.1 / 1006, 40: 133.0 / 1012, 44: 137.2 / 1042, 48: 131.9 / 1004,
//              52: 134.0 / 1031, 64: 134.9 / 1040, 96: 135.0, 128: 137.2                                  -> 48
//   bf16 maps (bound by the FMA pipe: the halo rows' arithmetic is what costs): 64 rows 118.7 us / 845 us,
//              96: 120.7 / 834, 112: 116.5 / 826, 128: 116.5 / 815, 144: 117.4 / 836, 192: 131.7 / 837, 256: 118.6 / 820 -> 128
// Shrinking ranges (guided self-scheduling) and claiming one range ahead to prefetch it were measured slower.
// Problems too small for a few waves of such ranges keep the static one-wave partition (32x512^2, 8x1024^2 and
// 16x1024^2: static 38.5 / 36.7 / 63.4 us per step against 38.3-42.5 / 38.0-41.5 / 64.2-70.1 with 12..32-row ranges).
static int bwd_dynamic_rows(long long total_rows, long long strips, long long resident_warps, bool narrow_maps) {
    static int forced = -2;
    if (forced == -2) {
        const char* e = getenv("PIL_BWD_ROWS");
        forced = e ? atoi(e) : -1;
    }
    if (forced >= 0) return (forced > 0 && forced < kMinRows) ? kMinRows : forced;
    auto waves = [&](int rows) { return (double)(((total_rows + rows - 1) / rows) * strips) / (double)resident_warps; };
    if (narrow_maps) {
        if (waves(128) >= 1.5) return 128;
        return waves(64) < 2.5 ? 0 : 64;
    }
    return waves(48) < 2.5 ? 0 : 48;
}
// How the aligned backward stages its rows: 1 = TMA boxes (cp.async.bulk.tensor + mbarrier), 0 = per-lane cp.async.
// pil_set_bwd_staging() / PIL_BWD_STAGE=tma|cpasync override the default.
static bool bwd_want_tma() {
    const int forced = host_state().bwd_stage.load();
    if (forced >= 0) return forced == 1;
    static int env = -1;
    if (env < 0) {
        const char* e = getenv("PIL_BWD_STAGE");
        env = (e && (e[0] == 'c' || e[0] == '0')) ? 0 : ((e && (e[0] == 't' || e[0] == '1')) ? 1 : kBwdTmaDefault);
    }
    return env == 1;
}

template <int KIND, typename XT, typename TT>
static cudaError_t launch_bwd_a(BwdArgs& a, int64_t B, int64_t H, int64_t W, bool aligned, cudaStream_t s, LaunchOut* out) {
    static std::atomic<int> per_sm_cache[3][kMaxDevices];  // per template instantiation x {scalar, cp.async, TMA} kernel x device
    auto go = [&](auto kernel, int smem, int which, const auto&... extra) -> cudaError_t {
        if (smem > 48 * 1024 && per_sm_cache[which][current_device()].load(std::memory_order_relaxed) == 0)
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        const int per_sm = blocks_per_sm_cached(kernel, kThreads, smem, per_sm_cache[which], true);
        const int tune_rps = host_state().tune_bwd_rps.load();
        const int resident = sm_count() * per_sm;
        const long long strips_ = (W + kStripCols - 1) / kStripCols;
        const int dyn_rows = a.task_counter != nullptr
                                 ? (tune_rps > 0 ? tune_rps : bwd_dynamic_rows(B * H, strips_, (long long)resident * kWarpsPerBlock, sizeof(XT) == 2))
                                 : 0;
        if (dyn_rows > 0) {
            // persistent grid of the resident blocks; short ranges claimed dynamically (see the kernel)
            a.g = make_geo(B, H, W, resident, dyn_rows, 1);
            // tail phase: the first tail_pct percent of the shard's rows (the last to be claimed) in ranges of tail_range rows
            static int tail_range = -1, tail_pct = -1;
            if (tail_range < 0) {
                const char* e = getenv("PIL_BWD_TAIL");  // "<rows>,<percent of the shard>"
                int r = kBwdTailRange, pc = kBwdTailPercent;
                if (e) sscanf(e, "%d,%d", &r, &pc);
                tail_pct = pc < 0 ? 0 : (pc > 90 ? 90 : pc);
                tail_range = r < 0 ? 0 : r;
            }
            // (measured at 64x1024^2 / 128x2048^2 with 48-row ranges, two interleaved repetitions: fp32 131.9 us / 1004 us against
            // 135.0-135.2 / 1009-1012 without, 24-row tail ranges 133.0 / 997-1004; bf16 maps, bound by the FMA pipe rather
            // than by the last ranges' memory latency, lose 2 us: fp32 maps only)
            if (tail_range >= kMinRows && tail_pct > 0 && tail_range < dyn_rows && sizeof(XT) == 4) {
                long long tr = (long long)B * H * tail_pct / 100;
                tr -= tr % tail_range;
                const long long body = (long long)B * H - tr;
                if (tr > 0 && body >= dyn_rows) {
                    a.g.tail_rows = tr;
                    a.g.tail_range = tail_range;
                    a.g.tail_groups = tr / tail_range;
                    a.g.groups = (body + dyn_rows - 1) / dyn_rows;
                    a.g.tasks = (a.g.groups + a.g.tail_groups) * a.g.strips;
                }
            }
            const long long need = (a.g.tasks + kWarpsPerBlock - 1) / kWarpsPerBlock;
            out->blocks = (int)(need < resident ? need : resident);
            a.first_dynamic = (long long)out->blocks * kWarpsPerBlock;
        } else {
            a.task_counter = nullptr;
            a.first_dynamic = 0;
            a.g = make_geo(B, H, W, resident, tune_rps, tuning_waves(true, B, H, W, resident));
            out->blocks = (int)((a.g.tasks + kWarpsPerBlock - 1) / kWarpsPerBlock);
        }
        out->rows = dyn_rows > 0 ? dyn_rows : (int)((a.g.total_rows + a.g.groups - 1) / a.g.groups);
        if (a.accumulate && (size_t)out->blocks * 2 * kPartialBytes > out->partials_avail) {
            out->status = PIL_ERR_WORKSPACE;
            return cudaSuccess;
        }
        return launch_pdl(kernel, out->blocks, kThreads, smem, s, a, extra...);
    };
    if (aligned) {
        using Ring = TmaRing<XT, TT>;
        CUtensorMap tmx, tmt;
        if (bwd_want_tma() &&
            make_tensor_map_2d(&tmx, a.x, dtype_code<XT>(), (long long)B * H, W, Ring::kBoxRows, TmaBox<XT>::kCols) &&
            make_tensor_map_2d(&tmt, a.t, dtype_code<TT>(), (long long)B * H, W, Ring::kBoxRows, TmaBox<TT>::kCols)) {
            out->tma = 1;
            return go(pil_bwd_kernel_tma<KIND, XT, TT>, Ring::kSmemBytesB, 2, tmx, tmt);
        }
        return go(pil_bwd_kernel<KIND, XT, TT, true>, kSmemPerBlock, 1);
    }
    return go(pil_bwd_kernel<KIND, XT, TT, false>, 0, 0);
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

// exported to pil_api.cu: one entry per input kind
#if !defined(PIL_KIND) || PIL_KIND == 0
cudaError_t launch_bwd_k0(int x_dtype, int t_dtype, PIL_BWD_ARGS) { return launch_bwd_x<PIL_X_PROB>(x_dtype, t_dtype, PIL_BWD_PASS); }
#endif
#if !defined(PIL_KIND) || PIL_KIND == 1
cudaError_t launch_bwd_k1(int x_dtype, int t_dtype, PIL_BWD_ARGS) { return launch_bwd_x<PIL_X_LOGITS_SIGMOID>(x_dtype, t_dtype, PIL_BWD_PASS); }
#endif
#if !defined(PIL_KIND) || PIL_KIND == 2
cudaError_t launch_bwd_k2(int x_dtype, int t_dtype, PIL_BWD_ARGS) { return launch_bwd_x<PIL_X_LOGITS_TANH>(x_dtype, t_dtype, PIL_BWD_PASS); }
#endif

}  // namespace pil
