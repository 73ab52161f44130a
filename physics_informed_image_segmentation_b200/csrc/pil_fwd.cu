// pil_fwd.cu -- K1, the fused full forward kernel (validation / no-grad path and the parameter-sweep moments),
// and its launcher.  Compiled once per input kind (-DPIL_KIND=0|1|2) so the instantiations build in parallel;
// without PIL_KIND (unity development builds) all three kinds are instantiated here.
#include <type_traits>

#include "pil_fwdrow.cuh"

namespace pil {

constexpr int kMomentsMode = 0;    // MOMENTS instantiation: static ranges + cp.async ring (A/B in DESIGN.md)
constexpr int kFwdTmaDefault = 1;  // staging of the aligned full-forward kernel when nothing is forced (A/B in DESIGN.md)

// One warp = one 120-column strip of a range of rows (see pil_common.cuh); after its first, statically assigned
// range a warp claims further (range, strip) tasks from a global counter (persistent grid, like the backward), because
// equal static shares finish far apart: SMs do not get equal shares of the memory system.  A task's sums are fp32
// per thread, reduced over the warp and added to per-warp DOUBLE accumulators in shared memory, so the result is
// reproducible to fp64 rounding whatever the task-to-warp assignment was.
// TMA (ALIGNED only): rows arrive as 2-D tensor-map boxes of 3 rows x 128 columns per map (cp.async.bulk.tensor, one
// elected lane, mbarrier completion), two stages per warp -- the same 6 KB of shared memory per warp as the cp.async
// ring, so the kernel keeps its 7 blocks per SM.
template <int KIND, typename XT, typename TT, bool ALIGNED, bool MOMENTS, bool TMA>
__device__ __forceinline__ void fwd_body(const FwdArgs& A, [[maybe_unused]] const CUtensorMap* tmx, [[maybe_unused]] const CUtensorMap* tmt) {
    static_assert(ALIGNED || !TMA, "the TMA stage ring needs the aligned layout");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;  // (a shuffled, provably uniform warp index helps K2 but makes ptxas
                                                                 // drop this kernel's 3-row steady loop: 117 -> 138 us)
    const Geo& g = A.g;
    constexpr int NACC = MOMENTS ? 16 : 8;
    using Ring = TmaRing<XT, TT, 3>;
    [[maybe_unused]] unsigned int tma_q = 0;  // TMA: boxes this warp has consumed so far
    __shared__ double s_acc[kWarpsPerBlock][NACC];
    if (lane < NACC) s_acc[warp][lane] = 0.0;
    if constexpr (TMA) {
        extern __shared__ __align__(128) unsigned char smem_tma[];
        if (threadIdx.x == 0) {
            const uint32_t bars = (uint32_t)__cvta_generic_to_shared(smem_tma + Ring::kBarOffset);
#pragma unroll
            for (int q = 0; q < 2 * kWarpsPerBlock; ++q) mbar_init(bars + 8 * q, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    __syncwarp();

    FwdRow<KIND, ALIGNED, MOMENTS> fr;
    fr.D = A.D;
    fr.a = A.a;
    fr.a1 = 1.0f + A.a;
    fr.c0 = -A.a - 4.0f * A.D;
    const int H = g.H, W = g.W;
    const bool out_lane = (lane >= 1) && (lane <= kOutLanes);

    long long task_cur = (long long)blockIdx.x * kWarpsPerBlock + warp;
#pragma unroll 1
    while (task_cur < g.tasks) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) fr.acc[k] = 0.f;
        fr.init_packed();
        const int strip = (int)(task_cur % g.strips);
        const long long grp = task_cur / g.strips;
        long long pos = (g.total_rows * grp) / g.groups;              // flattened image row b*H + r
        const long long end = (g.total_rows * (grp + 1)) / g.groups;
        const int col0 = strip * kStripCols + (lane - 1) * kVec;
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

        if constexpr (ALIGNED && TMA) {
            // ---- TMA staged path: box n of the segment holds map AND target rows r0+1+3n .. r0+3+3n, which iterations
            // i = r0+3n .. r0+2+3n consume (iteration i activates row i+1 and adds the stencil terms of row i).
            extern __shared__ __align__(128) unsigned char smem_tma[];
            const uint32_t wbase_s = (uint32_t)__cvta_generic_to_shared(smem_tma) + warp * 2 * Ring::kStageBytes;
            const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(smem_tma) + Ring::kBarOffset + 16 * warp;
            const int c0 = strip * kStripCols - kVec;
            const int cbx = TmaBox<XT>::first_col(c0), cbt = TmaBox<TT>::first_col(c0);
            const int xoff = (cx.colc - cbx) * (int)sizeof(XT), toff = (cx.colc - cbt) * (int)sizeof(TT) + Ring::kXSlotBytes;
            const int nb = (r1 - r0 - 1) / Ring::kBoxRows + 1;
            const int row0 = (int)((long long)b * H) + r0 + 1;  // flattened row of box 0
            const unsigned int q0 = tma_q;
            auto issue_box = [&](int n) {
                if (lane == 0) {
                    const uint32_t st = (q0 + (unsigned)n) & 1u;
                    const uint32_t dst = wbase_s + st * Ring::kStageBytes, bar = bar_s + 8 * st;
                    mbar_expect_tx(bar, Ring::kXBoxBytes + Ring::kTBoxBytes);
                    tma_load_2d(dst, tmx, cbx, row0 + Ring::kBoxRows * n, bar);
                    tma_load_2d(dst + Ring::kXSlotBytes, tmt, cbt, row0 + Ring::kBoxRows * n, bar);
                }
            };
            const float4 x0 = cx.template load<XT>(xrow(r0 - 1)), x1 = cx.template load<XT>(xrow(r0));
            const float4 t1 = cx.template load_plain<TT>(trow(r0));
            __syncwarp();  // every lane is done with the previous segment's stages
            issue_box(0);
            if (nb > 1) issue_box(1);
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 U0 = fr.point4(cx.fix(x0), zero4, false);  // row r0-1: halo row, contributes no pixel terms
            float4 U1 = fr.point4(cx.fix(x1), t1, true);      // row r0
            float4 U2;
            auto rd_x = [&](uint32_t sp, int row) { return lds4s<XT>(sp + xoff + row * Ring::kXRowBytes); };
            auto rd_t = [&](uint32_t sp, int row) { return lds4s<TT>(sp + toff + row * Ring::kTRowBytes); };
            auto step = [&](int i, auto check, const float4& xn, const float4& tn, const float4& um, const float4& uc, float4& up) {
                constexpr bool CHECK = decltype(check)::value;
                if (CHECK && i + 1 == H) up = um;  // row H := row H-2 (src/pde.py:67), which is row i-1; owns no pixel terms
                else up = fr.point4(cx.fix(xn), tn, !CHECK || (i + 1 < r1));
                fr.stencil4(um, uc, up);
            };
            int i = r0, n = 0;
#pragma unroll 1
            for (; i + Ring::kBoxRows - 1 <= r1 - 2; i += Ring::kBoxRows, ++n) {  // clean: rows i+1 .. i+3 < r1
                const unsigned int g_ = q0 + (unsigned)n;
                mbar_wait(bar_s + 8 * (g_ & 1u), (g_ >> 1) & 1u);
                const uint32_t sp = wbase_s + (g_ & 1u) * Ring::kStageBytes;
                step(i + 0, BoolC<false>{}, rd_x(sp, 0), rd_t(sp, 0), U0, U1, U2);
                step(i + 1, BoolC<false>{}, rd_x(sp, 1), rd_t(sp, 1), U1, U2, U0);
                step(i + 2, BoolC<false>{}, rd_x(sp, 2), rd_t(sp, 2), U2, U0, U1);
                __syncwarp();
                if (n + 2 < nb) issue_box(n + 2);
            }
            {
                int row = 0;
#pragma unroll 1
                for (; i < r1; ++i) {
                    const unsigned int g_ = q0 + (unsigned)n;
                    if (row == 0) mbar_wait(bar_s + 8 * (g_ & 1u), (g_ >> 1) & 1u);
                    const uint32_t sp = wbase_s + (g_ & 1u) * Ring::kStageBytes;
                    step(i, BoolC<true>{}, rd_x(sp, row), rd_t(sp, row), U0, U1, U2);
                    U0 = U1;
                    U1 = U2;
                    if (++row == Ring::kBoxRows) {
                        row = 0;
                        ++n;
                    }
                }
            }
            tma_q = q0 + (unsigned)nb;
        } else if constexpr (ALIGNED) {
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

        // ---- fold the task's fp32 sums into the warp's double accumulators; claim the next task ----
        if constexpr (ALIGNED) {
            fr.fold_packed();
            if (!counted) {
#pragma unroll
                for (int k = 0; k < NACC; ++k) fr.acc[k] = 0.f;
            }
        }
#pragma unroll
        for (int k = 0; k < NACC; ++k) {
            float v = fr.acc[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s_acc[warp][k] += (double)v;
        }
        if (A.task_counter == nullptr) break;
        unsigned int claimed = 0u;
        if (lane == 0) claimed = atomicAdd(A.task_counter, 1u);
        claimed = __shfl_sync(0xffffffffu, claimed, 0);
        task_cur = A.first_dynamic + (long long)claimed;
    }

    // ---- deterministic cross-block reduction of the block's doubles; the last block finalises ----
    __syncthreads();
    double blk = 0.0;
    if (threadIdx.x < NACC) {
#pragma unroll
        for (int w = 0; w < kWarpsPerBlock; ++w) blk += s_acc[w][threadIdx.x];
    }
    double raw[NACC];
    if (!blocks_to_last<kThreads, NACC>(blk, A.partials, A.ticket, raw)) return;
    [[maybe_unused]] __shared__ double s_push[MOMENTS ? 16 : 2];
    if (threadIdx.x == 0) {
        if constexpr (MOMENTS) {  // layout: include/pil.h PIL_NMOMENTS
            double* mo = s_push;
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
#pragma unroll
            for (int k = 0; k < 16; ++k) A.sums[k] = mo[k];
        } else {
            double s[PIL_NSUMS];
            sums_from_raw(raw, A.p.epsilon, (double)g.B * (double)g.H * (double)g.W, s);
#pragma unroll
            for (int k = 0; k < PIL_NSUMS; ++k) A.sums[k] = s[k];
            if (A.loss_out != nullptr) finalize_device(s, s[7], A.p, A.loss_out);
        }
        if (A.task_counter != nullptr) *A.task_counter = 0u;  // every warp has made its last claim
    }
    if constexpr (MOMENTS) {
        if (A.X.world > 0) {  // data-parallel sweep: this shard's 16 sums go to every rank as two 8-double vectors
            __syncthreads();
            xchg_push(A.X, 0, s_push);
            xchg_push(A.X, 1, s_push + 8);
        }
    }
}

template <int KIND, typename XT, typename TT, bool ALIGNED, bool MOMENTS>
__global__ void __launch_bounds__(kThreads, MOMENTS ? kFwdMinBlocks - 1 : kFwdMinBlocks) pil_fwd_kernel(const FwdArgs A) {
    fwd_body<KIND, XT, TT, ALIGNED, MOMENTS, false>(A, nullptr, nullptr);
}
// the same kernel with its rows staged by TMA boxes; the tensor maps of x and t are kernel parameters in constant space
template <int KIND, typename XT, typename TT, bool MOMENTS>
__global__ void __launch_bounds__(kThreads, MOMENTS ? kFwdMinBlocks - 1 : kFwdMinBlocks)
    pil_fwd_kernel_tma(const FwdArgs A, const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmt) {
    fwd_body<KIND, XT, TT, true, MOMENTS, true>(A, &tmx, &tmt);
}

// rows per dynamically claimed range (0 = one static range per warp): like the backward kernel, 64 rows when the shard
// gives every resident warp at least ~2.5 such tasks.  PIL_FWD_ROWS forces a value.
static int fwd_dynamic_rows(long long total_rows, long long strips, long long resident_warps) {
    static int forced = -2;
    if (forced == -2) {
        const char* e = getenv("PIL_FWD_ROWS");
        forced = e ? atoi(e) : -1;
    }
    if (forced >= 0) return (forced > 0 && forced < kMinRows) ? kMinRows : forced;
    const int rows = 64;
    const double waves = (double)(((total_rows + rows - 1) / rows) * strips) / (double)resident_warps;
    return waves < 2.0 ? 0 : rows;  // 7 blocks per SM: 64x1024^2 is 2.2 waves, measured faster dynamic (111.8 against 116.9 us)
}
static bool fwd_want_tma() {
    const int forced = host_state().bwd_stage.load();  // pil_set_bwd_staging covers both stencil kernels
    if (forced >= 0) return forced == 1;
    static int env = -1;
    if (env < 0) {
        const char* e = getenv("PIL_FWD_STAGE");
        env = (e && (e[0] == 'c' || e[0] == '0')) ? 0 : ((e && (e[0] == 't' || e[0] == '1')) ? 1 : kFwdTmaDefault);
    }
    return env == 1;
}

template <int KIND, typename XT, typename TT>
static cudaError_t launch_fwd_a(FwdArgs& a, int64_t B, int64_t H, int64_t W, bool aligned, size_t partial_bytes_avail,
                                cudaStream_t s, LaunchOut* out, bool moments) {
    static std::atomic<int> per_sm_cache[6][kMaxDevices];  // per template instantiation x {scalar, cp.async, TMA} x {sums, moments} x device
    auto go = [&](auto kernel, int smem, int which, const auto&... extra) -> cudaError_t {
        const int per_sm = blocks_per_sm_cached(kernel, kThreads, smem, per_sm_cache[which + (moments ? 3 : 0)], true);
        const int resident = sm_count() * per_sm;
        const int tune_rps = host_state().tune_fwd_rps.load();
        const long long strips_ = (W + kStripCols - 1) / kStripCols;
        // measured on B200 at 64x1024^2: the sums kernel gains 8% from TMA boxes + dynamic ranges; the MOMENTS
        // instantiation (16 accumulators to fold per task, 6 blocks per SM) is faster with one static range per warp
        static int mom_mode = -1;  // PIL_MOM_MODE: bit 0 = dynamic ranges, bit 1 = TMA boxes for the MOMENTS instantiation (experiments)
        if (mom_mode < 0) {
            const char* e = getenv("PIL_MOM_MODE");
            mom_mode = e ? atoi(e) : kMomentsMode;
        }
        const int dyn_rows = (tune_rps > 0 || (moments && !(mom_mode & 1))) ? 0 : fwd_dynamic_rows(B * H, strips_, (long long)resident * kWarpsPerBlock);
        if (dyn_rows > 0 && out->task_counter != nullptr) {
            a.g = make_geo(B, H, W, resident, dyn_rows, 1);
            const long long need = (a.g.tasks + kWarpsPerBlock - 1) / kWarpsPerBlock;
            out->blocks = (int)(need < resident ? need : resident);
            a.task_counter = out->task_counter;
            a.first_dynamic = (long long)out->blocks * kWarpsPerBlock;
        } else {
            a.task_counter = nullptr;
            a.first_dynamic = 0;
            a.g = make_geo(B, H, W, resident, tune_rps, tuning_waves(false, B, H, W, resident));
            out->blocks = (int)((a.g.tasks + kWarpsPerBlock - 1) / kWarpsPerBlock);
        }
        out->rows = (int)((a.g.total_rows + a.g.groups - 1) / a.g.groups);
        if ((size_t)out->blocks * (moments ? 16 : PIL_NSUMS) * kPartialBytes > partial_bytes_avail) {
            out->status = PIL_ERR_WORKSPACE;
            return cudaSuccess;
        }
        kernel<<<out->blocks, kThreads, smem, s>>>(a, extra...);
        return cudaGetLastError();
    };
    if (aligned) {
        using Ring = TmaRing<XT, TT, 3>;
        CUtensorMap tmx, tmt;
        static int mom_mode2 = -1;
        if (mom_mode2 < 0) {
            const char* e = getenv("PIL_MOM_MODE");
            mom_mode2 = e ? atoi(e) : kMomentsMode;
        }
        if (fwd_want_tma() && (!moments || (mom_mode2 & 2)) &&
            make_tensor_map_2d(&tmx, a.x, dtype_code<XT>(), (long long)B * H, W, Ring::kBoxRows, TmaBox<XT>::kCols) &&
            make_tensor_map_2d(&tmt, a.t, dtype_code<TT>(), (long long)B * H, W, Ring::kBoxRows, TmaBox<TT>::kCols)) {
            out->tma = 1;
            if (moments) return go(pil_fwd_kernel_tma<KIND, XT, TT, true>, Ring::kSmemBytes, 2, tmx, tmt);
            return go(pil_fwd_kernel_tma<KIND, XT, TT, false>, Ring::kSmemBytes, 2, tmx, tmt);
        }
        if (moments) return go(pil_fwd_kernel<KIND, XT, TT, true, true>, kSmemPerBlock, 1);
        return go(pil_fwd_kernel<KIND, XT, TT, true, false>, kSmemPerBlock, 1);
    }
    if (moments) return go(pil_fwd_kernel<KIND, XT, TT, false, true>, 0, 0);
    return go(pil_fwd_kernel<KIND, XT, TT, false, false>, 0, 0);
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

// exported to pil_api.cu: one entry per input kind (each lives in its own translation unit in release builds)
#if !defined(PIL_KIND) || PIL_KIND == 0
cudaError_t launch_fwd_k0(int x_dtype, int t_dtype, PIL_FWD_ARGS) { return launch_fwd_x<PIL_X_PROB>(x_dtype, t_dtype, PIL_FWD_PASS); }
#endif
#if !defined(PIL_KIND) || PIL_KIND == 1
cudaError_t launch_fwd_k1(int x_dtype, int t_dtype, PIL_FWD_ARGS) { return launch_fwd_x<PIL_X_LOGITS_SIGMOID>(x_dtype, t_dtype, PIL_FWD_PASS); }
#endif
#if !defined(PIL_KIND) || PIL_KIND == 2
cudaError_t launch_fwd_k2(int x_dtype, int t_dtype, PIL_FWD_ARGS) { return launch_fwd_x<PIL_X_LOGITS_TANH>(x_dtype, t_dtype, PIL_FWD_PASS); }
#endif

}  // namespace pil
