// pil_api.cu -- host side of the C ABI of include/pil.h: argument checks, workspace layout, the small kernels
// (finalize, sweep finalize, per-image metrics, gradient scaling, stand-alone PDE operators) and the exported
// entry points.  The fused kernels live in pil_fwd.cu / pil_point.cu / pil_bwd.cu.
#include "pil_common.cuh"

namespace pil {
// per-image Dice and IoU from the counts (src/metrics.py:66-70, src/evaluate.py:90-94)
__global__ void pil_image_metrics_kernel(const double* counts, long long B, double smooth, float* dice, float* iou) {
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
        const double I = counts[4 * b], P = counts[4 * b + 1], T = counts[4 * b + 2];
        if (dice) dice[b] = (float)((2.0 * I + smooth) / (P + T + smooth));
        if (iou) iou[b] = (float)((I + smooth) / (P + T - I + smooth));
    }
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

// the same from the mailbox: the ranks' 16 moment sums arrive as two 8-double vectors (phases 0 and 1, pushed by the last
// block of pil_forward_moments_xchg); the block adds them in rank order and evaluates the settings on the GLOBAL moments
__global__ void __launch_bounds__(kThreads) pil_sweep_finalize_xchg_kernel(XchgDev X, long long n_global, SweepParams sp, float* out,
                                                                           double* moments_out) {
    __shared__ double s_mo[16];
    const bool ok0 = xchg_wait_sum(X, 0, s_mo);
    const bool ok1 = xchg_wait_sum(X, 1, s_mo + 8);
    (void)ok0;
    (void)ok1;  // a timed-out wait leaves NaN moments: every loss of the sweep is NaN and the status word is set
    if (threadIdx.x == 0) xchg_advance_epoch(X);  // device-epoch mode: this rank has completed the sweep step
    if (moments_out != nullptr && threadIdx.x < 16) moments_out[threadIdx.x] = s_mo[threadIdx.x];
    const int k = threadIdx.x;
    if (k >= sp.n) return;
    const PilParams& p = sp.p[k];
    const double D = p.diffusion_coeff, a = p.reaction_threshold, eps = p.epsilon;
    const double* mo = s_mo;
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

// stand-alone push of a sums vector into every rank's mailbox (pil_exchange_push): for callers that assemble the
// shard's pointwise sums from several launches (the host-buffer session adds per-chunk sums first)
__global__ void __launch_bounds__(kThreads) pil_xchg_push_kernel(XchgDev X, int phase, const double* sums) {
    __shared__ double s_v[PIL_NSUMS];
    if (threadIdx.x < PIL_NSUMS) s_v[threadIdx.x] = sums[threadIdx.x];
    __syncthreads();
    xchg_push(X, phase, s_v);
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

thread_local PilLaunchInfo t_info = {};
HostState& host_state() {
    static HostState hs;
    return hs;
}
static inline void count_launch() { host_state().kernels_launched.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
    static std::atomic<int> cache[kMaxDevices];
    const int dev = current_device();
    int n = cache[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}
// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (libpil.so links cudart only)
bool make_tensor_map_2d(CUtensorMap* out, const void* base, int dtype, long long rows, long long cols, int box_rows, int box_cols) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static std::atomic<EncodeFn> cached{nullptr};
    static std::atomic<int> tried{0};
    EncodeFn fn = cached.load(std::memory_order_acquire);
    if (fn == nullptr) {
        if (tried.load(std::memory_order_acquire)) return false;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess && p != nullptr) {
            fn = reinterpret_cast<EncodeFn>(p);
            cached.store(fn, std::memory_order_release);
        }
        tried.store(1, std::memory_order_release);
        if (fn == nullptr) return false;
    }
    const size_t esz = dtype_size(dtype);
    if (((uintptr_t)base % 16) || ((size_t)cols * esz) % 16 || rows < 1 || cols < 1 || rows >= (1ll << 31) - 64) return false;
    const CUtensorMapDataType dt = dtype == PIL_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                                    : (dtype == PIL_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8);
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * esz};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

constexpr long long kL2ResidentMB = 110;  // maps + gradient up to this size are left L2-resident by the forward (L2: 126 MB)

static unsigned long long xchg_timeout_ns() {
    static unsigned long long v = 0;
    if (v == 0) {
        const char* e = getenv("PIL_XCHG_TIMEOUT_MS");
        const long long ms = (e && atoll(e) > 0) ? atoll(e) : 20000;
        v = (unsigned long long)ms * 1000000ull;
    }
    return v;
}
int make_xchg(const PilExchange* ex, XchgDev* X) {
    *X = XchgDev{};
    if (!ex) return PIL_OK;
    if (ex->world < 1 || ex->world > PIL_MAX_RANKS || ex->rank < 0 || ex->rank >= ex->world) return PIL_ERR_EXCHANGE;
    X->rank = ex->rank;
    X->world = ex->world;
    X->parity = (int)(ex->epoch & 1ull);
    X->defer = (ex->flags & PIL_XCHG_DEFER_FINALIZE) ? 1 : 0;
    X->device_epoch = (ex->flags & PIL_XCHG_DEVICE_EPOCH) ? 1 : 0;
    if (X->defer && X->device_epoch) return PIL_ERR_EXCHANGE;  // the deferred finalize kernel takes its epoch from the host
    X->want = (ex->epoch % 0xfffffffeull) + 1ull;  // 32-bit step tag, never 0 (mailboxes start zeroed)
    X->timeout_ns = xchg_timeout_ns();
    for (int r = 0; r < ex->world; ++r) {
        if (!ex->mailbox[r]) return PIL_ERR_EXCHANGE;
        X->box[r] = reinterpret_cast<unsigned char*>(ex->mailbox[r]);
    }
    return PIL_OK;
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
    l.total = l.partials_off + (size_t)(blocks > kMaxPointBlocks ? blocks : kMaxPointBlocks) * PIL_NMOMENTS * kPartialBytes;
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
    // header AND partials: a partial slot is valid when it carries the tag of the current launch (pil_common.cuh), so no
    // slot may start out with bytes that could pass for one
    return (int)cudaMemsetAsync(workspace, 0, workspace_bytes, (cudaStream_t)stream);
}

static int forward_impl(const void* x, const void* t, int64_t B, int64_t H, int64_t W, int x_dtype, int t_dtype, int x_kind,
                        const PilParams* p, double* sums, float* loss_out, void* workspace, size_t workspace_bytes,
                        void* stream, bool moments, const PilExchange* ex = nullptr) {
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
    st = make_xchg(ex, &a.X);
    if (st != PIL_OK) return st;
    PIL_SET_BOUNDS(x, t, nullptr, B * H * W, x_dtype, t_dtype);
    const bool aligned = is_aligned_case(x, t, nullptr, W, x_dtype, t_dtype);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t avail = workspace_bytes - wl.partials_off;
    LaunchOut lo;
    lo.task_counter = a.ticket + 1;
    a.task_counter = nullptr;
    a.first_dynamic = 0;
    cudaError_t e;
    switch (x_kind) {
        case PIL_X_PROB: e = launch_fwd_k0(x_dtype, t_dtype, a, B, H, W, aligned, avail, s, &lo, moments); break;
        case PIL_X_LOGITS_SIGMOID: e = launch_fwd_k1(x_dtype, t_dtype, a, B, H, W, aligned, avail, s, &lo, moments); break;
        default: e = launch_fwd_k2(x_dtype, t_dtype, a, B, H, W, aligned, avail, s, &lo, moments); break;
    }
    if (lo.status != PIL_OK) return lo.status;
    const int blocks = lo.blocks;
    t_info.fwd_blocks = blocks;
    t_info.fwd_threads = kThreads;
    t_info.fwd_rows_per_segment = lo.rows;
    t_info.fwd_aligned = aligned ? (lo.tma ? 2 : 1) : 0;
    count_launch();
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

int pil_forward_moments_xchg(const void* x, const void* t, int64_t B, int64_t H, int64_t W, int x_dtype, int t_dtype, int x_kind,
                             double* moments, void* workspace, size_t workspace_bytes, const PilExchange* ex, void* stream) {
    if (!ex) return PIL_ERR_NULL;
    PilParams neutral = {0.5, 0.5, 0.0, 0.0, 1.0, 0.5, 1.0, 1e-6};
    return forward_impl(x, t, B, H, W, x_dtype, t_dtype, x_kind, &neutral, moments, nullptr, workspace, workspace_bytes, stream, true, ex);
}

int pil_sweep_finalize_xchg(const PilExchange* ex, int64_t n_global, const PilParams* params, int n_params, float* loss_out,
                            double* moments_out, void* stream) {
    if (!ex || !params || !loss_out) return PIL_ERR_NULL;
    if (n_params < 1 || n_params > kSweepChunk) return PIL_ERR_SHAPE;
    for (int k = 0; k < n_params; ++k) {
        const int st = pil_validate_params(params + k);
        if (st != PIL_OK) return st;
    }
    XchgDev X;
    int st = make_xchg(ex, &X);
    if (st != PIL_OK) return st;
    SweepParams sp;
    sp.n = n_params;
    for (int k = 0; k < n_params; ++k) sp.p[k] = params[k];
    pil_sweep_finalize_xchg_kernel<<<1, kThreads, 0, (cudaStream_t)stream>>>(X, (long long)n_global, sp, loss_out, moments_out);
    count_launch();
    return (int)cudaGetLastError();
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
        count_launch();
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    return PIL_OK;
}

int pil_finalize(const double* sums, int64_t n_global, const PilParams* p, float* loss_out, void* stream) {
    if (!sums || !p || !loss_out) return PIL_ERR_NULL;
    pil_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, (long long)n_global, *p, loss_out);
    count_launch();
    return (int)cudaGetLastError();
}

static int backward_impl(const void* x, const void* t, void* grad, int64_t B, int64_t H, int64_t W, int x_dtype, int t_dtype,
                         int x_kind, const PilParams* p, const double* global_sums, int64_t n_global, const float* upstream,
                         float grad_scale, double* stencil_sums, float* loss_out, double* total_sums, void* acc_ws,
                         size_t acc_ws_bytes, void* stream, const PilExchange* ex = nullptr, bool skip_unit_upstream = false) {
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
    a.skip_unit_upstream = skip_unit_upstream ? 1 : 0;
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
    if (a.accumulate) lo.partials_avail = acc_ws_bytes - workspace_layout(B, H, W).partials_off;
    cudaError_t e;
    switch (x_kind) {
        case PIL_X_PROB: e = launch_bwd_k0(x_dtype, t_dtype, a, B, H, W, aligned, s, &lo); break;
        case PIL_X_LOGITS_SIGMOID: e = launch_bwd_k1(x_dtype, t_dtype, a, B, H, W, aligned, s, &lo); break;
        default: e = launch_bwd_k2(x_dtype, t_dtype, a, B, H, W, aligned, s, &lo); break;
    }
    if (lo.status != PIL_OK) return lo.status;
    const int blocks = lo.blocks;
    t_info.bwd_blocks = blocks;
    t_info.bwd_threads = kThreads;
    t_info.bwd_rows_per_segment = lo.rows;
    t_info.bwd_aligned = aligned ? (lo.tma ? 2 : 1) : 0;
    count_launch();
    return (int)e;
}

int pil_backward(const void* x, const void* t, void* grad, int64_t B, int64_t H, int64_t W, int x_dtype, int t_dtype,
                 int x_kind, const PilParams* p, const double* global_sums, int64_t n_global, const float* upstream,
                 float grad_scale, void* stream) {
    return backward_impl(x, t, grad, B, H, W, x_dtype, t_dtype, x_kind, p, global_sums, n_global, upstream, grad_scale,
                         nullptr, nullptr, nullptr, nullptr, 0, stream);
}

int pil_backward_if_scaled(const void* x, const void* t, void* grad, int64_t B, int64_t H, int64_t W, int x_dtype, int t_dtype,
                           int x_kind, const PilParams* p, const double* global_sums, int64_t n_global, const float* upstream,
                           float grad_scale, void* stream) {
    if (!upstream) return PIL_ERR_NULL;
    return backward_impl(x, t, grad, B, H, W, x_dtype, t_dtype, x_kind, p, global_sums, n_global, upstream, grad_scale,
                         nullptr, nullptr, nullptr, nullptr, 0, stream, nullptr, true);
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
                         stencil_sums, loss_out, nullptr, workspace, workspace_bytes, stream);
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
        long long keep_mb = host_state().l2_keep_mb.load();
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
        // Small shards (a data-parallel rank of a strong-scaled batch, the reference's own 8x128x128 batches): when
        // both maps and the gradient fit in the 126 MB L2 the stream is not "read once" at all -- the backward finds
        // x and t in L2 if this kernel does not mark them evict_first.  Plain loads then.
        static long long resident_mb = -1;
        if (resident_mb < 0) {
            const char* e = getenv("PIL_L2_RESIDENT_MB");
            resident_mb = e ? atoll(e) : kL2ResidentMB;
        }
        const long long footprint = a.n * (long long)(2 * dtype_size(x_dtype) + dtype_size(t_dtype));
        if (footprint <= (resident_mb << 20)) a.keep_from4 = (a.n >> 2);
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
        case PIL_X_PROB: e = launch_point_k0(x_dtype, t_dtype, a, aligned, s, &blocks); break;
        case PIL_X_LOGITS_SIGMOID: e = launch_point_k1(x_dtype, t_dtype, a, aligned, s, &blocks); break;
        default: e = launch_point_k2(x_dtype, t_dtype, a, aligned, s, &blocks); break;
    }
    t_info.fwd_blocks = blocks;
    t_info.fwd_threads = kPointThreads;
    t_info.fwd_rows_per_segment = 0;
    t_info.fwd_aligned = aligned ? 1 : 0;
    count_launch();
    return (int)e;
}

int pil_forward_pointwise(const void* x, const void* t, int64_t B, int64_t H, int64_t W, int x_dtype, int t_dtype,
                          int x_kind, const PilParams* p, double* sums, void* workspace, size_t workspace_bytes,
                          void* stream) {
    return pointwise_impl(x, t, B, H, W, x_dtype, t_dtype, x_kind, p, sums, workspace, workspace_bytes, stream, nullptr);
}

// ---- data-parallel training step over the peer-memory exchange (no NCCL call, 2 launches) --------
size_t pil_exchange_bytes(void) { return (size_t)kXchgStatusOffset + 128; }  // slots, status word, device epoch counter

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
                         stencil_sums, loss_out, total_sums, workspace, workspace_bytes, stream, ex);
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
    a.l2_stream = (host_state().l2_keep_mb.load() != 0 &&
                   a.n * (long long)(2 * dtype_size(x_dtype) + dtype_size(t_dtype)) > ((long long)kL2ResidentMB << 20)) ? 1 : 0;
    st = make_xchg(ex, &a.X);
    if (st != PIL_OK) return st;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(image_counts, 0, (size_t)B * 4 * sizeof(double), s);
    if (e != cudaSuccess) return (int)e;
    PIL_SET_BOUNDS(x, t, nullptr, B * H * W, x_dtype, t_dtype);
    const bool aligned = (a.hw % 4 == 0) && is_aligned_case(x, t, nullptr, 4, x_dtype, t_dtype);
    int blocks = 0;
    switch (x_kind) {
        case PIL_X_PROB: e = launch_point_metrics_k0(x_dtype, t_dtype, a, aligned, s, &blocks); break;
        case PIL_X_LOGITS_SIGMOID: e = launch_point_metrics_k1(x_dtype, t_dtype, a, aligned, s, &blocks); break;
        default: e = launch_point_metrics_k2(x_dtype, t_dtype, a, aligned, s, &blocks); break;
    }
    t_info.fwd_blocks = blocks;
    t_info.fwd_threads = kPointThreads;
    t_info.fwd_rows_per_segment = 0;
    t_info.fwd_aligned = aligned ? 1 : 0;
    count_launch();
    return (int)e;
}

int pil_image_metrics(const double* image_counts, int64_t B, double smooth, float* dice_out, float* iou_out, void* stream) {
    if (!image_counts || (!dice_out && !iou_out)) return PIL_ERR_NULL;
    if (B < 1) return PIL_ERR_SHAPE;
    const int blocks = (int)((B + 127) / 128 < 64 ? (B + 127) / 128 : 64);
    pil_image_metrics_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(image_counts, (long long)B, smooth, dice_out, iou_out);
    count_launch();
    return (int)cudaGetLastError();
}

int pil_exchange_finalize(const PilExchange* ex, int64_t n_global, const PilParams* p, float* loss_out, double* total_sums,
                          void* stream) {
    if (!ex || !p || (!loss_out && !total_sums)) return PIL_ERR_NULL;
    XchgDev X;
    int st = make_xchg(ex, &X);
    if (st != PIL_OK) return st;
    pil_xchg_finalize_kernel<<<1, kThreads, 0, (cudaStream_t)stream>>>(X, (long long)n_global, *p, loss_out, total_sums);
    count_launch();
    return (int)cudaGetLastError();
}

int pil_exchange_push(const PilExchange* ex, int phase, const double* sums, void* stream) {
    if (!ex || !sums) return PIL_ERR_NULL;
    if (phase != 0 && phase != 1) return PIL_ERR_EXCHANGE;
    XchgDev X;
    int st = make_xchg(ex, &X);
    if (st != PIL_OK) return st;
    pil_xchg_push_kernel<<<1, kThreads, 0, (cudaStream_t)stream>>>(X, phase, sums);
    count_launch();
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
                         sums, workspace, workspace_bytes, stream);
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
    count_launch();
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
    count_launch();
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
    count_launch();
    return (int)cudaGetLastError();
}

int pil_last_launch_info(PilLaunchInfo* out) {
    if (!out) return PIL_ERR_NULL;
    *out = t_info;
    out->kernels_launched = host_state().kernels_launched.load();
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

int pil_set_bwd_staging(int mode) {
    host_state().bwd_stage.store(mode < 0 ? -1 : (mode ? 1 : 0));
    return PIL_OK;
}

int pil_set_l2_keep_mb(int mb) {
    host_state().l2_keep_mb.store(mb);
    return PIL_OK;
}

int pil_set_tuning(int fwd_rows_per_segment, int bwd_rows_per_segment) {
    host_state().tune_fwd_rps.store(fwd_rows_per_segment > 0 ? fwd_rows_per_segment : 0);
    host_state().tune_bwd_rps.store(bwd_rows_per_segment > 0 ? bwd_rows_per_segment : 0);
    return PIL_OK;
}

}  // extern "C"
