// pil_boundary.cu -- the per-step boundary-F1 metric of the reference on the GPU (pil_boundary_* of include/pil.h).
//
// Reference: src/evaluate.py:102-229 (extract_boundaries, compute_boundary_f1, compute_boundary_f1_batch), called on
// every training and validation step (src/train.py:156, :259).  There, every image of the batch is pulled to the
// host (.cpu().numpy()), thresholded, handed to OpenCV (findContours RETR_EXTERNAL + drawContours, two
// distanceTransform DIST_L2 / mask 5) and reduced with NumPy -- B host round trips per step.  Here the same integer
// quantities are counted on the device, with no host synchronisation:
//
//   boundary(M) = drawContours(findContours(M, RETR_EXTERNAL, CHAIN_APPROX_NONE), thickness 1)
//               = the foreground pixels that are 4-adjacent to the OUTSIDE background, i.e. to a background pixel that is
//                 4-connected to the image frame (pixels beyond the image count as frame).  Suzuki-Abe border following
//                 with 8-connected foreground visits exactly the foreground pixels that have a 0-pixel of the surrounding
//                 background component in their 4-neighbourhood; RETR_EXTERNAL keeps the outer borders of the components
//                 the frame's background touches -- holes, and anything inside a hole, contribute nothing.
//   distanceTransform(1 - boundary, DIST_L2, 5) <= tol
//               = some boundary pixel lies at a 5x5-chamfer distance <= tol: OpenCV's fixed-point weights
//                 a = 65536 (1.0), b = 91750 (1.4), c = 143976 (2.1969) on (1,0), (1,1), (2,1) steps.  For tol = 2 that is
//                 the 13-pixel neighbourhood {(0,0), (+-1,0), (0,+-1), (+-1,+-1), (+-2,0), (0,+-2)}.
//
// Frame-connected background = connected components of the background (4-connectivity) by lock-free union-find
// (label = smallest pixel index of the component), one plane for the thresholded prediction and one for the target,
// then a flag per root that an edge pixel belongs to it.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "pil.h"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxOffsets = 192;  // (2*6+1)^2 = 169 for the largest supported tolerance

__device__ __forceinline__ float load_f(const void* p, int dtype, long long i) {
    if (dtype == PIL_F32) return reinterpret_cast<const float*>(p)[i];
    if (dtype == PIL_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
    return (float)reinterpret_cast<const uint8_t*>(p)[i];
}

struct BArgs {
    const void* x;
    const void* t;
    long long n;       // B*H*W
    int H, W;
    int x_dtype, t_dtype, x_kind;
    float threshold;
    unsigned char* fg;       // [n] bit0: thresholded prediction, bit1: target mask
    int* label;              // [2][n] union-find parent of background pixels (global pixel index), -1 for foreground
    unsigned char* outside;  // [2][n] indexed by ROOT: 1 if the component touches the image frame
    unsigned char* bnd;      // [n] bit0: prediction boundary, bit1: target boundary
    long long* counts;       // [B][4]: |Bp|, |Bt|, |{p in Bp near Bt}|, |{q in Bt near Bp}|   (tolerance 0: [2] = |Bp & Bt|)
};

// threshold both maps (src/evaluate.py:146 `predictions > threshold`; extract_boundaries :113 casts mask*255 to uint8 and
// findContours treats non-zero as foreground) and initialise the union-find forests.  A warp covers 32 consecutive pixels;
// a background pixel starts out labelled with the FIRST pixel of its horizontal run inside the warp's chunk (one ballot),
// so long runs arrive pre-merged and the merge kernel only has to stitch chunk boundaries and vertically adjacent runs.
__global__ void __launch_bounds__(kThreads) bf1_init(const BArgs A) {
    const int lane = threadIdx.x & 31;
    const long long n32 = (A.n + 31) / 32 * 32;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n32; i += (long long)gridDim.x * kThreads) {
        const bool valid = i < A.n;
        bool fp = true, ft = true;
        int col = 0;
        if (valid) {
            float u = load_f(A.x, A.x_dtype, i);
            if (A.x_kind != PIL_X_PROB) {
                const float s = (A.x_kind == PIL_X_LOGITS_TANH) ? 2.0f : 1.0f;
                u = 1.0f / (1.0f + expf(-s * u));  // src/unet.py:208-214
            }
            fp = u > A.threshold;
            const float tv = load_f(A.t, A.t_dtype, i) * 255.0f;
            ft = ((unsigned char)(int)tv) != 0;  // (mask * 255).astype(np.uint8) != 0
            col = (int)(i % A.W);
        }
        // "continues the run of the pixel to its left" (same row, same warp chunk), per plane
        const unsigned bgp = __ballot_sync(0xffffffffu, valid && !fp), bgt = __ballot_sync(0xffffffffu, valid && !ft);
        const unsigned row_start = __ballot_sync(0xffffffffu, col == 0);
        const unsigned contp = bgp & (bgp << 1) & ~row_start, contt = bgt & (bgt << 1) & ~row_start;
        if (valid) {
            // number of consecutive "continues" bits ending at this lane = distance to the run's first pixel in the chunk
            const int kp = __clz(~(contp << (31 - lane))), kt = __clz(~(contt << (31 - lane)));
            A.fg[i] = (unsigned char)((fp ? 1 : 0) | (ft ? 2 : 0));
            A.label[i] = fp ? -1 : (int)(i - kp);
            A.label[A.n + i] = ft ? -1 : (int)(i - kt);
            A.outside[i] = 0;
            A.outside[A.n + i] = 0;
            A.bnd[i] = 0;
        }
    }
}

__device__ __forceinline__ int uf_find(const int* L, int a) {
    int p = L[a];
    while (p != a) {
        a = p;
        p = L[a];
    }
    return a;
}
// lock-free union: the larger root is hooked under the smaller one
__device__ __forceinline__ void uf_union(int* L, int a, int b) {
    for (;;) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a < b) {
            const int s = a;
            a = b;
            b = s;
        }
        const int old = atomicMin(&L[a], b);
        if (old == a) return;
        a = old;
    }
}

// background 4-connectivity inside every image.  Runs are pre-merged inside 32-pixel chunks (bf1_init), so a pixel only
// merges with its LEFT neighbour across a chunk boundary, and with its UPPER neighbour where the pair (pixel, upper pixel)
// does not simply continue the pair to its left -- one union per pair of vertically touching runs instead of one per pixel.
__global__ void __launch_bounds__(kThreads) bf1_merge(const BArgs A) {
    const long long total = 2 * A.n;
    for (long long q = (long long)blockIdx.x * kThreads + threadIdx.x; q < total; q += (long long)gridDim.x * kThreads) {
        const int plane = q >= A.n;
        const long long i = q - (plane ? A.n : 0);
        int* L = A.label + (plane ? A.n : 0);
        if (L[i] < 0) continue;
        const int col = (int)(i % A.W);
        const int row = (int)((i / A.W) % A.H);
        const bool left = col > 0 && L[i - 1] >= 0;
        if (left && (i & 31) == 0) uf_union(L, (int)i, (int)(i - 1));
        if (row > 0 && L[i - A.W] >= 0) {
            const bool same_pair_as_left = left && L[i - A.W - 1] >= 0;
            if (!same_pair_as_left) uf_union(L, (int)i, (int)(i - A.W));
        }
    }
}

// path compression + the frame flag of every component that owns a pixel of the image's first/last row or column
__global__ void __launch_bounds__(kThreads) bf1_flatten_mark(const BArgs A) {
    const long long total = 2 * A.n;
    for (long long q = (long long)blockIdx.x * kThreads + threadIdx.x; q < total; q += (long long)gridDim.x * kThreads) {
        const int plane = q >= A.n;
        const long long i = q - (plane ? A.n : 0);
        int* L = A.label + (plane ? A.n : 0);
        if (L[i] < 0) continue;
        const int root = uf_find(L, (int)i);
        L[i] = root;  // benign race: every value written on the way is an ancestor
        const int col = (int)(i % A.W);
        const int row = (int)((i / A.W) % A.H);
        if (col == 0 || col == A.W - 1 || row == 0 || row == A.H - 1) A.outside[(plane ? A.n : 0) + root] = 1;
    }
}

// boundary pixels (see the header comment) and their number per image
__global__ void __launch_bounds__(kThreads) bf1_boundary(const BArgs A) {
    const long long hw = (long long)A.H * A.W;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < A.n; i += (long long)gridDim.x * kThreads) {
        const unsigned char f = A.fg[i];
        if (f == 0) continue;
        const int col = (int)(i % A.W);
        const int row = (int)((i / A.W) % A.H);
        unsigned char b = 0;
#pragma unroll
        for (int plane = 0; plane < 2; ++plane) {
            if (!(f & (1 << plane))) continue;
            const int* L = A.label + (plane ? A.n : 0);
            const unsigned char* out = A.outside + (plane ? A.n : 0);
            auto open = [&](bool inside, long long j) -> bool {
                if (!inside) return true;  // beyond the image: the frame
                const int l = L[j];
                return l >= 0 && out[uf_find(L, l)] != 0;
            };
            if (open(col > 0, i - 1) || open(col < A.W - 1, i + 1) || open(row > 0, i - A.W) || open(row < A.H - 1, i + A.W))
                b |= (unsigned char)(1 << plane);
        }
        if (b) {
            A.bnd[i] = b;
            long long* c = A.counts + (i / hw) * 4;
            if (b & 1) atomicAdd(reinterpret_cast<unsigned long long*>(c + 0), 1ull);
            if (b & 2) atomicAdd(reinterpret_cast<unsigned long long*>(c + 1), 1ull);
        }
    }
}

struct Offsets {
    int n;
    signed char dy[kMaxOffsets], dx[kMaxOffsets];
};

// boundary pixels of one map that have a boundary pixel of the OTHER map within the chamfer tolerance
__global__ void __launch_bounds__(kThreads) bf1_match(const BArgs A, const Offsets O) {
    const long long hw = (long long)A.H * A.W;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < A.n; i += (long long)gridDim.x * kThreads) {
        const unsigned char b = A.bnd[i];
        if (b == 0) continue;
        const int col = (int)(i % A.W);
        const int row = (int)((i / A.W) % A.H);
        unsigned char seen = 0;  // bits of the maps that have a boundary pixel in the neighbourhood
        for (int k = 0; k < O.n && seen != 3; ++k) {
            const int r = row + O.dy[k], c = col + O.dx[k];
            if (r < 0 || r >= A.H || c < 0 || c >= A.W) continue;
            seen |= A.bnd[i + (long long)O.dy[k] * A.W + O.dx[k]];
        }
        long long* cnt = A.counts + (i / hw) * 4;
        if ((b & 1) && (seen & 2)) atomicAdd(reinterpret_cast<unsigned long long*>(cnt + 2), 1ull);  // prediction boundary near the target's
        if ((b & 2) && (seen & 1)) atomicAdd(reinterpret_cast<unsigned long long*>(cnt + 3), 1ull);  // target boundary near the prediction's
    }
}

// src/evaluate.py:171-191 in float32, the way NumPy 2 evaluates it (float32 sums, Python-float smooth is weak)
__global__ void bf1_finalize(const long long* counts, long long B, int tolerance, float smooth, float* f1) {
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
        const float np_ = (float)counts[4 * b], nt = (float)counts[4 * b + 1];
        float v;
        if (tolerance > 0) {
            const float precision = __fdiv_rn((float)counts[4 * b + 2] + smooth, np_ + smooth);
            const float recall = __fdiv_rn((float)counts[4 * b + 3] + smooth, nt + smooth);
            v = __fdiv_rn(__fmul_rn(__fmul_rn(2.0f, precision), recall) + smooth, (precision + recall) + smooth);
        } else {
            v = __fdiv_rn(__fmul_rn(2.0f, (float)counts[4 * b + 2]) + smooth, (np_ + nt) + smooth);
        }
        f1[b] = v;
    }
}

size_t align_up(size_t v) { return (v + 255) / 256 * 256; }

// OpenCV distanceTransform DIST_L2 mask 5 in its fixed-point form (DIST_SHIFT 16): cost of the cheapest chain of
// (1,0) / (1,1) / (2,1) steps between two pixels dy rows and dx columns apart
long long chamfer5_fix(int dy, int dx) {
    const long long HV = 65536, DIAG = 91750, LONG = 143976;  // cvRound(1.0 * 2^16), (1.4f ...), (2.1969f ...)
    int a = dy < 0 ? -dy : dy, b = dx < 0 ? -dx : dx;
    if (a > b) {
        const int s = a;
        a = b;
        b = s;
    }  // a <= b
    return b >= 2 * a ? LONG * a + HV * (b - 2 * a) : LONG * (b - a) + DIAG * (2 * a - b);
}

}  // namespace

extern "C" {

size_t pil_boundary_workspace_bytes(int64_t B, int64_t H, int64_t W) {
    if (B < 1 || H < 1 || W < 1) return 0;
    const size_t n = (size_t)B * H * W;
    return align_up(n) + align_up(2 * n * sizeof(int)) + align_up(2 * n) + align_up(n);
}

int pil_boundary_tolerance_offsets(int tolerance, int8_t* dy_out, int8_t* dx_out, int capacity) {
    if (tolerance < 0 || tolerance > 6) return PIL_ERR_SHAPE;
    int n = 0;
    for (int dy = -tolerance; dy <= tolerance; ++dy)
        for (int dx = -tolerance; dx <= tolerance; ++dx)
            if (chamfer5_fix(dy, dx) <= (long long)tolerance * 65536) {
                if (dy_out && dx_out) {
                    if (n >= capacity) return PIL_ERR_WORKSPACE;
                    dy_out[n] = (int8_t)dy;
                    dx_out[n] = (int8_t)dx;
                }
                ++n;
            }
    return n;
}

int pil_boundary_counts(const void* x, const void* t, int64_t B, int64_t H, int64_t W, int x_dtype, int t_dtype, int x_kind,
                        float threshold, int tolerance, long long* counts, void* workspace, size_t workspace_bytes, void* stream) {
    if (!x || !t || !counts || !workspace) return PIL_ERR_NULL;
    if (B < 1 || H < 1 || W < 1 || B * H * W >= ((int64_t)1 << 31)) return PIL_ERR_SHAPE;
    if (!(x_dtype == PIL_F32 || x_dtype == PIL_BF16)) return PIL_ERR_DTYPE;
    if (!(t_dtype == PIL_F32 || t_dtype == PIL_BF16 || t_dtype == PIL_U8)) return PIL_ERR_DTYPE;
    if (x_kind < PIL_X_PROB || x_kind > PIL_X_LOGITS_TANH) return PIL_ERR_KIND;
    if (tolerance < 0 || tolerance > 6) return PIL_ERR_SHAPE;
    if (workspace_bytes < pil_boundary_workspace_bytes(B, H, W) || ((uintptr_t)workspace % 16)) return PIL_ERR_WORKSPACE;
    const size_t n = (size_t)B * H * W;
    BArgs a;
    a.x = x;
    a.t = t;
    a.n = (long long)n;
    a.H = (int)H;
    a.W = (int)W;
    a.x_dtype = x_dtype;
    a.t_dtype = t_dtype;
    a.x_kind = x_kind;
    a.threshold = threshold;
    unsigned char* w = reinterpret_cast<unsigned char*>(workspace);
    a.fg = w;
    w += align_up(n);
    a.label = reinterpret_cast<int*>(w);
    w += align_up(2 * n * sizeof(int));
    a.outside = w;
    w += align_up(2 * n);
    a.bnd = w;
    a.counts = counts;
    Offsets o;
    o.n = pil_boundary_tolerance_offsets(tolerance, reinterpret_cast<int8_t*>(o.dy), reinterpret_cast<int8_t*>(o.dx), kMaxOffsets);
    if (o.n < 1) return PIL_ERR_SHAPE;
    cudaStream_t s = (cudaStream_t)stream;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    auto grid = [&](size_t work) {
        size_t b = (work + kThreads - 1) / kThreads;
        const size_t cap = (size_t)sms * 16;
        return (int)(b < 1 ? 1 : (b > cap ? cap : b));
    };
    cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)B * 4 * sizeof(long long), s);
    if (e != cudaSuccess) return (int)e;
    bf1_init<<<grid(n), kThreads, 0, s>>>(a);
    bf1_merge<<<grid(2 * n), kThreads, 0, s>>>(a);
    bf1_flatten_mark<<<grid(2 * n), kThreads, 0, s>>>(a);
    bf1_boundary<<<grid(n), kThreads, 0, s>>>(a);
    bf1_match<<<grid(n), kThreads, 0, s>>>(a, o);
    return (int)cudaGetLastError();
}

int pil_boundary_f1(const long long* counts, int64_t B, int tolerance, double smooth, float* f1_out, void* stream) {
    if (!counts || !f1_out) return PIL_ERR_NULL;
    if (B < 1) return PIL_ERR_SHAPE;
    const int blocks = (int)((B + 127) / 128 < 64 ? (B + 127) / 128 : 64);
    bf1_finalize<<<blocks, 128, 0, (cudaStream_t)stream>>>(counts, (long long)B, tolerance, (float)smooth, f1_out);
    return (int)cudaGetLastError();
}

}  // extern "C"
