// pil_fwdrow.cuh -- per-pixel forward terms shared by the full forward kernel (pil_fwd.cu) and the pointwise
// forward kernels (pil_point.cu).
#ifndef PIL_FWDROW_CUH_
#define PIL_FWDROW_CUH_
#include "pil_common.cuh"

namespace pil {
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
    // hs = left + right neighbour, dx = right - left neighbour of the pair (both formed with SCALAR adds on the six values
    // {L, s0..s3, R} by the caller: packing the misaligned pairs (L,s0) (s1,s2) (s3,R) would cost register moves)
    __device__ __forceinline__ void stencil2(f2 u, f2 hs, f2 dx, f2 m, f2 p) {
        const f2 s4 = add2(hs, add2(m, p));                                                 // src/pde.py:73-77
        const f2 dy = sub2(p, m);                                                           // 2*gy (src/pde.py:173); dx = 2*gx (:172)
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
        const f2 hs0 = make_float2(L + uc.y, uc.x + uc.z), hs1 = make_float2(uc.y + uc.w, uc.z + R);
        const f2 dx0 = make_float2(uc.y - L, uc.z - uc.x), dx1 = make_float2(uc.w - uc.y, R - uc.z);
        stencil2(make_float2(uc.x, uc.y), hs0, dx0, make_float2(um.x, um.y), make_float2(up.x, up.y));
        stencil2(make_float2(uc.z, uc.w), hs1, dx1, make_float2(um.z, um.w), make_float2(up.z, up.w));
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

}  // namespace pil
#endif  // PIL_FWDROW_CUH_
