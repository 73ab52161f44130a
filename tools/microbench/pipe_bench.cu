// pipe_bench.cu -- throughput of the arithmetic instruction forms the fused kernels are built from, measured on
// the GPU it runs on (development tool; results recorded in DESIGN.md).  One block per SM, W warps per SMSP,
// 8 independent dependency chains per thread; prints cycles per warp-instruction per SMSP.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_bench pipe_bench.cu && ./pipe_bench
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>

constexpr int CH = 8;      // independent chains per thread
constexpr int IT = 2048;   // loop iterations; each iteration = CH instructions of the form under test

template <int MODE>
__global__ void bench(float* out, unsigned long long* cyc, float k0, float k1, float k2) {
    float a[CH], b[CH], c[CH];
    float2 A[CH], B[CH], C[CH];
    unsigned int H[CH], HB[CH], HC[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        a[i] = threadIdx.x * 1e-3f + i;
        b[i] = 1.0f + 1e-6f * (threadIdx.x + i);
        c[i] = 0.5f * i;
        A[i] = make_float2(a[i], a[i] + 1.f);
        B[i] = make_float2(b[i], b[i] + 1e-6f);
        C[i] = make_float2(c[i], c[i] + 1.f);
        H[i] = 0x3f803f80u + i;
        HB[i] = 0x3f813f81u + threadIdx.x;
        HC[i] = 0x3c003c00u + i;
    }
    __syncthreads();
    const unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < IT; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if constexpr (MODE == 0) {  // FFMA, three distinct registers
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(c[i]) : "f"(a[i]), "f"(b[i]));
            } else if constexpr (MODE == 1) {  // FFMA, one multiplicand from the constant bank (kernel parameter)
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(c[i]) : "f"(a[i]), "f"(k0));
            } else if constexpr (MODE == 2) {  // FFMA2 all-register
                asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%0, %1};"
                             " fma.rn.f32x2 rc, ra, rb, rc; mov.b64 {%0, %1}, rc;}"
                             : "+f"(C[i].x), "+f"(C[i].y) : "f"(A[i].x), "f"(A[i].y), "f"(B[i].x), "f"(B[i].y));
            } else if constexpr (MODE == 3) {  // FFMA2 with a broadcast scalar multiplicand
                asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %4}; mov.b64 rc, {%0, %1};"
                             " fma.rn.f32x2 rc, ra, rb, rc; mov.b64 {%0, %1}, rc;}"
                             : "+f"(C[i].x), "+f"(C[i].y) : "f"(A[i].x), "f"(A[i].y), "f"(k0));
            } else if constexpr (MODE == 4) {  // FADD2
                asm volatile("{.reg .b64 ra, rc; mov.b64 ra, {%2, %3}; mov.b64 rc, {%0, %1};"
                             " add.rn.f32x2 rc, ra, rc; mov.b64 {%0, %1}, rc;}"
                             : "+f"(C[i].x), "+f"(C[i].y) : "f"(A[i].x), "f"(A[i].y));
            } else if constexpr (MODE == 5) {  // FMUL2
                asm volatile("{.reg .b64 ra, rc; mov.b64 ra, {%2, %3}; mov.b64 rc, {%0, %1};"
                             " mul.rn.f32x2 rc, ra, rc; mov.b64 {%0, %1}, rc;}"
                             : "+f"(C[i].x), "+f"(C[i].y) : "f"(B[i].x), "f"(B[i].y));
            } else if constexpr (MODE == 6) {  // HFMA2.BF16, three registers
                asm volatile("fma.rn.bf16x2 %0, %1, %2, %0;" : "+r"(HC[i]) : "r"(H[i]), "r"(HB[i]));
            } else if constexpr (MODE == 7) {  // FADD scalar, two registers
                asm volatile("add.rn.f32 %0, %1, %0;" : "+f"(c[i]) : "f"(a[i]));
            } else if constexpr (MODE == 8) {  // MUFU.EX2
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c[i]));
            } else if constexpr (MODE == 9) {  // FFMA: a*b + c with b == a (two distinct source registers)
                asm volatile("fma.rn.f32 %0, %1, %1, %0;" : "+f"(c[i]) : "f"(a[i]));
            } else if constexpr (MODE == 10) {  // HFMA2 fp16, three registers
                asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(HC[i]) : "r"(H[i]), "r"(HB[i]));
            } else if constexpr (MODE == 11) {  // FFMA2 : rc = ra*ra + rc (two distinct 64-bit sources)
                asm volatile("{.reg .b64 ra, rc; mov.b64 ra, {%2, %3}; mov.b64 rc, {%0, %1};"
                             " fma.rn.f32x2 rc, ra, ra, rc; mov.b64 {%0, %1}, rc;}"
                             : "+f"(C[i].x), "+f"(C[i].y) : "f"(A[i].x), "f"(A[i].y));
            } else if constexpr (MODE == 12) {  // alternating FFMA2 and MUFU (do they overlap?)
                asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%0, %1};"
                             " fma.rn.f32x2 rc, ra, rb, rc; mov.b64 {%0, %1}, rc;}"
                             : "+f"(C[i].x), "+f"(C[i].y) : "f"(A[i].x), "f"(A[i].y), "f"(B[i].x), "f"(B[i].y));
                if (i % 4 == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c[i]));
            } else if constexpr (MODE >= 20 && MODE <= 23) {  // FFMA2 (broadcast form, 2 pipe cycles) + MUFU.EX2 on 1/8, 1/4, 1/2, 1/1 of them
                asm volatile("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %4}; mov.b64 rc, {%0, %1};"
                             " fma.rn.f32x2 rc, ra, rb, rc; mov.b64 {%0, %1}, rc;}"
                             : "+f"(C[i].x), "+f"(C[i].y) : "f"(A[i].x), "f"(A[i].y), "f"(k0));
                constexpr int every = MODE == 20 ? 8 : (MODE == 21 ? 4 : (MODE == 22 ? 2 : 1));
                if (i % every == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c[i]));
            } else if constexpr (MODE == 13) {  // alternating FFMA (const operand) and integer IADD3 (alu pipe)
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(c[i]) : "f"(a[i]), "f"(k0));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(H[i]) : "r"(HB[i]));
            }
        }
    }
    const unsigned long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c[i] + C[i].x + C[i].y + __uint_as_float(HC[i]) + __uint_as_float(H[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(const char* name, int instr_per_iter_per_chain, float* out, unsigned long long* cyc, int sms) {
    for (int wps : {1, 2, 4, 8}) {
        const int threads = wps * 4 * 32;
        bench<MODE><<<sms, threads>>>(out, cyc, 1.0000001f, 0.5f, 0.25f);
        cudaDeviceSynchronize();
        bench<MODE><<<sms, threads>>>(out, cyc, 1.0000001f, 0.5f, 0.25f);
        cudaDeviceSynchronize();
        unsigned long long h[256];
        cudaMemcpy(h, cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);
        double m = 0;
        for (int i = 0; i < sms; ++i) m += (double)h[i];
        m /= sms;
        const double winstr = (double)IT * CH * instr_per_iter_per_chain * wps;  // warp-instructions per SMSP
        printf("%-46s warps/SMSP %d : %.3f cycles per warp-instruction per SMSP\n", name, wps, m / winstr);
    }
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out;
    unsigned long long* cyc;
    cudaMalloc(&out, sizeof(float) * 256 * 1024);
    cudaMalloc(&cyc, sizeof(unsigned long long) * 256);
    run<0>("FFMA  R,R,R (3 distinct regs)", 1, out, cyc, sms);
    run<9>("FFMA  R,R(same),R", 1, out, cyc, sms);
    run<1>("FFMA  R,c[],R (constant-bank operand)", 1, out, cyc, sms);
    run<7>("FADD  R,R", 1, out, cyc, sms);
    run<2>("FFMA2 RR,RR,RR", 1, out, cyc, sms);
    run<11>("FFMA2 RR,RR(same),RR", 1, out, cyc, sms);
    run<3>("FFMA2 RR,bcast(c[]),RR", 1, out, cyc, sms);
    run<4>("FADD2 RR,RR", 1, out, cyc, sms);
    run<5>("FMUL2 RR,RR", 1, out, cyc, sms);
    run<6>("HFMA2.BF16 R,R,R", 1, out, cyc, sms);
    run<10>("HFMA2 (fp16) R,R,R", 1, out, cyc, sms);
    run<8>("MUFU.EX2", 1, out, cyc, sms);
    run<12>("FFMA2 + 1/4 MUFU.EX2 (per FFMA2)", 1, out, cyc, sms);
    run<13>("FFMA c[] + IADD (per pair)", 1, out, cyc, sms);
    // do the FMA and XU pipes overlap?  cycles per FFMA2 (2.0 alone); the MUFU share alone would cost 1.0 / 2.0 / 4.0 / 8.0
    run<20>("FFMA2 bcast + 1/8 MUFU.EX2 (per FFMA2)", 1, out, cyc, sms);
    run<21>("FFMA2 bcast + 1/4 MUFU.EX2 (per FFMA2)", 1, out, cyc, sms);
    run<22>("FFMA2 bcast + 1/2 MUFU.EX2 (per FFMA2)", 1, out, cyc, sms);
    run<23>("FFMA2 bcast + 1/1 MUFU.EX2 (per FFMA2)", 1, out, cyc, sms);
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return e == cudaSuccess ? 0 : 1;
}
