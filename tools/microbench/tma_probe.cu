// tma_probe.cu -- minimal 2-D TMA box load for f32 / bf16 / u8 maps (development probe for the TMA stage ring).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap tm, int col, int row, int box_bytes, unsigned char* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm);
    const uint32_t bar = dst + 8192;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(box_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                     "l"(&tm), "r"(col), "r"(row), "r"(bar) : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(bar) : "memory");
    for (int i = threadIdx.x; i < box_bytes; i += blockDim.x) out[i] = sm[i];
}

int main() {
    cuInit(0);
    for (int esz : {4, 2, 1}) {
        const int W = 256, R = 64, box_cols = 128, box_rows = 6;
        std::vector<unsigned char> h((size_t)W * R * esz);
        for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)(i * 7 + 3);
        unsigned char *d, *o;
        cudaMalloc(&d, h.size());
        cudaMalloc(&o, 8192);
        cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
        CUtensorMap tm;
        const cuuint64_t dims[2] = {(cuuint64_t)W, (cuuint64_t)R};
        const cuuint64_t strides[1] = {(cuuint64_t)W * esz};
        const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
        const cuuint32_t es[2] = {1, 1};
        const CUtensorMapDataType dt = esz == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8);
        CUresult r = cuTensorMapEncodeTiled(&tm, dt, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const int bytes = box_cols * box_rows * esz;
        for (int col : {0, 116, -4}) {
            cudaMemset(o, 0xee, 8192);
            probe<<<1, 128, 8192 + 64>>>(tm, col, 10, bytes, o);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<unsigned char> got(bytes);
            cudaMemcpy(got.data(), o, bytes, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int rr = 0; rr < box_rows; ++rr)
                for (int cc = 0; cc < box_cols * esz; ++cc) {
                    const long gc = (long)col * esz + cc;
                    const unsigned char want = (gc < 0 || gc >= (long)W * esz) ? 0 : h[(size_t)(10 + rr) * W * esz + gc];
                    bad += got[(size_t)rr * box_cols * esz + cc] != want;
                }
            printf("esz %d encode %d col %d : %s, mismatches %d\n", esz, (int)r, col, cudaGetErrorString(e), bad);
            if (e != cudaSuccess) return 1;
        }
    }
    return 0;
}
