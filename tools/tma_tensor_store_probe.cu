// Rate of 2-D TMA tensor stores (cp.async.bulk.tensor.2d shared -> global) of [ROWS x 16] double boxes per SM: would an
// epilogue that hands 8-row x 128-byte boxes to the TMA unit beat 32-byte-per-lane stores?  (development probe)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tma_tensor_store_probe tools/tma_tensor_store_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void tensor_store_2d(const CUtensorMap* tm, const void* ssrc, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(tm),
                 "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(c0), "r"(c1)
                 : "memory");
}

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tm, int rows_per_cta, int box_rows, int nlon, int iters,
                                             int issuers) {
    extern __shared__ __align__(1024) unsigned char smem[];
    for (int i = threadIdx.x; i < 32768 / 8; i += blockDim.x) reinterpret_cast<double*>(smem)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    if ((int)threadIdx.x >= issuers) return;
    const int row0 = blockIdx.x * rows_per_cta;
    const int boxes_per_row = nlon / 16;
    const int n_box = rows_per_cta / box_rows * boxes_per_row;
    for (int it = 0; it < iters; ++it) {
        int k = 0;
        for (int b = threadIdx.x; b < n_box; b += issuers, ++k) {
            const int r = b / boxes_per_row, c = b % boxes_per_row;
            tensor_store_2d(&tm, smem + (b % 4) * box_rows * 128, c * 16, row0 + r * box_rows);
            if (k % 8 == 7) {
                asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 4;\n" ::: "memory");
            }
        }
        asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
    }
}

int main() {
    const int nlon = 720, rows_per_cta = 584, rows = rows_per_cta * 148;     // 86 432 rows: 498 MB
    double* d;
    cudaMalloc(&d, (size_t)rows * nlon * 8);
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q);
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int box_rows : {8, 32})
        for (int issuers : {1, 4, 32, 128}) {
            CUtensorMap tm;
            cuuint64_t dims[2] = {(cuuint64_t)nlon, (cuuint64_t)rows};
            cuuint64_t strides[1] = {(cuuint64_t)nlon * 8};
            cuuint32_t box[2] = {16, (cuuint32_t)box_rows};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
            cudaEvent_t a, b;
            cudaEventCreate(&a); cudaEventCreate(&b);
            const int iters = 4;
            probe<<<148, 128, 65536>>>(tm, rows_per_cta, box_rows, nlon, 1, issuers);
            cudaDeviceSynchronize();
            cudaEventRecord(a);
            probe<<<148, 128, 65536>>>(tm, rows_per_cta, box_rows, nlon, iters, issuers);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms = 0;
            cudaEventElapsedTime(&ms, a, b);
            const double n_box = (double)(rows_per_cta / box_rows) * (nlon / 16) * iters;     // per CTA
            const double bytes = n_box * box_rows * 128 * 148;
            printf("box %2d rows x 128 B, %3d issuing threads: %7.1f GB/s total, %5.1f B/clk/SM, %6.1f cycles per box per SM  [%s]\n", box_rows,
                   issuers, bytes / ms / 1e6, bytes / 148 / (ms * 1e-3) / 1.965e9, ms * 1e-3 * 1.965e9 / n_box,
                   cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
