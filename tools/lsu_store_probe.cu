// Store rate of one SM's LSU path for 32-byte-per-lane streaming stores, by the shape of the 1 KB a warp instruction
// writes: ROWS rows (far apart) x 1024 / ROWS contiguous bytes (development probe).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/lsu_store_probe tools/lsu_store_probe.cu && build/lsu_store_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int ROWS>
__global__ void __launch_bounds__(512) probe(double* out, size_t per_cta_bytes, int iters, int row_stride) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    constexpr int LPR = 32 / ROWS;                       // lanes per row
    unsigned char* base = reinterpret_cast<unsigned char*>(out) + (size_t)blockIdx.x * per_cta_bytes;
    // a warp instruction covers ROWS rows x (LPR * 32) bytes; successive instructions of a warp move along the rows
    const size_t lane_off = (size_t)(lane / LPR) * row_stride + (size_t)(lane % LPR) * 32;
    const int cols_per_row = row_stride / (LPR * 32);    // instructions until a row block is full
    const size_t block_bytes = (size_t)ROWS * row_stride;
    const int n_blocks = (int)(per_cta_bytes / block_bytes);
    const double v = lane;
    for (int it = 0; it < iters; ++it)
        for (int b = warp; b < n_blocks; b += nwarps) {
            unsigned char* p = base + (size_t)b * block_bytes + lane_off;
#pragma unroll 4
            for (int c = 0; c < cols_per_row; ++c)
                asm volatile("st.global.cs.v4.f64 [%0], {%1, %1, %1, %1};\n" ::"l"(p + (size_t)c * LPR * 32), "d"(v) : "memory");
        }
}

template <int ROWS>
void run(double* d, size_t per_cta, int stride, int threads) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = (int)((64u << 20) / per_cta);      // 64 MB per CTA in all
    probe<ROWS><<<148, threads>>>(d, per_cta, 1, stride);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    probe<ROWS><<<148, threads>>>(d, per_cta, iters, stride);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const size_t block_bytes = (size_t)ROWS * stride;
    const double bytes = (double)(per_cta / block_bytes) * block_bytes * 148 * iters;
    printf("%4zu KB per CTA, %2d rows x %4d B per instruction, %3d threads: %7.1f GB/s total, %5.1f B/clk/SM, %5.1f cycles per 1 KB instruction  [%s]\n",
           per_cta >> 10, ROWS, 1024 / ROWS, threads, bytes / ms / 1e6, bytes / 148 / (ms * 1e-3) / 1.965e9,
           1024.0 / (bytes / 148 / (ms * 1e-3) / 1.965e9), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    double* d;
    cudaMalloc(&d, (size_t)(4u << 20) * 148 + (1 << 20));
    // 4 MB per CTA: 592 MB, DRAM-bound; 256 KB per CTA: 37 MB rewritten in place, L2-resident (the LSU / L2 path alone)
    for (size_t per_cta : {(size_t)4 << 20, (size_t)256 << 10})
        for (int threads : {128, 512}) {
            run<8>(d, per_cta, 5760, threads);     // the Fourier stage: 8 output rows x 128 B
            run<4>(d, per_cta, 5760, threads);
            run<2>(d, per_cta, 5760, threads);
            run<1>(d, per_cta, 5760, threads);     // one row x 1 KB
        }
    return 0;
}
