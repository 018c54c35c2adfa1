// FP64 peak probes for B200 (sm_100a): DFMA vector pipe vs DMMA tensor pipe.
// Stand-alone (nvcc only). Prints one JSON object. The roofline denominators that
// MEASURED_PEAKS.json lacks (no FP64 figure) come from here and from cuBLAS DGEMM.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peaks tools/fp64_peaks.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x * 1e-9 + c;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int TILES>
__global__ void __launch_bounds__(256) dmma884_kernel(double* out, int iters, double a, double b) {
    double acc[TILES][2];
#pragma unroll
    for (int c = 0; c < TILES; ++c) { acc[c][0] = threadIdx.x * 1e-9; acc[c][1] = c; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < TILES; ++c) dmma884(acc[c][0], acc[c][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < TILES; ++c) s += acc[c][0] + acc[c][1];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int TILES>
__global__ void __launch_bounds__(256) dmma1688_kernel(double* out, int iters, double a0, double b0) {
    double acc[TILES][4];
    double a[4] = {a0, a0 * 0.5, a0 * 0.25, a0 * 0.125};
    double b[2] = {b0, b0 * 0.5};
#pragma unroll
    for (int c = 0; c < TILES; ++c) { acc[c][0] = threadIdx.x * 1e-9; acc[c][1] = c; acc[c][2] = 1; acc[c][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < TILES; ++c) dmma1688(acc[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < TILES; ++c) s += acc[c][0] + acc[c][1] + acc[c][2] + acc[c][3];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int TILES>
__global__ void __launch_bounds__(256) dmma16816_kernel(double* out, int iters, double a0, double b0) {
    double acc[TILES][4];
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a0 / (i + 1);
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = b0 / (i + 1);
#pragma unroll
    for (int c = 0; c < TILES; ++c) { acc[c][0] = threadIdx.x * 1e-9; acc[c][1] = c; acc[c][2] = 1; acc[c][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < TILES; ++c) dmma16816(acc[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < TILES; ++c) s += acc[c][0] + acc[c][1] + acc[c][2] + acc[c][3];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Mixed: DMMA and DFMA interleaved in one warp -- do the two pipes add up or share hardware?
template <int TILES, int CHAINS>
__global__ void __launch_bounds__(256) mixed_kernel(double* out, int iters, double a, double b) {
    double acc[TILES][2];
    double f[CHAINS];
#pragma unroll
    for (int c = 0; c < TILES; ++c) { acc[c][0] = threadIdx.x * 1e-9; acc[c][1] = c; }
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) f[c] = threadIdx.x * 1e-9 + c;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < TILES; ++c) dmma884(acc[c][0], acc[c][1], a, b);
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) f[c] = fma(f[c], a, b);
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < TILES; ++c) s += acc[c][0] + acc[c][1];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += f[c];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

struct Result { double best_ms; double med_ms; };

template <typename F>
Result time_it(F launch, int reps = 7) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    std::vector<float> ms(reps);
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms[r], e0, e1));
    }
    CK(cudaGetLastError());
    std::sort(ms.begin(), ms.end());
    return {ms[0], ms[reps / 2]};
}

int main(int argc, char** argv) {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    int sms = prop.multiProcessorCount;
    double* out;
    CK(cudaMalloc(&out, sizeof(double) * 1 << 20));
    const int iters = 20000;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", prop.name, sms, prop.clockRate);

    // blocks per SM sweep: 256 threads per block
    for (int bps : {1, 2, 4, 8}) {
        int grid = sms * bps;
        {
            auto r = time_it([&] { dfma_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); });
            double fl = 2.0 * 8 * iters * 256.0 * grid;
            printf(" \"dfma_c8_bps%d_tflops\": %.3f,\n", bps, fl / (r.best_ms * 1e-3) / 1e12);
        }
        {
            auto r = time_it([&] { dmma884_kernel<8><<<grid, 256>>>(out, iters / 4, 1.0000001, 1e-9); });
            double fl = 2.0 * 256 * 8 * (iters / 4) * 8.0 * grid;  // 256 FMA per warp-instr, 8 warps per block
            printf(" \"dmma884_t8_bps%d_tflops\": %.3f,\n", bps, fl / (r.best_ms * 1e-3) / 1e12);
        }
        {
            auto r = time_it([&] { dmma1688_kernel<4><<<grid, 256>>>(out, iters / 8, 1.0000001, 1e-9); });
            double fl = 2.0 * 1024 * 4 * (iters / 8) * 8.0 * grid;
            printf(" \"dmma1688_t4_bps%d_tflops\": %.3f,\n", bps, fl / (r.best_ms * 1e-3) / 1e12);
        }
        {
            auto r = time_it([&] { dmma16816_kernel<4><<<grid, 256>>>(out, iters / 16, 1.0000001, 1e-9); });
            double fl = 2.0 * 2048 * 4 * (iters / 16) * 8.0 * grid;
            printf(" \"dmma16816_t4_bps%d_tflops\": %.3f,\n", bps, fl / (r.best_ms * 1e-3) / 1e12);
        }
    }
    {
        int grid = sms * 4;
        auto r = time_it([&] { dmma884_kernel<16><<<grid, 256>>>(out, iters / 4, 1.0000001, 1e-9); });
        double fl = 2.0 * 256 * 16 * (iters / 4) * 8.0 * grid;
        printf(" \"dmma884_t16_bps4_tflops\": %.3f,\n", fl / (r.best_ms * 1e-3) / 1e12);
        auto r2 = time_it([&] { dmma884_kernel<2><<<grid, 256>>>(out, iters / 4, 1.0000001, 1e-9); });
        double fl2 = 2.0 * 256 * 2 * (iters / 4) * 8.0 * grid;
        printf(" \"dmma884_t2_bps4_tflops\": %.3f,\n", fl2 / (r2.best_ms * 1e-3) / 1e12);
        auto r3 = time_it([&] { dmma884_kernel<1><<<sms, 32>>>(out, iters, 1.0000001, 1e-9); });
        // latency probe: one warp per SM, one dependent chain: ns per DMMA
        printf(" \"dmma884_dep_chain_ns\": %.2f,\n", r3.best_ms * 1e6 / iters);
        auto r4 = time_it([&] { dfma_kernel<1><<<sms, 32>>>(out, iters, 1.0000001, 1e-9); });
        printf(" \"dfma_dep_chain_ns\": %.2f,\n", r4.best_ms * 1e6 / iters);
    }
    {
        int grid = sms * 4;
        auto r = time_it([&] { mixed_kernel<8, 8><<<grid, 256>>>(out, iters / 4, 1.0000001, 1e-9); });
        double fl_mma = 2.0 * 256 * 8 * (iters / 4) * 8.0 * grid;
        double fl_fma = 2.0 * 8 * (iters / 4) * 256.0 * grid;
        printf(" \"mixed_t8_c8_tflops_total\": %.3f, \"mixed_mma_share\": %.3f,\n",
               (fl_mma + fl_fma) / (r.best_ms * 1e-3) / 1e12, fl_mma / (fl_mma + fl_fma));
        auto r2 = time_it([&] { mixed_kernel<8, 32><<<grid, 256>>>(out, iters / 4, 1.0000001, 1e-9); });
        double fl_fma2 = 2.0 * 32 * (iters / 4) * 256.0 * grid;
        printf(" \"mixed_t8_c32_tflops_total\": %.3f, \"mixed2_mma_share\": %.3f,\n",
               (fl_mma + fl_fma2) / (r2.best_ms * 1e-3) / 1e12, fl_mma / (fl_mma + fl_fma2));
    }
    // sustained: run DFMA and DMMA back-to-back for ~2 s each, report average
    {
        int grid = sms * 4;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        int n = 0; float ms = 0;
        CK(cudaEventRecord(e0));
        do { for (int k = 0; k < 10; ++k) dfma_kernel<8><<<grid, 256>>>(out, iters, 1.0000001, 1e-9); n += 10;
             CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1)); } while (ms < 2000);
        printf(" \"dfma_sustained_tflops\": %.3f,\n", 2.0 * 8 * iters * 256.0 * grid * n / (ms * 1e-3) / 1e12);
        n = 0;
        CK(cudaEventRecord(e0));
        do { for (int k = 0; k < 10; ++k) dmma884_kernel<8><<<grid, 256>>>(out, iters / 4, 1.0000001, 1e-9); n += 10;
             CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1)); } while (ms < 2000);
        printf(" \"dmma884_sustained_tflops\": %.3f,\n", 2.0 * 256 * 8 * (iters / 4) * 8.0 * grid * n / (ms * 1e-3) / 1e12);
    }
    // PCIe pinned copies
    {
        size_t bytes = 512ull << 20;
        void *h, *d;
        CK(cudaMallocHost(&h, bytes)); CK(cudaMalloc(&d, bytes));
        auto r1 = time_it([&] { CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice)); }, 5);
        auto r2 = time_it([&] { CK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost)); }, 5);
        printf(" \"h2d_pinned_gbs\": %.2f, \"d2h_pinned_gbs\": %.2f,\n", bytes / (r1.best_ms * 1e-3) / 1e9, bytes / (r2.best_ms * 1e-3) / 1e9);
        void* hp = malloc(bytes);
        memset(hp, 1, bytes);
        auto r3 = time_it([&] { CK(cudaMemcpy(hp, d, bytes, cudaMemcpyDeviceToHost)); }, 3);
        printf(" \"d2h_pageable_gbs\": %.2f,\n", bytes / (r3.best_ms * 1e-3) / 1e9);
        // device copy
        void* d2; CK(cudaMalloc(&d2, bytes));
        auto r4 = time_it([&] { CK(cudaMemcpyAsync(d2, d, bytes, cudaMemcpyDeviceToDevice)); }, 7);
        printf(" \"d2d_copy_gbs_rw\": %.1f,\n", 2.0 * bytes / (r4.best_ms * 1e-3) / 1e9);
    }
    printf(" \"done\": true}\n");
    return 0;
}
