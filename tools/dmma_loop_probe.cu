// Probe: which ingredient of the symmetric stage-2 consumer loop costs DMMA throughput?
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// MODE 0: same a,b   MODE 1: 4 a x 2 b distinct regs   MODE 2: + LDS each step   MODE 3: 4 accumulator sets, switch every 6 steps
template <int MODE>
__global__ void __launch_bounds__(384, 1) probe(double* out, int steps, double a0, double b0) {
    __shared__ double smem[28 * 168];
    for (int i = threadIdx.x; i < 28 * 168; i += blockDim.x) smem[i] = 1e-3 * (i % 7);
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3, warp = threadIdx.x >> 5;
    double acc[4][4][2][2];
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) acc[s][mi][ni][0] = acc[s][mi][ni][1] = 0.0;
    double a[4] = {a0, a0 * 0.5, a0 * 0.25, a0 * 0.125}, b[2] = {b0, b0 * 0.5};
    const double* sA = smem + (warp & 3) * 32 + g;
    const double* sB = smem + 28 * 132 + (warp & 1) * 16 + g;
    if (MODE <= 2) {
        for (int t = 0; t < steps; ++t) {
            if (MODE == 2) {
                const int kk = (t % 7) * 4;
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) a[mi] = sA[(kk + q) * 132 + mi * 8];
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) b[ni] = sB[(kk + q) * 36 + ni * 8];
            }
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni)
                    dmma(acc[0][mi][ni][0], acc[0][mi][ni][1], MODE == 0 ? a[0] : a[mi], MODE == 0 ? b[0] : b[ni]);
        }
    } else {
        for (int t0 = 0; t0 < steps; t0 += 24) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
#pragma unroll 1
                for (int t = 0; t < 6; ++t) {
                    const int kk = t * 4;
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) a[mi] = sA[(kk + q) * 132 + mi * 8];
#pragma unroll
                    for (int ni = 0; ni < 2; ++ni) b[ni] = sB[(kk + q) * 36 + ni * 8];
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                        for (int ni = 0; ni < 2; ++ni) dmma(acc[s][mi][ni][0], acc[s][mi][ni][1], a[mi], b[ni]);
                }
            }
        }
    }
    double z = 0;
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) z += acc[s][mi][ni][0] + acc[s][mi][ni][1];
    if (z == 123.456) out[threadIdx.x] = z;
}

template <int MODE>
double run(int threads, int sms, double* out) {
    const int steps = 24 * 2000;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e9;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        probe<MODE><<<sms, threads>>>(out, steps, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r > 0) best = std::min(best, ms);
    }
    CK(cudaGetLastError());
    return 2.0 * 256 * 8 * (double)steps * (threads / 32) * sms / (best * 1e-3) / 1e12;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    double* out; CK(cudaMalloc(&out, 4096));
    for (int threads : {256, 384}) {
        printf("threads=%d: same-ab %.2f | distinct-ab %.2f | +LDS %.2f | +4 sets/unroll1 %.2f TFLOP/s\n", threads,
               run<0>(threads, prop.multiProcessorCount, out), run<1>(threads, prop.multiProcessorCount, out),
               run<2>(threads, prop.multiProcessorCount, out), run<3>(threads, prop.multiProcessorCount, out));
    }
    return 0;
}
