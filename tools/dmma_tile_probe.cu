// Probe: what does the bare consumer loop of gb_fourier_stage2_sym reach without barriers, copies and stores?
// VAR 0: as shipped (4 sets, chunks of <= 7 k-steps with an early exit, accumulators cleared per tile)
// VAR 1: fixed 7-step chunks (no early exit)          VAR 2: fragment loads of step i+1 issued before the DMMAs of step i
// VAR 3: VAR 0 with 12 warps (3 per sub-partition)     VAR 4: VAR 2 with 12 warps
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
constexpr int KC = 28, LDA = 132, LDB = 36, STAGES = 4;
struct Groups { int off[5]; };

template <int VAR>
__global__ void __launch_bounds__(VAR >= 3 ? 384 : 256, 1) probe(double* out, int tiles, Groups grp) {
    extern __shared__ double smem[];
    for (int i = threadIdx.x; i < STAGES * KC * (LDA + LDB); i += blockDim.x) smem[i] = 1e-3 * (i % 7);
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3, warp = threadIdx.x >> 5;
    const int wm = warp & 3, wn = (warp >> 2) & 1;
    double keep = 0.0;
    int stage = 0;
    for (int t = 0; t < tiles; ++t) {
        double acc[4][4][2][2];
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) acc[s][mi][ni][0] = acc[s][mi][ni][1] = 0.0;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            for (int k0 = grp.off[s]; k0 < grp.off[s + 1];) {
                const int kc = (VAR == 1) ? KC : min(KC, grp.off[s + 1] - k0);
                const double* sA = smem + (size_t)stage * KC * (LDA + LDB) + wm * 32 + g;
                const double* sB = smem + (size_t)stage * KC * (LDA + LDB) + KC * LDA + wn * 16 + g;
                if (VAR == 2 || VAR == 4) {
                    double a[4], b[2];
#pragma unroll
                    for (int mi = 0; mi < 4; ++mi) a[mi] = sA[q * LDA + mi * 8];
#pragma unroll
                    for (int ni = 0; ni < 2; ++ni) b[ni] = sB[q * LDB + ni * 8];
#pragma unroll
                    for (int kk = 0; kk < KC; kk += 4) {
                        if (kk >= kc) break;
                        double an[4], bn[2];
                        const int kn = (kk + 4 < kc) ? kk + 4 : kk;
#pragma unroll
                        for (int mi = 0; mi < 4; ++mi) an[mi] = sA[(kn + q) * LDA + mi * 8];
#pragma unroll
                        for (int ni = 0; ni < 2; ++ni) bn[ni] = sB[(kn + q) * LDB + ni * 8];
#pragma unroll
                        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                            for (int ni = 0; ni < 2; ++ni) dmma(acc[s][mi][ni][0], acc[s][mi][ni][1], a[mi], b[ni]);
#pragma unroll
                        for (int mi = 0; mi < 4; ++mi) a[mi] = an[mi];
#pragma unroll
                        for (int ni = 0; ni < 2; ++ni) b[ni] = bn[ni];
                    }
                } else {
#pragma unroll
                    for (int kk = 0; kk < KC; kk += 4) {
                        if (VAR != 1 && kk >= kc) break;
                        double a[4], b[2];
#pragma unroll
                        for (int mi = 0; mi < 4; ++mi) a[mi] = sA[(kk + q) * LDA + mi * 8];
#pragma unroll
                        for (int ni = 0; ni < 2; ++ni) b[ni] = sB[(kk + q) * LDB + ni * 8];
#pragma unroll
                        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                            for (int ni = 0; ni < 2; ++ni) dmma(acc[s][mi][ni][0], acc[s][mi][ni][1], a[mi], b[ni]);
                    }
                }
                __syncwarp();
                if (++stage == STAGES) stage = 0;
                k0 += (VAR == 1) ? min(KC, grp.off[s + 1] - k0) : kc;
            }
        }
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) keep += acc[s][mi][ni][0] + acc[s][mi][ni][1];
    }
    if (keep == 123.456) out[threadIdx.x] = keep;
}

template <int VAR>
void run(int sms, double* out, const Groups& grp, const char* what) {
    const int tiles = 28 * 20;
    const int threads = VAR >= 3 ? 384 : 256;
    const size_t smem = (size_t)STAGES * KC * (LDA + LDB) * sizeof(double);
    CK(cudaFuncSetAttribute(probe<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e9;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        probe<VAR><<<sms, threads, smem>>>(out, tiles, grp);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r > 0) best = std::min(best, ms);
    }
    CK(cudaGetLastError());
    // executed k-steps per tile (VAR 1 runs whole 7-step chunks)
    long steps = 0;
    for (int s = 0; s < 4; ++s)
        for (int k0 = grp.off[s]; k0 < grp.off[s + 1]; k0 += KC) steps += (VAR == 1) ? 7 : (std::min(KC, grp.off[s + 1] - k0) / 4);
    const double flops = 2.0 * 256 * 8 * (double)steps * tiles * (threads / 32) * sms;
    printf("VAR %d (%s): %.3f ms per 28 tiles, %.2f TFLOP/s\n", VAR, what, best / 20, flops / (best * 1e-3) / 1e12);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    double* out; CK(cudaMalloc(&out, 4096));
    Groups grp{{0, 52, 100, 148, 196}};
    run<0>(prop.multiProcessorCount, out, grp, "as shipped, 8 warps");
    run<1>(prop.multiProcessorCount, out, grp, "fixed 7-step chunks");
    run<2>(prop.multiProcessorCount, out, grp, "register double buffer");
    run<3>(prop.multiProcessorCount, out, grp, "as shipped, 12 warps");
    run<4>(prop.multiProcessorCount, out, grp, "double buffer, 12 warps");
    return 0;
}
