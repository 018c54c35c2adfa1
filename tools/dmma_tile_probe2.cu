// Probe 2: does running TWO small CTAs per SM (64-row tiles, 16 x 16 x 4-set warp tiles) hide the per-tile
// overheads (epilogue stores) of the symmetric stage 2 behind the other CTA's k-steps?
// Data resident in shared memory, no barriers / copies; the epilogue does the real butterfly + streaming stores.
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void st_cs_v2(double* p, double a, double b) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};\n" ::"l"(p), "d"(a), "d"(b) : "memory");
}
constexpr int KC = 28, LDB = 36, STAGES = 4;
struct Groups { int off[5]; };

// MI = 8-row fragments per warp in M (4: 128-row CTA tile, 2: 64-row CTA tile)
template <int MI, int MINB>
__global__ void __launch_bounds__(256, MINB) probe(double* out, int tiles_total, Groups grp, int nlon, int stores) {
    constexpr int TM = MI * 8 * 4;            // 4 warps in M
    constexpr int LDA = TM + 4;
    extern __shared__ double smem[];
    for (int i = threadIdx.x; i < STAGES * KC * (LDA + LDB); i += blockDim.x) smem[i] = 1e-3 * (i % 7);
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, q = lane & 3, warp = threadIdx.x >> 5;
    const int wm = warp >> 1, wn = warp & 1;
    const int h = nlon >> 1;
    int stage = 0;
    for (int t = blockIdx.x; t < tiles_total; t += gridDim.x) {
        const int mt = t / 6, nt = t % 6;
        double acc[4][MI][2][2];
#pragma unroll
        for (int s = 0; s < 4; ++s)
#pragma unroll
            for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) acc[s][mi][ni][0] = acc[s][mi][ni][1] = 0.0;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            for (int k0 = grp.off[s]; k0 < grp.off[s + 1];) {
                const int kc = min(KC, grp.off[s + 1] - k0);
                const double* sA = smem + (size_t)stage * KC * (LDA + LDB) + wm * (MI * 8) + g;
                const double* sB = smem + (size_t)stage * KC * (LDA + LDB) + KC * LDA + wn * 16 + g;
#pragma unroll
                for (int kk = 0; kk < KC; kk += 4) {
                    if (kk >= kc) break;
                    double a[MI], b[2];
#pragma unroll
                    for (int mi = 0; mi < MI; ++mi) a[mi] = sA[(kk + q) * LDA + mi * 8];
#pragma unroll
                    for (int ni = 0; ni < 2; ++ni) b[ni] = sB[(kk + q) * LDB + ni * 8];
#pragma unroll
                    for (int mi = 0; mi < MI; ++mi)
#pragma unroll
                        for (int ni = 0; ni < 2; ++ni) dmma(acc[s][mi][ni][0], acc[s][mi][ni][1], a[mi], b[ni]);
                }
                __syncwarp();
                if (++stage == STAGES) stage = 0;
                k0 += kc;
            }
        }
        const long long row_base = (long long)mt * TM + wm * (MI * 8) + g;
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) {
            double* orow = out + (size_t)(row_base + mi * 8) * nlon;
#pragma unroll
            for (int ni = 0; ni < 2; ++ni) {
                const int jq = nt * 32 + wn * 16 + ni * 8 + 2 * q;
                if (jq >= nlon / 4) continue;
                double v1[2], v2[2], v3[2], v4[2];
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const double ce = acc[0][mi][ni][r], co = acc[1][mi][ni][r], se = acc[2][mi][ni][r], so = acc[3][mi][ni][r];
                    const double cp = ce + co, cm = ce - co, sp = se + so, sm = se - so;
                    v1[r] = cp + sp; v2[r] = cm - sm; v3[r] = cp - sp; v4[r] = cm + sm;
                }
                if (stores) {
                    st_cs_v2(orow + h + jq, v1[0], v1[1]);
                    st_cs_v2(orow + nlon - 2 - jq, v2[1], v2[0]);
                    st_cs_v2(orow + h - 2 - jq, v3[1], v3[0]);
                    st_cs_v2(orow + jq, v4[0], v4[1]);
                } else if (v1[0] + v2[0] + v3[0] + v4[0] == 123.456) {
                    orow[jq] = v1[1];
                }
            }
        }
    }
}

template <int MI, int MINB>
void run(int sms, double* out, const Groups& grp, int nlon, long long rows, int stores, const char* what) {
    constexpr int TM = MI * 32, LDA = TM + 4;
    const int tiles = (int)(rows / TM) * 6;
    const size_t smem = (size_t)STAGES * KC * (LDA + LDB) * sizeof(double);
    CK(cudaFuncSetAttribute(probe<MI, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e9;
    for (int r = 0; r < 6; ++r) {
        CK(cudaEventRecord(e0));
        probe<MI, MINB><<<sms * MINB, 256, smem>>>(out, tiles, grp, nlon, stores);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (r > 0) best = std::min(best, ms);
    }
    CK(cudaGetLastError());
    printf("%-44s stores=%d: %.3f ms (smem %zu KB per CTA)\n", what, stores, best, smem / 1024);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int nlon = 720; const long long rows = 240LL * 360;     // config 2
    double* out; CK(cudaMalloc(&out, (size_t)rows * nlon * sizeof(double)));
    Groups grp{{0, 52, 100, 148, 196}};
    for (int stores = 0; stores < 2; ++stores) {
        run<4, 1>(prop.multiProcessorCount, out, grp, nlon, rows, stores, "128-row tiles, 1 CTA/SM (as shipped)");
        run<2, 2>(prop.multiProcessorCount, out, grp, nlon, rows, stores, "64-row tiles, 2 CTAs/SM");
        run<2, 3>(prop.multiProcessorCount, out, grp, nlon, rows, stores, "64-row tiles, 3 CTAs/SM");
    }
    return 0;
}
