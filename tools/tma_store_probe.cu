// Throughput of small cp.async.bulk shared -> global stores per SM (development probe for DESIGN.md §9/§10: would a
// TMA-drained epilogue beat st.global for the 128-byte row segments of the Fourier stage?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tma_store_probe tools/tma_store_probe.cu && build/tma_store_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst),
                 "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
                 : "memory");
}

template <int SEG, bool TMA>
__global__ void __launch_bounds__(128) probe(double* out, size_t per_cta_bytes, int iters, int row_stride_bytes) {
    extern __shared__ __align__(128) unsigned char smem[];
    for (int i = threadIdx.x; i < 65536 / 8; i += blockDim.x) reinterpret_cast<double*>(smem)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    unsigned char* base = reinterpret_cast<unsigned char*>(out) + (size_t)blockIdx.x * per_cta_bytes;
    // every thread owns a stream of segments; consecutive segments of a thread are row_stride_bytes apart (scattered rows)
    const int n_seg = (int)(per_cta_bytes / SEG);
    for (int it = 0; it < iters; ++it) {
        for (int s = threadIdx.x; s < n_seg; s += blockDim.x) {
            const size_t off = ((size_t)s * row_stride_bytes) % per_cta_bytes / SEG * SEG;
            const unsigned char* src = smem + (s * SEG) % 65536;
            if (TMA) {
                bulk_s2g(base + off, src, SEG);
                if ((s / blockDim.x) % 8 == 7) {
                    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
                    asm volatile("cp.async.bulk.wait_group.read 4;\n" ::: "memory");
                }
            } else {
                // the LSU path as the kernels use it: 32 bytes per lane, SEG / 32 neighbouring lanes per segment, so one warp
                // instruction writes 1 KB as 1024 / SEG segments of different rows
                constexpr int LPS = SEG / 32;
                const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
                const int s2 = (s / blockDim.x) * blockDim.x * 1 + warp * 32 + lane / LPS + (s / blockDim.x) * 0;
                const size_t off2 = ((size_t)(s2 % n_seg) * row_stride_bytes) % per_cta_bytes / SEG * SEG + (lane % LPS) * 32;
                const double4 v = *reinterpret_cast<const double4*>(src);
                asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};\n" ::"l"(base + off2), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
            }
        }
        if (TMA) {
            asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");
        }
    }
}

template <int SEG, bool TMA>
void run(double* d, size_t per_cta, int stride) {
    cudaFuncSetAttribute(probe<SEG, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 8;
    probe<SEG, TMA><<<148, 128, 65536>>>(d, per_cta, 1, stride);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    probe<SEG, TMA><<<148, 128, 65536>>>(d, per_cta, iters, stride);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double bytes = (double)per_cta * 148 * iters * (TMA ? 1.0 : 32.0 / SEG);   // a thread of the LSU variant writes 32 bytes per step
    printf("%s seg %5d B, stride %6d: %8.1f GB/s total, %6.1f B/clk/SM (1.965 GHz), %6.2f ns per segment per SM  [%s]\n",
           TMA ? "bulk S2G " : "st.cs.v4 ", SEG, stride, bytes / ms / 1e6, bytes / 148 / (ms * 1e-3) / 1.965e9,
           ms * 1e6 / ((double)per_cta / SEG * iters), cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const size_t per_cta = 4u << 20;       // 4 MB per CTA, 592 MB in all (larger than L2)
    double* d;
    cudaMalloc(&d, per_cta * 148);
    for (int stride : {5760, 128}) {       // 5760: one segment per output row of the 0.5 degree grid; 128: contiguous
        run<128, true>(d, per_cta, stride);
        run<256, true>(d, per_cta, stride);
        run<512, true>(d, per_cta, stride);
        run<1024, true>(d, per_cta, stride);
        run<128, false>(d, per_cta, stride);
        run<256, false>(d, per_cta, stride);
    }
    return 0;
}
