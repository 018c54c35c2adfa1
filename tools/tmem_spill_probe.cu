#include <cstdint>
#include <cstdio>
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const double* v) {
    asm volatile(
        "{\n.reg .b32 l<16>, h<16>;\n"
        "mov.b64 {l0,h0}, %1; mov.b64 {l1,h1}, %2; mov.b64 {l2,h2}, %3; mov.b64 {l3,h3}, %4;\n"
        "mov.b64 {l4,h4}, %5; mov.b64 {l5,h5}, %6; mov.b64 {l6,h6}, %7; mov.b64 {l7,h7}, %8;\n"
        "mov.b64 {l8,h8}, %9; mov.b64 {l9,h9}, %10; mov.b64 {l10,h10}, %11; mov.b64 {l11,h11}, %12;\n"
        "mov.b64 {l12,h12}, %13; mov.b64 {l13,h13}, %14; mov.b64 {l14,h14}, %15; mov.b64 {l15,h15}, %16;\n"
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {l0,h0,l1,h1,l2,h2,l3,h3,l4,h4,l5,h5,l6,h6,l7,h7,"
        "l8,h8,l9,h9,l10,h10,l11,h11,l12,h12,l13,h13,l14,h14,l15,h15};\n}\n"
        :: "r"(taddr), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]), "d"(v[4]), "d"(v[5]), "d"(v[6]), "d"(v[7]),
           "d"(v[8]), "d"(v[9]), "d"(v[10]), "d"(v[11]), "d"(v[12]), "d"(v[13]), "d"(v[14]), "d"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, double* v) {
    asm volatile(
        "{\n.reg .b32 l<16>, h<16>;\n"
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {l0,h0,l1,h1,l2,h2,l3,h3,l4,h4,l5,h5,l6,h6,l7,h7,"
        "l8,h8,l9,h9,l10,h10,l11,h11,l12,h12,l13,h13,l14,h14,l15,h15}, [%16];\n"
        "tcgen05.wait::ld.sync.aligned;\n"
        "mov.b64 %0, {l0,h0}; mov.b64 %1, {l1,h1}; mov.b64 %2, {l2,h2}; mov.b64 %3, {l3,h3};\n"
        "mov.b64 %4, {l4,h4}; mov.b64 %5, {l5,h5}; mov.b64 %6, {l6,h6}; mov.b64 %7, {l7,h7};\n"
        "mov.b64 %8, {l8,h8}; mov.b64 %9, {l9,h9}; mov.b64 %10, {l10,h10}; mov.b64 %11, {l11,h11};\n"
        "mov.b64 %12, {l12,h12}; mov.b64 %13, {l13,h13}; mov.b64 %14, {l14,h14}; mov.b64 %15, {l15,h15};\n}\n"
        : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]), "=d"(v[4]), "=d"(v[5]), "=d"(v[6]), "=d"(v[7]),
          "=d"(v[8]), "=d"(v[9]), "=d"(v[10]), "=d"(v[11]), "=d"(v[12]), "=d"(v[13]), "=d"(v[14]), "=d"(v[15])
        : "r"(taddr) : "memory");
}
__global__ void __launch_bounds__(288, 1) k(double* out) {
    __shared__ uint32_t s_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;\n" :: "r"((uint32_t)__cvta_generic_to_shared(&s_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    const uint32_t base = s_base;
    if (warp < 8) {
        const uint32_t taddr = base + ((uint32_t)(32 * (warp & 3)) << 16) + (warp >> 2) * 128;
        double v[16], w[16];
        for (int r = 0; r < 4; ++r) {
            for (int i = 0; i < 16; ++i) v[i] = 1000.0 * threadIdx.x + 16 * r + i + 0.25;
            tmem_st16(taddr + 32 * r, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        for (int r = 0; r < 4; ++r) {
            tmem_ld16(taddr + 32 * r, w);
            for (int i = 0; i < 16; ++i) out[(threadIdx.x * 4 + r) * 16 + i] = w[i];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;\n" :: "r"(base));
}
int main() {
    double* d; cudaMalloc(&d, 256 * 64 * 8);
    k<<<2, 288>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    double* h = new double[256 * 64]; cudaMemcpy(h, d, 256 * 64 * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int t = 0; t < 256; ++t) for (int j = 0; j < 64; ++j) if (h[t * 64 + j] != 1000.0 * t + j + 0.25) ++bad;
    printf("bad=%d sample %f %f\n", bad, h[5 * 64 + 3], h[255 * 64 + 63]);
    return bad != 0;
}
