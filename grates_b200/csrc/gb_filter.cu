// Order-wise block filter, batched over epochs (reference filter.py:180-189).
//
// One CTA per (block g, epoch tile).  Block g = 0 acts on C_n0; g = 2m-1 / 2m on C_nm / S_nm.
// The block's top-left k x k corner (k = nmax+1-m) multiplies the coefficient column of every
// epoch in the tile: y[e][r] = sum_c W[r][c] x[e][c].  A warp owns a row r, its lanes stride over
// the columns c (coalesced reads of W), per-epoch partial sums are combined with warp shuffles.
// HBM-bound: the blocks (9.4 MB at N=120) are read once per epoch tile, coefficients once.
#include "gb_common.cuh"

namespace {

constexpr int FE = 8;  // epochs per CTA

__global__ void __launch_bounds__(256)
gb_orderwise_filter_kernel(const double* __restrict__ blocks, const long long* __restrict__ offsets, int nf,
                           const double* __restrict__ in, double* __restrict__ out, int E, int nmax) {
    extern __shared__ double s_x[];  // [FE][k]
    const int L = nmax + 1;
    const int g = blockIdx.x;                 // 0 .. 2*nmax
    const int m = (g + 1) >> 1;
    const bool sine = (g > 0) && ((g & 1) == 0);
    const int k = L - m;                      // rows/cols used
    const int kf = nf + 1 - m;                // leading dimension of the stored block
    const double* W = blocks + offsets[g];
    const int e0 = blockIdx.y * FE;
    const int ne = min(FE, E - e0);

    // element (n = m + c) of the coefficient column: C_nm = anm[n][m], S_nm = anm[m-1][n]
    auto elem = [&](int c) -> size_t {
        const int n = m + c;
        return sine ? (size_t)(m - 1) * L + n : (size_t)n * L + m;
    };
    for (int idx = threadIdx.x; idx < FE * k; idx += blockDim.x) {
        const int e = idx / k, c = idx % k;
        s_x[idx] = (e < ne) ? in[(size_t)(e0 + e) * L * L + elem(c)] : 0.0;
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int r = warp; r < k; r += nwarps) {
        double acc[FE];
#pragma unroll
        for (int e = 0; e < FE; ++e) acc[e] = 0.0;
        const double* wrow = W + (size_t)r * kf;
        for (int c = lane; c < k; c += 32) {
            const double w = __ldg(wrow + c);
#pragma unroll
            for (int e = 0; e < FE; ++e) acc[e] = fma(w, s_x[e * k + c], acc[e]);
        }
#pragma unroll
        for (int e = 0; e < FE; ++e) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], off);
        }
        if (lane < ne) {
            double v = 0.0;
#pragma unroll
            for (int e = 0; e < FE; ++e)
                if (lane == e) v = acc[e];
            const int n = m + r;
            const size_t pos = (size_t)(e0 + lane) * L * L + elem(r);
            // degrees 0 and 1 pass through unchanged (filter.py:189)
            out[pos] = (n < 2) ? in[pos] : v;
        }
    }
}

}  // namespace

extern "C" int gb_orderwise_filter(const double* d_blocks, const int64_t* block_offsets, int nf, const double* d_anm_in,
                                   int n_epochs, int nmax, double* d_anm_out, int device, void* stream) {
    GB_REQUIRE(nf >= 0 && nmax >= 0, "gb_orderwise_filter: negative degree");
    GB_REQUIRE(nmax <= nf, "gb_orderwise_filter: max_degree=%d exceeds the filter's maximum degree %d", nmax, nf);
    GB_REQUIRE(n_epochs >= 0, "gb_orderwise_filter: n_epochs=%d is negative", n_epochs);
    if (n_epochs == 0) return GB_OK;
    GB_REQUIRE(d_blocks && block_offsets && d_anm_in && d_anm_out, "gb_orderwise_filter: NULL pointer");
    GB_REQUIRE(d_anm_in != d_anm_out, "gb_orderwise_filter: input and output may not alias");
    GB_CUDA(cudaSetDevice(device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nblocks = 2 * nf + 1;
    long long* d_off = nullptr;
    GB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_off), (nblocks + 1) * sizeof(long long), st));
    GB_CUDA(cudaMemcpyAsync(d_off, block_offsets, (nblocks + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    const int L = nmax + 1;
    dim3 grid(2 * nmax + 1, (n_epochs + FE - 1) / FE);
    const size_t smem = (size_t)FE * L * sizeof(double);
    gb_orderwise_filter_kernel<<<grid, 256, smem, st>>>(d_blocks, d_off, nf, d_anm_in, d_anm_out, n_epochs, nmax);
    GB_LAUNCH_CHECK();
    GB_CUDA(cudaFreeAsync(d_off, st));
    return GB_OK;
}
