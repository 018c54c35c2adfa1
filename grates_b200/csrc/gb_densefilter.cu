// Dense filter matrices (GeneralMatrix / VDK, reference filter.py:430-572), batched over epochs:
//   Y[r][e] = sum_c W[r][c] x_e[c],   x_e = ravel(anm_e, nmin, nmax)   (degree-wise order, utilities.py:310-360)
// as ONE GEMM on the persistent DMMA kernel of gb_gemm.cuh (2 K^2 flops per epoch; the reference does a
// dgemv per epoch):
//   gb_dense_filter_prepare   W [K][K] row-major -> A tiles [row tile][c][132] (A[k = c][row = r] = W[r][c]), once per filter
//   gb_dense_filter           ravel the batch straight into B tiles [epoch tile][c][124], GEMM, epilogue unravels
//                             into the packed output; degrees below nmin are copied through (filter.py:473-477)
#include "gb_common.cuh"
#include "gb_gemm.cuh"

namespace {

using gbgemm::degreewise_position;

// 32 x 32 transposing tiles: reads rows of W, writes runs of 32 rows r per column c
__global__ void __launch_bounds__(256)
gb_dense_tiles_kernel(const double* __restrict__ W, double* __restrict__ At, long long K, int kp4) {
    __shared__ double s_t[32][33];
    const long long r0 = (long long)blockIdx.y * 32, c0 = (long long)blockIdx.x * 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int rr = w; rr < 32; rr += 8) {
        const long long r = r0 + rr, c = c0 + lane;
        s_t[rr][lane] = (r < K && c < K) ? W[(size_t)r * K + c] : 0.0;
    }
    __syncthreads();
    for (int cc = w; cc < 32; cc += 8) {
        const long long c = c0 + cc, r = r0 + lane;
        if (c < kp4) At[gb_ab_offset(r, (int)c, kp4)] = s_t[lane][cc];
    }
}

// B tiles: Bt[(e / 120)][c][e % 120] = anm[e][position of degree-wise index c + nmin^2], zero beyond the input degree
__global__ void __launch_bounds__(256)
gb_dense_ravel_tiles(const double* __restrict__ anm, double* __restrict__ Bt, int Lin, int nmin, long long K, int kp4, int E) {
    const int e = blockIdx.y * blockDim.x + threadIdx.x;
    const long long c = blockIdx.x;
    if (e >= E) return;
    double v = 0.0;
    if (c < K) {
        int row, col, n;
        degreewise_position(c + (long long)nmin * nmin, row, col, n);
        if (n < Lin) v = anm[((size_t)e * Lin + row) * Lin + col];
    }
    Bt[((size_t)(e / GB_S2_TN) * kp4 + c) * GB_S2_LDB + e % GB_S2_TN] = v;
}

// degrees below nmin pass through: the top-left nmin x nmin corner of the packed array (filter.py:477)
__global__ void gb_dense_passthrough(const double* __restrict__ in, double* __restrict__ out, int Lin, int Lout, int nmin, int E) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= E * nmin * nmin) return;
    const int e = idx / (nmin * nmin), rc = idx % (nmin * nmin), r = rc / nmin, c = rc % nmin;
    out[((size_t)e * Lout + r) * Lout + c] = in[((size_t)e * Lin + r) * Lin + c];
}

}  // namespace

// shared with the point-set synthesis (gb_points.cu): a batch ravelled into GEMM B tiles [epoch tile][c][124]
int gb_launch_ravel_tiles(const double* d_anm, double* d_bt, int Lin, int nmin, long long K, int kp4, int E, cudaStream_t st) {
    dim3 grid((unsigned)K, (E + 255) / 256);
    gb_dense_ravel_tiles<<<grid, 256, 0, st>>>(d_anm, d_bt, Lin, nmin, K, kp4, E);
    GB_LAUNCH_CHECK();
    return GB_OK;
}

extern "C" int64_t gb_dense_filter_tile_elements(int64_t k) {
    const long long kp4 = (k + 3) / 4 * 4;
    return (int64_t)((k + GB_TM - 1) / GB_TM) * kp4 * GB_LDA;
}

extern "C" int gb_dense_filter_prepare(const double* d_matrix, int64_t k, double* d_tiles, int device, void* stream) {
    GB_REQUIRE(k > 0 && k < (1LL << 30), "gb_dense_filter_prepare: matrix size %lld out of range", (long long)k);
    GB_REQUIRE(d_matrix && d_tiles, "gb_dense_filter_prepare: NULL pointer");
    GB_CUDA(cudaSetDevice(device));
    const int kp4 = (int)((k + 3) / 4 * 4);
    const long long rows = (k + GB_TM - 1) / GB_TM * GB_TM;
    dim3 grid((unsigned)((kp4 + 31) / 32), (unsigned)(rows / 32));
    gb_dense_tiles_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_matrix, d_tiles, k, kp4);
    GB_LAUNCH_CHECK();
    return GB_OK;
}

extern "C" int gb_dense_filter(const double* d_tiles, int nmin, int nmax_filter, const double* d_anm_in, int n_epochs,
                               int nmax_in, double* d_anm_out, int device, void* stream) {
    GB_REQUIRE(nmin >= 0 && nmin <= nmax_filter && nmax_in >= 0, "gb_dense_filter: bad degree range");
    GB_REQUIRE(n_epochs >= 0, "gb_dense_filter: n_epochs=%d is negative", n_epochs);
    if (n_epochs == 0) return GB_OK;
    GB_REQUIRE(d_tiles && d_anm_in && d_anm_out, "gb_dense_filter: NULL pointer");
    GB_REQUIRE(d_anm_in != d_anm_out, "gb_dense_filter: input and output may not alias");
    GB_CUDA(cudaSetDevice(device));
    gb_retain_pool_memory(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int E = n_epochs;
    const int Lin = nmax_in + 1;
    const int nmax_out = nmax_in < nmax_filter ? nmax_in : nmax_filter;        // filter.py:469
    const int Lout = nmax_out + 1;
    const long long K = (long long)(nmax_filter + 1) * (nmax_filter + 1) - (long long)nmin * nmin;
    const int kp4 = (int)((K + 3) / 4 * 4);
    const int n_ct = (E + GB_S2_TN - 1) / GB_S2_TN;
    int sm_count = 0;       // cudaGetDeviceProperties would cost milliseconds per call
    GB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    double* d_bt = nullptr;
    const size_t bt_elems = (size_t)n_ct * kp4 * GB_S2_LDB;
    gb_scratch scratch(st);
    GB_CUDA(scratch.alloc(&d_bt, bt_elems));
    GB_CUDA(cudaMemsetAsync(d_bt, 0, bt_elems * sizeof(double), st));
    GB_CUDA(cudaMemsetAsync(d_anm_out, 0, (size_t)E * Lout * Lout * sizeof(double), st));
    {
        int rc = gb_launch_ravel_tiles(d_anm_in, d_bt, Lin, nmin, K, kp4, E, st);
        if (rc) return rc;
    }
    {
        gbgemm::Shape sh;
        sh.A_t = d_tiles;
        sh.a_rows = kp4;
        sh.a_koff_mul = 0;
        sh.tiles_per_group = 1;
        sh.B_t = d_bt;
        sh.b_rows = kp4;
        sh.klen = kp4;
        sh.n_mtiles = (int)((K + GB_TM - 1) / GB_TM);
        sh.n_ntiles = n_ct;
        int rc = gbgemm::launch(sh, gbgemm::UnravelStore<false>{d_anm_out, K, nmin, Lout, E}, sm_count, st);
        if (rc) return rc;
    }
    const int nm = nmin < Lout ? nmin : Lout;
    if (nm > 0) {
        const int total = E * nm * nm;
        // the corner is copied with the smaller of the two strides in mind: rows / columns below nmin exist in both
        gb_dense_passthrough<<<(total + 255) / 256, 256, 0, st>>>(d_anm_in, d_anm_out, Lin, Lout, nm, E);
        GB_LAUNCH_CHECK();
    }
    return GB_OK;
}
