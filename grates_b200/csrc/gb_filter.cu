// Order-wise block filter, batched over epochs (reference filter.py:180-189).
//
// Block g = 0 acts on C_n0; g = 2m-1 / 2m on C_nm / S_nm.  The block's top-left k x k corner
// (k = nmax+1-m) multiplies the coefficient column of every epoch: y[r][e] = sum_c W[r][c] x[c][e].
// In the order-wise packed layout of gb_pack.cu the columns of all epochs of one (order, cos|sin)
// form a dense k x E matrix with the epochs contiguous, so the filter is one small GEMM per block:
//
//   gb_pack_kernel           anm [E][L][L] -> X            (HBM-bound transpose)
//   gb_filter_blocks_kernel  Y_g = W_g X_g                 one CTA per (block g, 32 epochs): the X tile
//                            stays in shared memory, W streams through it in 16-column chunks
//                            (register-staged double buffer); both tiles are k-major with pitches
//                            = 4 mod 16, i.e. conflict-free DMMA.8x8x4 operands: 8 warps x (16 rows
//                            x 32 epochs) of FP64 tensor-core tiles, three CTAs per SM
//   gb_pack_kernel<false>    Y -> anm [E][L][L]
//
// Bytes: anm in + out once each (2 x 8 E L^2), X/Y once each way, blocks once per epoch tile
// (L2-resident).  Flops 2 k^2 per block and epoch (2.36 MFLOP per epoch at N = 120).
#include "gb_common.cuh"

namespace {

constexpr int FT_E = 32;            // epochs per CTA
constexpr int FT_R = 128;           // rows per pass
constexpr int FT_KC = 16;           // columns of W per chunk
constexpr int FT_LDX = FT_E + 4;    // pitch of the X tile
constexpr int FT_LDW = FT_R + 4;    // pitch of a (transposed) W chunk

// TILED: Y is written in the layout of the table-fed Legendre stage of the synthesis (gb_pack.cu, x_tiled_position):
// [order][column tile of tn][degree row, even | odd n - m per 8][tn + 4] -- filter + synthesis then needs neither the
// unpack nor a second pack.  The buffer was cleared for this layout (the sine plane of order 0 stays zero).
struct FilterTiling {
    const int* roff;        // [L + 1] first row of every order
    int tn, n_ct;
};
template <bool TILED>
__global__ void __launch_bounds__(256, 3)
gb_filter_blocks_kernel(const double* __restrict__ blocks, const long long* __restrict__ offsets, int nf,
                        const double* __restrict__ X, double* __restrict__ Y, int E, int nmax, FilterTiling ft) {
    extern __shared__ __align__(16) double s_f[];
    const int L = nmax + 1;
    const int g = blockIdx.x;                 // 0 .. 2*nmax
    const int m = (g + 1) >> 1;
    const int cs = (g > 0) && ((g & 1) == 0);
    const int k = L - m;                      // rows/cols used
    const int kf = nf + 1 - m;                // leading dimension of the stored block
    const int kp = (k + FT_KC - 1) / FT_KC * FT_KC;
    const double* W = blocks + offsets[g];
    const int e0 = blockIdx.y * FT_E;
    const int ne = min(FT_E, E - e0);
    const long long xo = 2LL * E * ((long long)m * L - (long long)m * (m - 1) / 2) + (long long)cs * E + e0;
    const double* Xg = X + xo;                // element (c, e) at Xg[c * 2E + e]
    double* Yg = Y + xo;

    double* s_x = s_f;                        // [kp][FT_LDX]
    double* s_w = s_f + (size_t)kp * FT_LDX;  // [2][FT_KC][FT_LDW]
    const int tid = threadIdx.x;

    // coefficient tile: rows beyond k and epochs beyond E are zero
    for (int idx = tid; idx < kp * FT_E; idx += 256) {
        const int c = idx / FT_E, e = idx % FT_E;
        s_x[c * FT_LDX + e] = (c < k && e < ne) ? Xg[(size_t)c * 2 * E + e] : 0.0;
    }

    const int warp = tid >> 5, lane = tid & 31, fg = lane >> 2, q = lane & 3;   // fragment row / k index
    const int lr = tid >> 1, lc = (tid & 1) * 8;   // W chunk loader: row lr, columns lc .. lc+7
    const int n_chunks = kp / FT_KC;

    for (int r0 = 0; r0 < k; r0 += FT_R) {
        double acc[2][4][2];                           // warp tile: rows 16 warp .. +15, all 32 epochs
#pragma unroll
        for (int mi = 0; mi < 2; ++mi)
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
        const bool active = r0 + 16 * warp < k;        // warp-uniform
        double wreg[8];
        auto load_w = [&](int chunk) {
            const int r = r0 + lr;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = chunk * FT_KC + lc + j;
                wreg[j] = (r < k && c < k) ? __ldg(W + (size_t)r * kf + c) : 0.0;
            }
        };
        auto store_w = [&](int buf) {
#pragma unroll
            for (int j = 0; j < 8; ++j) s_w[(buf * FT_KC + lc + j) * FT_LDW + lr] = wreg[j];
        };
        load_w(0);
        __syncthreads();                               // previous pass finished with s_w; s_x is complete
        store_w(0);
        __syncthreads();
        for (int chunk = 0; chunk < n_chunks; ++chunk) {
            const int buf = chunk & 1;
            if (chunk + 1 < n_chunks) load_w(chunk + 1);
            if (active) {
                const double* wp = s_w + (size_t)buf * FT_KC * FT_LDW + 16 * warp + fg;
                const double* xp = s_x + (size_t)chunk * FT_KC * FT_LDX + fg;
#pragma unroll
                for (int kk = 0; kk < FT_KC; kk += 4) {
                    double a[2], b[4];
#pragma unroll
                    for (int mi = 0; mi < 2; ++mi) a[mi] = wp[(kk + q) * FT_LDW + mi * 8];
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) b[ni] = xp[(kk + q) * FT_LDX + ni * 8];
#pragma unroll
                    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                        for (int ni = 0; ni < 4; ++ni) gb::dmma_884(acc[mi][ni][0], acc[mi][ni][1], a[mi], b[ni]);
                }
            }
            if (chunk + 1 < n_chunks) {
                store_w(buf ^ 1);                      // the other buffer was last read before the previous barrier
                __syncthreads();
            }
        }
        if (active) {
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                const int r = r0 + 16 * warp + mi * 8 + fg;
                if (r >= k) continue;
                const bool pass = (m + r) < 2;         // degrees 0 and 1 pass through unchanged (filter.py:189)
                long long trow = 0;
                int kn_pad = 0;
                if (TILED) {
                    const int t0 = ft.roff[m];
                    kn_pad = ft.roff[m + 1] - t0;
                    trow = (long long)t0 * ft.n_ct + ((r & ~7) + ((r & 1) << 2) + ((r & 7) >> 1));
                }
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int e = ni * 8 + 2 * q + j;
                        if (e >= ne) continue;
                        const double v = pass ? s_x[r * FT_LDX + e] : acc[mi][ni][j];
                        if (TILED) {
                            const int col = cs * E + e0 + e, ct = col / ft.tn;
                            Y[(trow + (long long)ct * kn_pad) * (ft.tn + 4) + (col - ct * ft.tn)] = v;
                        } else {
                            Yg[(size_t)r * 2 * E + e] = v;
                        }
                    }
            }
        }
    }
    // order 0 has no sine coefficients: keep that plane of Y defined (zero)
    if (!TILED && g == 0) {
        for (int idx = tid; idx < k * FT_E; idx += 256) {
            const int c = idx / FT_E, e = idx % FT_E;
            if (e < ne) Y[(size_t)c * 2 * E + E + e0 + e] = 0.0;
        }
    }
}

}  // namespace

// pack -> filter; the filtered coefficients land in d_y, order-wise packed (roff == nullptr) or in the tiled layout
static int filter_packed(const double* d_blocks, const int64_t* block_offsets, int nf, const double* d_anm_in, int E, int nmax,
                         double* d_y, const int* d_roff, int tn, int n_ct, gb_scratch& scratch, cudaStream_t st) {
    const int nblocks = 2 * nf + 1;
    const int L = nmax + 1;
    const size_t x_elems = (size_t)L * (L + 1) * (size_t)E;      // 2E doubles per (order, degree) pair
    const int kp_max = (L + FT_KC - 1) / FT_KC * FT_KC;
    const size_t smem = ((size_t)kp_max * FT_LDX + 2 * FT_KC * FT_LDW) * sizeof(double);
    GB_REQUIRE(smem <= 227 * 1024, "gb_orderwise_filter: max_degree=%d exceeds the shared-memory tile of the filter kernel", nmax);
    long long* d_off = nullptr;
    double* d_x = nullptr;
    GB_CUDA(scratch.alloc(&d_off, (size_t)nblocks + 1));
    GB_CUDA(scratch.alloc(&d_x, x_elems));
    GB_CUDA(cudaMemcpyAsync(d_off, block_offsets, (nblocks + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
    int rc = gb_launch_pack(d_anm_in, d_x, L, E, st);
    if (rc) return rc;
    const FilterTiling ft{d_roff, tn, n_ct};
    dim3 grid(2 * nmax + 1, (E + FT_E - 1) / FT_E);
    if (d_roff) {
        if (smem > 48 * 1024)
            GB_CUDA(cudaFuncSetAttribute(gb_filter_blocks_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gb_filter_blocks_kernel<true><<<grid, 256, smem, st>>>(d_blocks, d_off, nf, d_x, d_y, E, nmax, ft);
    } else {
        if (smem > 48 * 1024)
            GB_CUDA(cudaFuncSetAttribute(gb_filter_blocks_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gb_filter_blocks_kernel<false><<<grid, 256, smem, st>>>(d_blocks, d_off, nf, d_x, d_y, E, nmax, ft);
    }
    GB_LAUNCH_CHECK();
    return GB_OK;
}

// Filtered coefficients straight into the synthesis workspace X of a plan (gb_synthesis.cu: launch_synthesis).
int gb_filter_into_x(const double* d_blocks, const int64_t* block_offsets, int nf, const double* d_anm, int E, int nmax,
                     double* d_x, const int* d_roff, int tn, int n_ct, cudaStream_t st) {
    GB_REQUIRE(nmax <= nf, "gb_synthesis_orderwise_filtered: max_degree=%d exceeds the filter's maximum degree %d", nmax, nf);
    gb_scratch scratch(st);
    return filter_packed(d_blocks, block_offsets, nf, d_anm, E, nmax, d_x, d_roff, tn, n_ct, scratch, st);
}

extern "C" int gb_orderwise_filter(const double* d_blocks, const int64_t* block_offsets, int nf, const double* d_anm_in,
                                   int n_epochs, int nmax, double* d_anm_out, int device, void* stream) {
    GB_REQUIRE(nf >= 0 && nmax >= 0, "gb_orderwise_filter: negative degree");
    GB_REQUIRE(nmax <= nf, "gb_orderwise_filter: max_degree=%d exceeds the filter's maximum degree %d", nmax, nf);
    GB_REQUIRE(n_epochs >= 0, "gb_orderwise_filter: n_epochs=%d is negative", n_epochs);
    if (n_epochs == 0) return GB_OK;
    GB_REQUIRE(d_blocks && block_offsets && d_anm_in && d_anm_out, "gb_orderwise_filter: NULL pointer");
    GB_REQUIRE(d_anm_in != d_anm_out, "gb_orderwise_filter: input and output may not alias");
    GB_CUDA(cudaSetDevice(device));
    gb_retain_pool_memory(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int L = nmax + 1;
    const int E = n_epochs;
    const size_t x_elems = (size_t)L * (L + 1) * (size_t)E;
    double* d_y = nullptr;
    gb_scratch scratch(st);
    GB_CUDA(scratch.alloc(&d_y, x_elems));
    int rc = filter_packed(d_blocks, block_offsets, nf, d_anm_in, E, nmax, d_y, nullptr, 0, 0, scratch, st);
    if (!rc) rc = gb_launch_unpack(d_y, d_anm_out, L, E, st);
    return rc;
}
