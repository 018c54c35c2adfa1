// Consumers of gridded epoch batches that keep the data on the device (HBM-bound reductions).
//
//   gb_temporal_rms       out[p] = sqrt( sum_e v[e][p]^2 / E )        (reference gravityfield.py:1143-1172:
//                         the epochs are added in order with separate multiply and add, so the
//                         result is bit-identical to the reference's accumulation loop)
//   gb_weighted_moments   per epoch:  sum_p w[p] (v[e][p] - c_e),  sum_p w[p] (v[e][p] - c_e)^2
//                         with w = area element * mask: the sums behind Grid.mean / rms / std
//                         (reference grid.py:174-260); c_e = 0 for mean and rms, the mean for std.
//                         Two deterministic stages (per-block partials, then one block per epoch).
#include "gb_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
gb_temporal_rms_kernel(const double* __restrict__ v, int E, long long P, double* __restrict__ out) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    double s = 0.0;
    for (int e = 0; e < E; ++e) {
        const double x = v[(size_t)e * P + p];
        s = __dadd_rn(s, __dmul_rn(x, x));
    }
    out[p] = sqrt(s / (double)E);
}

constexpr int RB = 64;    // partial sums per epoch

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
    return x;
}

__global__ void __launch_bounds__(256)
gb_moments_partial(const double* __restrict__ v, const double* __restrict__ w, const double* __restrict__ shift, long long P,
                   double* __restrict__ partial) {
    __shared__ double s_a[8], s_b[8];
    const int e = blockIdx.y;
    const double c = shift ? shift[e] : 0.0;
    const double* ve = v + (size_t)e * P;
    double a = 0.0, b = 0.0;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        const double wp = w[p];
        if (wp == 0.0) continue;             // outside the mask: the value does not take part, whatever it is (NaN on land)
        const double d = ve[p] - c;
        a = fma(wp, d, a);
        b = fma(wp * d, d, b);
    }
    a = warp_sum(a);
    b = warp_sum(b);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_a[warp] = a; s_b[warp] = b; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0;
        for (int i = 0; i < 8; ++i) { ta += s_a[i]; tb += s_b[i]; }
        partial[((size_t)e * gridDim.x + blockIdx.x) * 2] = ta;
        partial[((size_t)e * gridDim.x + blockIdx.x) * 2 + 1] = tb;
    }
}

__global__ void gb_moments_finish(const double* __restrict__ partial, int nb, double* __restrict__ out) {
    const int e = blockIdx.x;
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0;
        for (int i = 0; i < nb; ++i) {
            ta += partial[((size_t)e * nb + i) * 2];
            tb += partial[((size_t)e * nb + i) * 2 + 1];
        }
        out[2 * e] = ta;
        out[2 * e + 1] = tb;
    }
}

}  // namespace

extern "C" int gb_temporal_rms(const double* d_values, int n_epochs, int64_t n_points, double* d_out, int device,
                               void* stream) {
    GB_REQUIRE(n_epochs > 0 && n_points >= 0, "gb_temporal_rms: needs at least one epoch");
    if (n_points == 0) return GB_OK;
    GB_REQUIRE(d_values && d_out, "gb_temporal_rms: NULL pointer");
    GB_CUDA(cudaSetDevice(device));
    gb_temporal_rms_kernel<<<(unsigned)((n_points + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        d_values, n_epochs, n_points, d_out);
    GB_LAUNCH_CHECK();
    return GB_OK;
}

extern "C" int gb_weighted_moments(const double* d_values, const double* d_weights, const double* d_shift, int n_epochs,
                                   int64_t n_points, double* d_out, int device, void* stream) {
    GB_REQUIRE(n_epochs >= 0 && n_points >= 0, "gb_weighted_moments: negative size");
    if (n_epochs == 0) return GB_OK;
    GB_REQUIRE(d_values && d_weights && d_out, "gb_weighted_moments: NULL pointer");
    GB_REQUIRE(n_epochs <= 65535, "gb_weighted_moments: at most 65535 epochs per call");
    GB_CUDA(cudaSetDevice(device));
    gb_retain_pool_memory(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double* d_partial = nullptr;
    gb_scratch scratch(st);
    GB_CUDA(scratch.alloc(&d_partial, (size_t)n_epochs * RB * 2));
    dim3 grid(RB, n_epochs);
    gb_moments_partial<<<grid, 256, 0, st>>>(d_values, d_weights, d_shift, n_points, d_partial);
    GB_LAUNCH_CHECK();
    gb_moments_finish<<<n_epochs, 32, 0, st>>>(d_partial, RB, d_out);
    GB_LAUNCH_CHECK();
    return GB_OK;
}

// ---------------------------------------------------------------------------------------------
// Basin (functional) variances  var_b = a_b' Sigma a_b  for a handful of coefficient-space functionals
// a_b (the area-weighted basin mean of a synthesised field is a linear functional of the coefficients,
// a_b = A' w_b; the adjoint synthesis A' w runs through the analysis kernels with adjoint operators).
//   gb_ravel_coefficients   packed [B][L][L] -> degree-wise vectors [B][K'] (reference utilities.py:310-360)
//   gb_quadratic_forms      var_b = sum_r a_b[r] (sum_c Sigma[r][c] a_b[c]): one warp per row of Sigma,
//                           Sigma is read once (HBM-bound: 708 MB at degree 96), the vectors stay in L2
// ---------------------------------------------------------------------------------------------
namespace {

__global__ void __launch_bounds__(256)
gb_ravel_kernel(const double* __restrict__ anm, double* __restrict__ vec, int L, int nmin, long long K, int B) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= K * B) return;
    const int b = (int)(idx / K);
    const long long c = idx - (long long)b * K + (long long)nmin * nmin;     // degree-wise index from degree 0
    const int n = (int)floor(sqrt((double)c));
    int nn = n;
    if ((long long)nn * nn > c) --nn;
    if ((long long)(nn + 1) * (nn + 1) <= c) ++nn;
    const int j = (int)(c - (long long)nn * nn);                              // 0: C_n0, 2m-1: C_nm, 2m: S_nm
    const int m = (j + 1) >> 1;
    const bool sine = j > 0 && (j & 1) == 0;
    const double* a = anm + (size_t)b * L * L;
    vec[idx] = sine ? a[(size_t)(m - 1) * L + nn] : a[(size_t)nn * L + m];
}

constexpr int QF_B = 8;    // functionals per pass

// partial[(b0 + b) * gridDim.x + blockIdx.x]: one writer per (functional, CTA); gb_quadratic_forms_finish adds them in order
__global__ void __launch_bounds__(256)
gb_quadratic_forms_kernel(const double* __restrict__ sigma, const double* __restrict__ vec, long long K, int b0, int nb,
                          double* __restrict__ partial) {
    __shared__ double s_part[8][QF_B];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double tot[QF_B];
#pragma unroll
    for (int b = 0; b < QF_B; ++b) tot[b] = 0.0;
    for (long long r = (long long)blockIdx.x * 8 + warp; r < K; r += (long long)gridDim.x * 8) {
        const double* row = sigma + (size_t)r * K;
        double acc[QF_B];
#pragma unroll
        for (int b = 0; b < QF_B; ++b) acc[b] = 0.0;
        for (long long c = lane; c < K; c += 32) {
            const double s = row[c];
#pragma unroll
            for (int b = 0; b < QF_B; ++b)
                if (b < nb) acc[b] = fma(s, __ldg(vec + (size_t)(b0 + b) * K + c), acc[b]);
        }
#pragma unroll
        for (int b = 0; b < QF_B; ++b) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], off);
            if (b < nb) tot[b] = fma(acc[b], vec[(size_t)(b0 + b) * K + r], tot[b]);
        }
    }
    if (lane == 0)
        for (int b = 0; b < QF_B; ++b) s_part[warp][b] = tot[b];
    __syncthreads();
    if (threadIdx.x < nb) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_part[w][threadIdx.x];
        partial[(size_t)(b0 + threadIdx.x) * gridDim.x + blockIdx.x] = t;
    }
}

__global__ void gb_quadratic_forms_finish(const double* __restrict__ partial, int nparts, int n_vec, double* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_vec) return;
    double t = 0.0;
    for (int i = 0; i < nparts; ++i) t += partial[(size_t)b * nparts + i];
    out[b] = t;
}

}  // namespace

extern "C" int gb_ravel_coefficients(const double* d_anm, int n_sets, int nmax, int nmin, double* d_vec, int device,
                                     void* stream) {
    GB_REQUIRE(n_sets >= 0 && nmax >= 0 && nmin >= 0 && nmin <= nmax, "gb_ravel_coefficients: bad degree range");
    if (n_sets == 0) return GB_OK;
    GB_REQUIRE(d_anm && d_vec, "gb_ravel_coefficients: NULL pointer");
    GB_CUDA(cudaSetDevice(device));
    const int L = nmax + 1;
    const long long K = (long long)L * L - (long long)nmin * nmin;
    const long long total = K * n_sets;
    gb_ravel_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_anm, d_vec, L, nmin,
                                                                                                   K, n_sets);
    GB_LAUNCH_CHECK();
    return GB_OK;
}

extern "C" int gb_quadratic_forms(const double* d_sigma, int64_t k, const double* d_vec, int n_vec, double* d_out,
                                  int device, void* stream) {
    GB_REQUIRE(k >= 0 && n_vec >= 0, "gb_quadratic_forms: negative size");
    if (n_vec == 0) return GB_OK;
    GB_REQUIRE(d_sigma && d_vec && d_out, "gb_quadratic_forms: NULL pointer");
    GB_CUDA(cudaSetDevice(device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (k == 0) {
        GB_CUDA(cudaMemsetAsync(d_out, 0, (size_t)n_vec * sizeof(double), st));
        return GB_OK;
    }
    int sm_count = 0;       // cudaGetDeviceProperties would cost milliseconds per call
    GB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    const int grid = (int)((k + 7) / 8 < (long long)sm_count * 8 ? (k + 7) / 8 : sm_count * 8);
    gb_retain_pool_memory(device);
    gb_scratch scratch(st);
    double* d_partial = nullptr;
    GB_CUDA(scratch.alloc(&d_partial, (size_t)n_vec * grid));
    for (int b0 = 0; b0 < n_vec; b0 += QF_B) {
        const int nb = n_vec - b0 < QF_B ? n_vec - b0 : QF_B;
        gb_quadratic_forms_kernel<<<grid, 256, 0, st>>>(d_sigma, d_vec, k, b0, nb, d_partial);
        GB_LAUNCH_CHECK();
    }
    gb_quadratic_forms_finish<<<(n_vec + 127) / 128, 128, 0, st>>>(d_partial, grid, n_vec, d_out);
    GB_LAUNCH_CHECK();
    return GB_OK;
}
