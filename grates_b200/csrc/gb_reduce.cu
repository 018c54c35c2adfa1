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
    GB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_partial), (size_t)n_epochs * RB * 2 * sizeof(double), st));
    dim3 grid(RB, n_epochs);
    gb_moments_partial<<<grid, 256, 0, st>>>(d_values, d_weights, d_shift, n_points, d_partial);
    GB_LAUNCH_CHECK();
    gb_moments_finish<<<n_epochs, 32, 0, st>>>(d_partial, RB, d_out);
    GB_LAUNCH_CHECK();
    GB_CUDA(cudaFreeAsync(d_partial, st));
    return GB_OK;
}
