// Covariance propagation to grid-point variances on B200: diag(F Sigma F') for a regular grid.
//
// Replaces the per-parallel products of RegularGrid.covariance_propagation (reference
// grid.py:833-835), F[(i,j), a] = U_i[a] * t_a(j) with U_i[a] = kn[i,n] P_nm(theta_i) and
// t_a(j) = cos/sin(m lon_j).  The reference forms F Sigma F' (nlon x nlon) per parallel and keeps
// its diagonal: 2 P K^2 flops.  On a regular grid F factors, so the same numbers follow from
//
//   H_i[k,k'] = sum_{a in k} sum_{b in k'} U_i[a] Sigma[a,b] U_i[b]        (k = (m, cos|sin) group)
//   var[i,j]  = sum_k T[k,j] * ( sum_k' H_i[k,k'] T[k',j] )
//
// i.e. 2 nlat K^2 + 2 P kpad^2 flops (c4: 83.6 GF instead of 45.9 TF).  This restructuring is
// declared in DESIGN.md and bench.py reports algorithmic and executed flops separately.
//
// Steps (all on the caller's stream):
//   1. gb_cov_permute   Sigma (degree-wise order) -> Sigma' (order-wise: groups contiguous, each
//                       padded to a multiple of 4 rows/cols with zeros)
//   2. gb_cov_legendre  U[p][i]: on-the-fly Legendre recursion * kn for the requested parallels,
//                       rows in the same order-wise padded order
//   3. gb_cov_quadform  H[i][k][k'] for all group pairs
//   4. gb_cov_longitude var[i][j] (optionally sqrt)
#include <vector>
#include <cmath>
#include "gb_common.cuh"

namespace {

template <typename F>
__device__ __forceinline__ void legendre_column(int m, int L, double ct, double pmm, const double* __restrict__ ra,
                                                const double* __restrict__ rb, const double* __restrict__ rc, F&& f) {
    double p2 = pmm;
    f(m, p2);
    if (m + 1 >= L) return;
    double p1 = __dmul_rn(__dmul_rn(rc[m + 1], ct), p2);
    f(m + 1, p1);
    for (int n = m + 2; n < L; ++n) {
        const double p = __dsub_rn(__dmul_rn(__dmul_rn(ra[(size_t)n * L + m], ct), p1),
                                   __dmul_rn(rb[(size_t)n * L + m], p2));
        f(n, p);
        p2 = p1;
        p1 = p;
    }
}

__global__ void __launch_bounds__(256)
gb_cov_permute(const double* __restrict__ sigma, double* __restrict__ sp, const int* __restrict__ perm, int Kp,
               long long K) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= Kp) return;
    const int pr = perm[r], pc = perm[c];
    sp[(size_t)r * Kp + c] = (pr < 0 || pc < 0) ? 0.0 : sigma[(size_t)pr * K + pc];
}

// U[goff[2m+cs] + (n - n0)][i_local] = kn[i][n] * P_nm(theta_i),  n >= n0 = max(m, nmin)
__global__ void __launch_bounds__(128)
gb_cov_legendre(double* __restrict__ U, const int* __restrict__ goff, const double* __restrict__ ct,
                const double* __restrict__ kn, const double* __restrict__ pmm, const double* __restrict__ ra,
                const double* __restrict__ rb, const double* __restrict__ rc, int L, int nmin, int row0, int nrows,
                int ldu) {
    const int il = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y;
    if (il >= nrows) return;
    const int i = row0 + il;
    const int n0 = max(m, nmin);
    const double* kn_i = kn + (size_t)i * L;
    const int gc = goff[2 * m], gs = goff[2 * m + 1];
    legendre_column(m, L, ct[i], pmm[(size_t)i * L + m], ra, rb, rc, [&](int n, double p) {
        if (n < n0) return;
        const double v = __dmul_rn(p, kn_i[n]);
        U[(size_t)(gc + n - n0) * ldu + il] = v;
        if (m > 0) U[(size_t)(gs + n - n0) * ldu + il] = v;
    });
}

// H[i][k][k'] = sum_{r,c} U[gk+r][i] Sp[gk+r][gk'+c] U[gk'+c][i]; CTA per (k', k), threads over parallels.
constexpr int QA = 32;  // rows of the Sigma' block staged per pass

__global__ void __launch_bounds__(128)
gb_cov_quadform(const double* __restrict__ sp, const double* __restrict__ U, double* __restrict__ H,
                const int* __restrict__ goff, const int* __restrict__ gcnt, int Kp, int kpad, int nrows, int ldu) {
    extern __shared__ double s_s[];  // [QA][cntc]
    const int kc = blockIdx.x, kr = blockIdx.y;
    const int cntr = gcnt[kr], cntc = gcnt[kc];
    if (cntr == 0 || cntc == 0) {
        for (int i = threadIdx.x; i < nrows; i += blockDim.x) H[((size_t)i * kpad + kr) * kpad + kc] = 0.0;
        return;
    }
    const int gr = goff[kr], gcol = goff[kc];
    for (int i0 = 0; i0 < nrows; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        const bool live = i < nrows;
        double h = 0.0;
        for (int a0 = 0; a0 < cntr; a0 += QA) {
            const int na = min(QA, cntr - a0);
            __syncthreads();
            for (int idx = threadIdx.x; idx < na * cntc; idx += blockDim.x) {
                const int a = idx / cntc, c = idx % cntc;
                s_s[idx] = sp[(size_t)(gr + a0 + a) * Kp + gcol + c];
            }
            __syncthreads();
            if (live) {
                for (int a = 0; a < na; ++a) {
                    double z0 = 0.0, z1 = 0.0;
                    const double* srow = s_s + a * cntc;
                    int c = 0;
                    for (; c + 1 < cntc; c += 2) {
                        z0 = fma(srow[c], U[(size_t)(gcol + c) * ldu + i], z0);
                        z1 = fma(srow[c + 1], U[(size_t)(gcol + c + 1) * ldu + i], z1);
                    }
                    if (c < cntc) z0 = fma(srow[c], U[(size_t)(gcol + c) * ldu + i], z0);
                    h = fma(U[(size_t)(gr + a0 + a) * ldu + i], z0 + z1, h);
                }
            }
        }
        if (live) H[((size_t)i * kpad + kr) * kpad + kc] = h;
    }
}

// var[i][j] = sum_k T[k][j] * sum_k' H[i][k][k'] T[k'][j]; CTA per (j tile, i)
__global__ void __launch_bounds__(128)
gb_cov_longitude(const double* __restrict__ H, const double* __restrict__ trig, double* __restrict__ out, int kpad,
                 int k_used, int nlon, int nlp, int take_sqrt) {
    extern __shared__ double s_h[];  // one row of H_i at a time: [kpad]
    const int i = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = j < nlon;
    const double* Hi = H + (size_t)i * kpad * kpad;
    double var = 0.0;
    for (int k = 0; k < k_used; ++k) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < k_used; idx += blockDim.x) s_h[idx] = Hi[(size_t)k * kpad + idx];
        __syncthreads();
        if (live) {
            double w0 = 0.0, w1 = 0.0;
            int kk = 0;
            for (; kk + 1 < k_used; kk += 2) {
                w0 = fma(s_h[kk], trig[(size_t)kk * nlp + j], w0);
                w1 = fma(s_h[kk + 1], trig[(size_t)(kk + 1) * nlp + j], w1);
            }
            if (kk < k_used) w0 = fma(s_h[kk], trig[(size_t)kk * nlp + j], w0);
            var = fma(trig[(size_t)k * nlp + j], w0 + w1, var);
        }
    }
    if (live) out[(size_t)i * nlon + j] = take_sqrt ? sqrt(var) : var;
}

}  // namespace

extern "C" int gb_covariance_propagation(gb_plan* plan, const double* d_sigma, int nmin, int row0, int nrows,
                                         double* d_out, int take_sqrt, void* stream) {
    GB_REQUIRE(plan != nullptr, "gb_covariance_propagation: plan is NULL");
    gb_plan* p = plan;
    GB_REQUIRE(nmin >= 0 && nmin <= p->nmax, "gb_covariance_propagation: min_degree=%d outside [0, %d]", nmin, p->nmax);
    GB_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= p->nlat,
               "gb_covariance_propagation: parallels [%d, %d) outside the grid (%d parallels)", row0, row0 + nrows, p->nlat);
    if (nrows == 0) return GB_OK;
    GB_REQUIRE(d_sigma && d_out, "gb_covariance_propagation: NULL device pointer");
    GB_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int L = p->L, kpad = p->kpad;
    const long long K = (long long)L * L - (long long)nmin * nmin;

    // order-wise padded layout: group k = 2m + cs holds degrees n0(m)..nmax
    std::vector<int> goff(kpad + 1, 0), gcnt(kpad, 0);
    int Kp = 0;
    for (int k = 0; k < kpad; ++k) {
        const int m = k >> 1, cs = k & 1;
        int cnt = 0;
        if (m < L && !(m == 0 && cs == 1)) cnt = L - (m > nmin ? m : nmin);
        goff[k] = Kp;
        gcnt[k] = cnt;
        Kp += (cnt + 3) / 4 * 4;
    }
    goff[kpad] = Kp;
    std::vector<int> perm(Kp, -1);
    for (int k = 0; k < kpad; ++k) {
        const int m = k >> 1, cs = k & 1;
        const int n0 = (m > nmin ? m : nmin);
        for (int r = 0; r < gcnt[k]; ++r) {
            const int n = n0 + r;
            const long long idx = (long long)n * n + (m == 0 ? 0 : 2 * m - 1 + cs) - (long long)nmin * nmin;
            perm[goff[k] + r] = (int)idx;
        }
    }
    const int ldu = (nrows + 7) / 8 * 8;
    int *d_goff = nullptr, *d_gcnt = nullptr, *d_perm = nullptr;
    double *d_sp = nullptr, *d_U = nullptr, *d_H = nullptr;
    GB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_goff), (kpad + 1) * sizeof(int), st));
    GB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_gcnt), kpad * sizeof(int), st));
    GB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_perm), Kp * sizeof(int), st));
    GB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_sp), (size_t)Kp * Kp * sizeof(double), st));
    GB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_U), (size_t)Kp * ldu * sizeof(double), st));
    GB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&d_H), (size_t)nrows * kpad * kpad * sizeof(double), st));
    GB_CUDA(cudaMemcpyAsync(d_goff, goff.data(), (kpad + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
    GB_CUDA(cudaMemcpyAsync(d_gcnt, gcnt.data(), kpad * sizeof(int), cudaMemcpyHostToDevice, st));
    GB_CUDA(cudaMemcpyAsync(d_perm, perm.data(), Kp * sizeof(int), cudaMemcpyHostToDevice, st));
    GB_CUDA(cudaStreamSynchronize(st));  // host vectors go out of scope below; copies are tiny
    GB_CUDA(cudaMemsetAsync(d_U, 0, (size_t)Kp * ldu * sizeof(double), st));
    {
        dim3 grid((Kp + 255) / 256, Kp);
        gb_cov_permute<<<grid, 256, 0, st>>>(d_sigma, d_sp, d_perm, Kp, K);
        GB_LAUNCH_CHECK();
    }
    {
        dim3 grid((nrows + 127) / 128, L);
        gb_cov_legendre<<<grid, 128, 0, st>>>(d_U, d_goff, p->d_ct, p->d_kn, p->d_pmm, p->d_ra, p->d_rb, p->d_rc, L,
                                              nmin, row0, nrows, ldu);
        GB_LAUNCH_CHECK();
    }
    {
        dim3 grid(kpad, kpad);
        const size_t smem = (size_t)QA * L * sizeof(double);
        if (smem > 48 * 1024)
            GB_CUDA(cudaFuncSetAttribute(gb_cov_quadform, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gb_cov_quadform<<<grid, 128, smem, st>>>(d_sp, d_U, d_H, d_goff, d_gcnt, Kp, kpad, nrows, ldu);
        GB_LAUNCH_CHECK();
    }
    {
        dim3 grid((p->nlon + 127) / 128, nrows);
        gb_cov_longitude<<<grid, 128, kpad * sizeof(double), st>>>(d_H, p->d_trig, d_out, kpad, 2 * L, p->nlon, p->nlp,
                                                                   take_sqrt);
        GB_LAUNCH_CHECK();
    }
    GB_CUDA(cudaFreeAsync(d_goff, st));
    GB_CUDA(cudaFreeAsync(d_gcnt, st));
    GB_CUDA(cudaFreeAsync(d_perm, st));
    GB_CUDA(cudaFreeAsync(d_sp, st));
    GB_CUDA(cudaFreeAsync(d_U, st));
    GB_CUDA(cudaFreeAsync(d_H, st));
    return GB_OK;
}
