// Explicit operators of a regular grid in degree-wise column / row order:
//   gb_synthesis_matrix   A[(i,j)][a] = kn[i,n] P_nm(theta_i) trig_a(lon_j)
//                         (reference Grid.synthesis_matrix grid.py:412-443 over
//                          RegularGrid.synthesis_matrix_per_order grid.py:627-663)
//   gb_analysis_matrix    F[a][(i,j)] = lat_op_m[n - n0][i] * lon_op[(m,cs)][j]
//                         (reference RegularGrid.analysis_matrix grid.py:698-730; the Kronecker form of
//                          the per-order solve(A'WA, A'W), see gb_analysis.cu)
// Dense O(P K) outputs: meant for small grids / window matrices, as in the reference.
#include "gb_common.cuh"

namespace {

using gb::legendre_column;

// CTA = (meridian tile of 128, parallel i, order m); the Legendre column is recomputed per thread
// (cheap: < 2N dependent operations) so that the stores along j are coalesced per coefficient row.
__global__ void __launch_bounds__(128)
gb_synthesis_matrix_kernel(double* __restrict__ A, const double* __restrict__ ct, const double* __restrict__ kn,
                           const double* __restrict__ pmm, const double* __restrict__ ra, const double* __restrict__ rb,
                           const double* __restrict__ rc, const double* __restrict__ trig, int nlp, int L, int nmin,
                           int nlon, long long K) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y, m = blockIdx.z;
    if (j >= nlon) return;
    const double cm = trig[(size_t)(2 * m) * nlp + j];
    const double sm = trig[(size_t)(2 * m + 1) * nlp + j];
    const double* kn_i = kn + (size_t)i * L;
    double* row = A + ((size_t)i * nlon + j) * K;
    const long long off = (long long)nmin * nmin;
    legendre_column(m, L, ct[i], pmm[(size_t)i * L + m], ra, rb, rc, [&](int n, double pn) {
        if (n < nmin) return;
        const double pk = __dmul_rn(pn, kn_i[n]);
        const long long a = (long long)n * n + (m == 0 ? 0 : 2 * m - 1) - off;
        row[a] = __dmul_rn(pk, cm);
        if (m > 0) row[a + 1] = __dmul_rn(pk, sm);
    });
}

__global__ void __launch_bounds__(256)
gb_analysis_matrix_kernel(double* __restrict__ F, const double* __restrict__ lonT, const double* __restrict__ lat_ops,
                          const long long* __restrict__ lat_off, int L, int nmin, int nlat, int nlon, int kpad) {
    // blockIdx.y = coefficient row a (degree-wise), threads over grid points
    const long long a = blockIdx.y;
    const long long full = a + (long long)nmin * nmin;
    int n = (int)sqrt((double)full);
    while ((long long)(n + 1) * (n + 1) <= full) ++n;
    while ((long long)n * n > full) --n;
    const int r = (int)(full - (long long)n * n);
    const int m = (r + 1) >> 1;
    const int cs = (r > 0 && (r & 1) == 0) ? 1 : 0;
    const int n0 = max(m, nmin);
    const double* op = lat_ops + lat_off[m] + (size_t)(n - n0) * nlat;
    const long long P = (long long)nlat * nlon;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(p / nlon), j = (int)(p % nlon);
        F[a * P + p] = op[i] * lonT[(size_t)j * kpad + 2 * m + cs];
    }
}

}  // namespace

extern "C" int gb_synthesis_matrix(gb_plan* p, int nmin, double* d_out, void* stream) {
    GB_REQUIRE(p != nullptr && d_out != nullptr, "gb_synthesis_matrix: NULL argument");
    GB_REQUIRE(nmin >= 0 && nmin <= p->nmax, "gb_synthesis_matrix: min_degree=%d outside [0, %d]", nmin, p->nmax);
    GB_CUDA(cudaSetDevice(p->device));
    const long long K = (long long)p->L * p->L - (long long)nmin * nmin;
    dim3 grid((p->nlon + 127) / 128, p->nlat, p->L);
    GB_REQUIRE(p->nlat <= 65535 && p->L <= 65535, "gb_synthesis_matrix: grid too large for a dense operator");
    gb_synthesis_matrix_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        d_out, p->d_ct, p->d_kn, p->d_pmm, p->d_ra, p->d_rb, p->d_rc, p->d_trig, p->nlp, p->L, nmin, p->nlon, K);
    GB_LAUNCH_CHECK();
    return GB_OK;
}

extern "C" int gb_analysis_matrix(gb_plan* p, double* d_out, void* stream) {
    GB_REQUIRE(p != nullptr && d_out != nullptr, "gb_analysis_matrix: NULL argument");
    GB_REQUIRE(p->ana_nmin >= 0, "gb_analysis_matrix: gb_plan_set_analysis has not been called for this plan");
    GB_CUDA(cudaSetDevice(p->device));
    const long long K = (long long)p->L * p->L - (long long)p->ana_nmin * p->ana_nmin;
    GB_REQUIRE(K <= 65535, "gb_analysis_matrix: %lld coefficients are too many for a dense operator", K);
    const long long P = (long long)p->nlat * p->nlon;
    dim3 grid((unsigned)((P + 255) / 256 < 1024 ? (P + 255) / 256 : 1024), (unsigned)K);
    gb_analysis_matrix_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        d_out, p->d_lon_ops, p->d_lat_ops, p->d_lat_off, p->L, p->ana_nmin, p->nlat, p->nlon, p->kpad);
    GB_LAUNCH_CHECK();
    return GB_OK;
}
