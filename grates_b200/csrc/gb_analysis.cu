// Spherical-harmonic analysis on B200, epoch-batched.
//
// Replaces RegularGrid.to_potential_coefficients (reference grid.py:776-785), whose per-order
// operator solve(A'WA, A'W) (grid.py:690-696) factors on a regular grid with rank-one area
// weights W_ij = w_i u_j into
//
//   G[e,i,(m,cs)] = sum_j V[e,i,j] * lon_op[(m,cs)][j]                         (longitude stage)
//   x[e,n,(m,cs)] = sum_i lat_op_m[n - n0(m)][i] * G[e,i,(m,cs)]              (latitude stage)
//
// with lon_op = u_j trig(m lon_j) / sum_j u_j trig^2 and lat_op_m = solve(P'WP, P'W), built once
// per plan on the host (grates_b200/plan.py: analysis_operators).
//
// HBM layout:  V [E][nlat][nlon] (input), lonT [nlp][kpad] (lon_op transposed, zero padded),
//              G [kpad][mpad] (same spectral layout as the synthesis intermediate AB),
//              lat_ops concatenated [cnt_m][nlat] blocks, anm [E][L][L] packed (output).
#include <vector>
#include "gb_common.cuh"

namespace {

// ---- longitude stage: G[k][row] = sum_j V[row][j] * lonT[j][k] ;  64 x 64 tile, 4 x 4 per thread
constexpr int LT = 64, LK = 16;

__global__ void __launch_bounds__(256)
gb_analysis_lon_kernel(const double* __restrict__ V, const double* __restrict__ lonT, double* __restrict__ G,
                       long long M, int nlon, int kpad, long long mpad) {
    __shared__ double sV[LK][LT + 1];   // [j][row]
    __shared__ double sT[LK][LT];       // [j][k]
    const long long row0 = (long long)blockIdx.x * LT;
    const int k0 = blockIdx.y * LT;
    const int tx = threadIdx.x & 15;    // k direction
    const int ty = threadIdx.x >> 4;    // row direction
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;

    for (int j0 = 0; j0 < nlon; j0 += LK) {
        // V tile: 64 rows x 16 j, read with j fastest
        for (int idx = threadIdx.x; idx < LT * LK; idx += 256) {
            const int r = idx / LK, jj = idx % LK;
            const long long row = row0 + r;
            const int j = j0 + jj;
            sV[jj][r] = (row < M && j < nlon) ? V[(size_t)row * nlon + j] : 0.0;
        }
        for (int idx = threadIdx.x; idx < LK * LT; idx += 256) {
            const int jj = idx / LT, kk = idx % LT;
            const int j = j0 + jj, k = k0 + kk;
            sT[jj][kk] = (j < nlon && k < kpad) ? lonT[(size_t)j * kpad + k] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int jj = 0; jj < LK; ++jj) {
            double v[4], t[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) v[a] = sV[jj][ty * 4 + a];
#pragma unroll
            for (int b = 0; b < 4; ++b) t[b] = sT[jj][tx * 4 + b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fma(v[a], t[b], acc[a][b]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const int k = k0 + tx * 4 + b;
        if (k >= kpad) continue;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const long long row = row0 + ty * 4 + a;
            if (row < M) G[(size_t)k * mpad + row] = acc[a][b];
        }
    }
}

// ---- latitude stage: one CTA per (order m, epoch tile); a warp owns an output degree, lanes stride
//      over the parallels, sums are combined with warp shuffles.
constexpr int AE = 4;  // epochs per CTA

__global__ void __launch_bounds__(256)
gb_analysis_lat_kernel(const double* __restrict__ G, const double* __restrict__ lat_ops,
                       const long long* __restrict__ lat_off, double* __restrict__ anm, int L, int nlat, int E,
                       int nmin, long long mpad) {
    extern __shared__ double s_g[];  // [2][AE][nlat]
    const int m = blockIdx.x;
    const int e0 = blockIdx.y * AE;
    const int ne = min(AE, E - e0);
    const int n0 = max(m, nmin);
    const int cnt = L - n0;
    if (cnt <= 0) return;
    const double* op = lat_ops + lat_off[m];
    for (int idx = threadIdx.x; idx < 2 * AE * nlat; idx += blockDim.x) {
        const int cs = idx / (AE * nlat);
        const int rem = idx % (AE * nlat);
        const int e = rem / nlat, i = rem % nlat;
        s_g[idx] = (e < ne) ? G[(size_t)(2 * m + cs) * mpad + (size_t)(e0 + e) * nlat + i] : 0.0;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int r = warp; r < cnt; r += nwarps) {
        double acc[2][AE];
#pragma unroll
        for (int cs = 0; cs < 2; ++cs)
#pragma unroll
            for (int e = 0; e < AE; ++e) acc[cs][e] = 0.0;
        const double* row = op + (size_t)r * nlat;
        for (int i = lane; i < nlat; i += 32) {
            const double w = __ldg(row + i);
#pragma unroll
            for (int cs = 0; cs < 2; ++cs)
#pragma unroll
                for (int e = 0; e < AE; ++e) acc[cs][e] = fma(w, s_g[(cs * AE + e) * nlat + i], acc[cs][e]);
        }
#pragma unroll
        for (int cs = 0; cs < 2; ++cs)
#pragma unroll
            for (int e = 0; e < AE; ++e)
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) acc[cs][e] += __shfl_xor_sync(0xffffffffu, acc[cs][e], off);
        if (lane == 0) {
            const int n = n0 + r;
#pragma unroll
            for (int e = 0; e < AE; ++e) {
                if (e >= ne) break;
                double* a = anm + (size_t)(e0 + e) * L * L;
                a[(size_t)n * L + m] = acc[0][e];
                if (m > 0) a[(size_t)(m - 1) * L + n] = acc[1][e];
            }
        }
    }
}

}  // namespace

extern "C" int gb_plan_set_analysis(gb_plan* plan, int nmin, const double* lon_ops, const double* lat_ops,
                                    const int64_t* lat_op_offsets) {
    GB_REQUIRE(plan != nullptr, "gb_plan_set_analysis: plan is NULL");
    GB_REQUIRE(nmin >= 0 && nmin <= plan->nmax, "gb_plan_set_analysis: min_degree=%d outside [0, %d]", nmin, plan->nmax);
    GB_REQUIRE(lon_ops && lat_ops && lat_op_offsets, "gb_plan_set_analysis: NULL pointer");
    gb_plan* p = plan;
    GB_CUDA(cudaSetDevice(p->device));
    GB_CUDA(cudaDeviceSynchronize());
    cudaFree(p->d_lon_ops); p->d_lon_ops = nullptr;
    cudaFree(p->d_lat_ops); p->d_lat_ops = nullptr;
    cudaFree(p->d_lat_off); p->d_lat_off = nullptr;
    delete[] p->h_lat_off; p->h_lat_off = nullptr;
    p->ana_nmin = -1;
    const int L = p->L;
    // transposed + zero padded longitude operator: lonT[j][k], k = 2m + cs
    std::vector<double> lonT((size_t)p->nlp * p->kpad, 0.0);
    for (int k = 0; k < 2 * L; ++k)
        for (int j = 0; j < p->nlon; ++j) lonT[(size_t)j * p->kpad + k] = lon_ops[(size_t)k * p->nlon + j];
    GB_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->d_lon_ops), lonT.size() * sizeof(double)));
    GB_CUDA(cudaMemcpy(p->d_lon_ops, lonT.data(), lonT.size() * sizeof(double), cudaMemcpyHostToDevice));
    p->h_lat_off = new long long[L + 1];
    for (int m = 0; m <= L; ++m) p->h_lat_off[m] = lat_op_offsets[m];
    for (int m = 0; m < L; ++m) {
        const long long expect = (long long)(L - (m > nmin ? m : nmin)) * p->nlat;
        GB_REQUIRE(p->h_lat_off[m + 1] - p->h_lat_off[m] == (expect > 0 ? expect : 0),
                   "gb_plan_set_analysis: operator of order %d has %lld elements, expected %lld", m,
                   p->h_lat_off[m + 1] - p->h_lat_off[m], expect);
    }
    const size_t total = (size_t)p->h_lat_off[L];
    GB_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->d_lat_ops), (total ? total : 1) * sizeof(double)));
    GB_CUDA(cudaMemcpy(p->d_lat_ops, lat_ops, total * sizeof(double), cudaMemcpyHostToDevice));
    GB_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->d_lat_off), (L + 1) * sizeof(long long)));
    GB_CUDA(cudaMemcpy(p->d_lat_off, p->h_lat_off, (L + 1) * sizeof(long long), cudaMemcpyHostToDevice));
    p->ana_nmin = nmin;
    return GB_OK;
}

static int launch_analysis(gb_plan* p, const double* d_grid, int E, double* d_anm, cudaStream_t st) {
    const int L = p->L;
    const long long M = (long long)E * p->nlat;
    const long long mpad = p->ws_mpad;
    GB_CUDA(cudaMemsetAsync(d_anm, 0, (size_t)E * L * L * sizeof(double), st));
    {
        dim3 grid((unsigned)((M + LT - 1) / LT), (p->kpad + LT - 1) / LT);
        gb_analysis_lon_kernel<<<grid, 256, 0, st>>>(d_grid, p->d_lon_ops, p->d_ab, M, p->nlon, p->kpad, mpad);
        GB_LAUNCH_CHECK();
    }
    {
        dim3 grid(L, (E + AE - 1) / AE);
        const size_t smem = (size_t)2 * AE * p->nlat * sizeof(double);
        if (smem > 48 * 1024)
            GB_CUDA(cudaFuncSetAttribute(gb_analysis_lat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        gb_analysis_lat_kernel<<<grid, 256, smem, st>>>(p->d_ab, p->d_lat_ops, p->d_lat_off, d_anm, L, p->nlat, E,
                                                        p->ana_nmin, mpad);
        GB_LAUNCH_CHECK();
    }
    return GB_OK;
}

extern "C" int gb_analysis(gb_plan* plan, const double* d_grid, int n_epochs, double* d_anm, void* stream) {
    GB_REQUIRE(plan != nullptr, "gb_analysis: plan is NULL");
    GB_REQUIRE(plan->ana_nmin >= 0, "gb_analysis: gb_plan_set_analysis has not been called for this plan");
    GB_REQUIRE(n_epochs >= 0, "gb_analysis: n_epochs=%d is negative", n_epochs);
    if (n_epochs == 0) return GB_OK;
    GB_REQUIRE(d_grid && d_anm, "gb_analysis: NULL device pointer");
    GB_CUDA(cudaSetDevice(plan->device));
    int rc = gb_plan_ensure_workspace(plan, n_epochs);
    if (rc) return rc;
    return launch_analysis(plan, d_grid, n_epochs, d_anm, static_cast<cudaStream_t>(stream));
}

extern "C" int gb_analysis_host(gb_plan* plan, const double* h_grid, int n_epochs, double* h_anm) {
    GB_REQUIRE(plan != nullptr, "gb_analysis_host: plan is NULL");
    GB_REQUIRE(plan->ana_nmin >= 0, "gb_analysis_host: gb_plan_set_analysis has not been called for this plan");
    GB_REQUIRE(n_epochs >= 0, "gb_analysis_host: n_epochs=%d is negative", n_epochs);
    if (n_epochs == 0) return GB_OK;
    GB_REQUIRE(h_grid && h_anm, "gb_analysis_host: NULL host pointer");
    gb_plan* p = plan;
    GB_CUDA(cudaSetDevice(p->device));
    const size_t pts = (size_t)p->nlat * p->nlon, coef = (size_t)p->L * p->L;
    // epoch chunks: the upload of chunk c+1 overlaps the kernels of chunk c
    int chunk = (n_epochs + 7) / 8;
    if (chunk < 1) chunk = 1;
    int rc = gb_plan_ensure_workspace(p, chunk);
    if (rc) return rc;
    if (!p->s_compute) GB_CUDA(cudaStreamCreateWithFlags(&p->s_compute, cudaStreamNonBlocking));
    if (!p->s_copy) GB_CUDA(cudaStreamCreateWithFlags(&p->s_copy, cudaStreamNonBlocking));
    for (auto& e : p->ev)
        if (!e) GB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    double* d_in[2] = {nullptr, nullptr};
    double* d_out = nullptr;
    for (int b = 0; b < 2; ++b) GB_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_in[b]), (size_t)chunk * pts * sizeof(double)));
    GB_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_out), (size_t)n_epochs * coef * sizeof(double)));
    int c = 0;
    for (int e0 = 0; e0 < n_epochs; e0 += chunk, ++c) {
        const int ne = (n_epochs - e0 < chunk) ? (n_epochs - e0) : chunk;
        const int b = c & 1;
        if (c >= 2) GB_CUDA(cudaStreamWaitEvent(p->s_copy, p->ev[2 + b], 0));  // kernels done with buffer b
        GB_CUDA(cudaMemcpyAsync(d_in[b], h_grid + (size_t)e0 * pts, (size_t)ne * pts * sizeof(double),
                                cudaMemcpyHostToDevice, p->s_copy));
        GB_CUDA(cudaEventRecord(p->ev[b], p->s_copy));
        GB_CUDA(cudaStreamWaitEvent(p->s_compute, p->ev[b], 0));
        if ((rc = launch_analysis(p, d_in[b], ne, d_out + (size_t)e0 * coef, p->s_compute))) break;
        GB_CUDA(cudaEventRecord(p->ev[2 + b], p->s_compute));
    }
    if (rc == GB_OK) {
        cudaError_t e = cudaMemcpyAsync(h_anm, d_out, (size_t)n_epochs * coef * sizeof(double), cudaMemcpyDeviceToHost,
                                        p->s_compute);
        if (e != cudaSuccess) rc = gb_set_error(GB_ERR_CUDA, "gb_analysis_host: result copy failed: %s", cudaGetErrorString(e));
    }
    cudaStreamSynchronize(p->s_compute);
    cudaStreamSynchronize(p->s_copy);
    cudaFree(d_in[0]); cudaFree(d_in[1]); cudaFree(d_out);
    return rc;
}
