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
// i.e. 2 nlat K^2 + 2 P (2L)^2 flops (config 4: 83.6 GF instead of 45.9 TF).  This restructuring is
// DECLARED in DESIGN.md; bench_configs.py reports contract and executed flops separately, and the
// direct kernel without any structure assumption is gb_points_covariance (gb_points.cu).
//
// Both contractions run on the persistent DMMA GEMM of gb_gemm.cuh (operands staged by the TMA unit
// with cp.async.bulk, 3-stage mbarrier pipeline); the second factor and the reduction over rows are
// folded into the epilogues with warp shuffles.  Every partial sum has ONE writer and the partials are
// added in a fixed order by small reduction kernels: the result is bit-identical from run to run.
//
//   1. gb_cov_permute    Sigma (degree-wise order) -> tiled operand St[row tile a'][b][132]: order-wise
//                        order, groups padded to 8 rows (a') / 4 columns (b), zeros in the padding.
//                        For a symmetric Sigma only the blocks k <= k' are read and written; optionally
//                        slice by slice right in front of the GEMM that consumes the slice (see the host code)
//   2. gb_cov_legendre   on-the-fly Legendre recursion * kn for the requested parallels -> U as GEMM
//                        B tiles [group k' * nti + parallel tile][c][124]
//   3. GEMM + QuadEpilogue   C_k'[a', i] = sum_{b in k'} St[a'][b] U[b][i];  Hpart[i][k'][piece(a')] = sum U[a'][i] C
//      gb_cov_reduce_h       H[i][k][k'] = sum of the pieces of group k, in order
//   4. GEMM + LonEpilogue    W_i[k, j] = sum_k' H_i[k][k'] T[k'][j];          varpart[i][slot][j] = sum T[k][j] W
//   5. gb_cov_finish     var = sum of the slots in order, optional sqrt
#include <cstdlib>
#include <vector>
#include <cmath>
#include "gb_common.cuh"
#include "gb_gemm.cuh"

namespace {

using gb::legendre_column;

// St[gb_ab_offset(a', b, Kp4)] = Sigma[perm8[a']][perm4[b]]  (0 where either index is padding)
__global__ void __launch_bounds__(256)
gb_cov_permute(const double* __restrict__ sigma, double* __restrict__ St, const int* __restrict__ perm8,
               const int* __restrict__ perm4, int rows_a, int Kp4, long long K) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;   // a' fastest: coalesced writes
    const int b = blockIdx.y;
    if (a >= rows_a) return;
    const int pa = perm8[a], pb = perm4[b];
    St[gb_ab_offset(a, b, Kp4)] = (pa < 0 || pb < 0) ? 0.0 : sigma[(size_t)pa * K + pb];
}

// The same permutation, one CTA per (32 rows a', degree n of the columns): the 2n+1 columns of a degree
// are contiguous in the degree-wise order, so Sigma is read in row segments and St written in 256-byte
// runs (the element-wise gather above fetches a 32-byte sector per double).
// a_first: first row a' of the slice (St holds the slice's row tiles only).  Symmetric Sigma (rowgroup != nullptr): a
// row tile only meets column groups k' >= its own first group kmin, i.e. the columns j >= jmin of every degree: the
// rest of the row block is neither read nor written (the GEMM skips those column tiles, gbgemm::Shape::mt_first_nt).
constexpr int CP_ROWS = 32;
__global__ void __launch_bounds__(256)
gb_cov_permute_degree(const double* __restrict__ sigma, double* __restrict__ St, const int* __restrict__ perm8,
                      const int* __restrict__ goff4, int Kp4, long long K, int nmin, int a_first,
                      const int* __restrict__ first_group) {
    extern __shared__ double s_p[];   // [CP_ROWS][2n+1]
    const int a0 = blockIdx.x * CP_ROWS;              // inside the slice
    const int n = nmin + blockIdx.y;
    const int width = 2 * n + 1;      // odd pitch: conflict-free column reads
    int jmin = 0;
    if (first_group) {
        const int kmin = first_group[(a_first + a0) >> 7];         // first group of the 128-row tile
        jmin = kmin <= 1 ? 0 : kmin - 1;                              // group 2m (cos) is column 2m-1, group 2m+1 (sin) column 2m
    }
    if (jmin >= width) return;
    const long long col0 = (long long)n * n - (long long)nmin * nmin;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int r = warp; r < CP_ROWS; r += 8) {
        const int pa = perm8[a_first + a0 + r];
        const double* src = sigma + (size_t)(pa < 0 ? 0 : pa) * K + col0;
        // Sigma is read once: streaming loads keep the re-tiled slice (and U) in L2 instead
        for (int j = jmin + lane; j < width; j += 32) s_p[r * width + j] = pa < 0 ? 0.0 : __ldcs(src + j);
    }
    __syncthreads();
    for (int j = jmin + warp; j < width; j += 8) {
        const int m = (j + 1) >> 1;
        const int k = (j == 0) ? 0 : 2 * m + ((j & 1) ? 0 : 1);          // j = 2m-1: cos, j = 2m: sin
        const int b = goff4[k] + n - max(m, nmin);
        double* dst = St + ((size_t)(a0 >> 7) * Kp4 + b) * GB_LDA + (a0 & (GB_TM - 1));
#pragma unroll
        for (int r = lane; r < CP_ROWS; r += 32) dst[r] = s_p[r * width + j];
    }
}

// rows b of St that pad a group to a multiple of four
__global__ void __launch_bounds__(128)
gb_cov_zero_rows(double* __restrict__ St, const int* __restrict__ padrows, int Kp4) {
    St[((size_t)blockIdx.x * Kp4 + padrows[blockIdx.y]) * GB_LDA + threadIdx.x] = 0.0;
}

// U as B tiles: Ut[((k * nti + i/120) * Kg + (n - n0)) * 124 + i % 120] = kn[i][n] * P_nm(theta_i)
__global__ void __launch_bounds__(128)
gb_cov_legendre(double* __restrict__ Ut, const double* __restrict__ ct, const double* __restrict__ kn,
                const double* __restrict__ pmm, const double* __restrict__ ra, const double* __restrict__ rb,
                const double* __restrict__ rc, int L, int nmin, int row0, int nrows, int nti, int Kg,
                const double* __restrict__ wn) {
    const int il = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y;
    if (il >= nrows) return;
    const int i = row0 + il;
    const int n0 = max(m, nmin);
    const double* kn_i = kn + (size_t)i * L;
    const int it = il / GB_S2_TN, ic = il % GB_S2_TN;
    double* uc = Ut + ((size_t)(2 * m) * nti + it) * Kg * GB_S2_LDB + ic;
    double* us = Ut + ((size_t)(2 * m + 1) * nti + it) * Kg * GB_S2_LDB + ic;
    legendre_column(m, L, ct[i], pmm[(size_t)i * L + m], ra, rb, rc, [&](int n, double p) {
        if (n < n0) return;
        double v = __dmul_rn(p, kn_i[n]);
        if (wn) v *= wn[n];                       // isotropic filter: F = diag(w_n)
        uc[(size_t)(n - n0) * GB_S2_LDB] = v;
        if (m > 0) us[(size_t)(n - n0) * GB_S2_LDB] = v;
    });
}

// Filtered covariances: diag(A F Sigma F' A') = diag((A F) Sigma (A F)').  An order-wise filter F only mixes
// coefficients of one (order, cos|sin) group, so A F keeps the product structure of A with
//   U'[b][i] = sum_{a in group} W_g[a][b] U[a][i]          (W_g: the group's block of F, reference filter.py:193-222)
// One thread per parallel, eight output degrees per CTA; the U tile of the group is re-read from L2.
constexpr int CF_B = 8;
__global__ void __launch_bounds__(128)
gb_cov_apply_blocks(const double* __restrict__ Ut, double* __restrict__ Ut2, const double* __restrict__ blocks,
                    const long long* __restrict__ offsets, int nf, int L, int nmin, int nti, int Kg) {
    const int k = blockIdx.x;                 // group 2m + cs
    const int it = blockIdx.y;
    const int b0 = blockIdx.z * CF_B;
    const int m = k >> 1, cs = k & 1;
    if (m >= L || (m == 0 && cs == 1)) return;
    const int n0 = max(m, nmin);
    const int cnt = L - n0;
    if (b0 >= cnt) return;
    const int g = (m == 0) ? 0 : 2 * m - 1 + cs;          // block index of the filter (filter.py:180-187)
    const int kf = nf + 1 - m;
    const double* W = blocks + offsets[g] + (size_t)(n0 - m) * kf + (n0 - m);   // rows / columns of degree >= n0
    const size_t tile = ((size_t)k * nti + it) * Kg * GB_S2_LDB;
    const double* u = Ut + tile + threadIdx.x;
    double acc[CF_B];
#pragma unroll
    for (int j = 0; j < CF_B; ++j) acc[j] = 0.0;
    if (threadIdx.x < GB_S2_LDB) {
        for (int a = 0; a < cnt; ++a) {
            const double ua = u[(size_t)a * GB_S2_LDB];
            const double* wrow = W + (size_t)a * kf + b0;
#pragma unroll
            for (int j = 0; j < CF_B; ++j)
                if (b0 + j < cnt) acc[j] = fma(__ldg(wrow + j), ua, acc[j]);
        }
#pragma unroll
        for (int j = 0; j < CF_B; ++j)
            if (b0 + j < cnt) Ut2[tile + (size_t)(b0 + j) * GB_S2_LDB + threadIdx.x] = acc[j];
    }
}

// sum over the 8 rows of a fragment slab (lanes with equal lane % 4)
__device__ __forceinline__ double slab_sum(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 16);
    return v;
}

// Hpart[(i * kpad + k') * n_pieces + piece]: the four 8-row slabs a thread holds are multiplied by U and summed in the
// thread as long as they belong to the same group k; such a run of slabs inside a warp's 32 rows is a "piece" with exactly
// one writer (host table piece_of[(row tile * 4 + warp row) * 4 + first slab of the run]).  gb_cov_reduce_h adds the
// pieces of every group in order and writes H in the A-tile layout of stage 4.
struct QuadEpilogue {
    static constexpr bool whole_tile = true;
    const double* Ut;        // B tiles of U (also the source of the row-side factor)
    double* Hpart;
    const int* rowgroup;     // [rows_a / 8] group k of each 8-row slab, -1 for padding
    const int* goff8;        // [kpad] first a' of each group
    const int* piece_of;     // [rows_a / 8] piece of the run that starts at this slab (valid at run heads)
    int nti, Kg, kpad, n_pieces, nrows;
    int symmetric;           // only group pairs k <= k' are kept (the off-diagonal ones count twice in the reduction)
    int row_offset;          // first row a' of the slice the GEMM runs on
    struct Pre {
        int kslab[4];          // group of each of the thread's four 8-row slabs (uniform over the warp), -1: padding
        int uoff[4];           // offset of the slab's row inside the U tiles of its group (checked on the host)
        int piece[4];          // piece of the run starting at slab mi
        int kprime, i0;        // column group of the tile, first parallel of this thread (column pair i0, i0 + 1)
    };
    __device__ __forceinline__ Pre prepare(long long row_base, int nt, int col_base) const {
        Pre pr;
        const int it = nt % nti;
        pr.kprime = nt / nti;
        const int cc = col_base - nt * GB_S2_TN;              // column of the thread inside the tile (even)
        pr.i0 = it * GB_S2_TN + cc;
        const int ubase = it * Kg * GB_S2_LDB + cc;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            const long long row = row_base + row_offset + mi * 8;
            const int k = rowgroup[row >> 3];
            pr.kslab[mi] = k;
            pr.piece[mi] = piece_of[row >> 3];
            pr.uoff[mi] = k >= 0 ? ubase + (k * nti * Kg + (int)(row - goff8[k])) * GB_S2_LDB : 0;
        }
        // the row-side factors are needed after the K loop: start pulling them into L1 now
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
            if (pr.kslab[mi] >= 0) {
#pragma unroll
                for (int ni = 0; ni < 5; ++ni)
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(Ut + pr.uoff[mi] + ni * 8));
            }
        return pr;
    }
    __device__ __forceinline__ void tile(const Pre& pr, long long, int, double (&acc)[4][5][2]) const {
        // 1. multiply by the row-side factor in place: twenty independent loads in flight, no extra registers
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            if (pr.kslab[mi] < 0) continue;                    // warp-uniform
#pragma unroll
            for (int ni = 0; ni < 5; ++ni) {
                const double2 u = __ldg(reinterpret_cast<const double2*>(Ut + pr.uoff[mi] + ni * 8));   // 16-byte aligned
                acc[mi][ni][0] *= u.x;
                acc[mi][ni][1] *= u.y;
            }
        }
        // 2. every run of slabs with one group: sum in the thread, reduce across the slab, store the piece
#pragma unroll
        for (int head = 0; head < 4; ++head) {
            const int k = pr.kslab[head];
            if (k < 0 || (head > 0 && k == pr.kslab[head - 1])) continue;     // padding, or not the first slab of its run
            if (symmetric && k > pr.kprime) continue;
            double* h = Hpart + ((size_t)pr.i0 * kpad + pr.kprime) * n_pieces + pr.piece[head];
            const size_t hstep = (size_t)kpad * n_pieces;                     // to the next parallel
#pragma unroll
            for (int ni = 0; ni < 5; ++ni) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int mi = head; mi < 4; ++mi)
                    if (pr.kslab[mi] == k) {
                        s0 += acc[mi][ni][0];
                        s1 += acc[mi][ni][1];
                    }
                s0 = slab_sum(s0);
                s1 = slab_sum(s1);
                if ((threadIdx.x & 31) < 4) {
                    const int i = pr.i0 + ni * 8;
                    if (i < nrows) h[(size_t)(ni * 8) * hstep] = s0;
                    if (i + 1 < nrows) h[(size_t)(ni * 8 + 1) * hstep] = s1;
                }
            }
        }
    }
};

// Ht[((i * hmt + k/128) * kpad + k') * 132 + k%128] = scale * sum_{pieces of k} Hpart[(i * kpad + k') * n_pieces + piece]
// (the A-tile layout of stage 4; every element of Ht is written, padding rows with zero).  Symmetric: pairs k > k' are
// zero, pairs k < k' count twice.  CTA = (parallel i, column group k'), thread = row group k.
__global__ void __launch_bounds__(256)
gb_cov_reduce_h(const double* __restrict__ Hpart, double* __restrict__ Ht, const int* __restrict__ pstart, int kpad,
                int n_pieces, int hmt, int symmetric) {
    const int i = blockIdx.x, kp = blockIdx.y;
    const double* src = Hpart + ((size_t)i * kpad + kp) * n_pieces;
    for (int k = threadIdx.x; k < hmt * GB_LDA; k += blockDim.x) {
        const int kt = k / GB_LDA, kk = k - kt * GB_LDA;
        const int kg = kt * GB_TM + kk;
        double v = 0.0;
        if (kk < GB_TM && kg < kpad && !(symmetric && kg > kp)) {
            for (int pc = pstart[kg]; pc < pstart[kg + 1]; ++pc) v += src[pc];
            if (symmetric && kg < kp) v *= 2.0;
        }
        Ht[(((size_t)i * hmt + kt) * kpad + kp) * GB_LDA + kk] = v;
    }
}

// varpart[(i * 4 hmt + slot) * nlon + j] = sum_k T[k][j] * W_i[k][j] over the 32 spectral rows of one warp row of one
// row tile (slot = 4 * (row tile of the parallel) + warp row: one writer); rows of the GEMM = (i, k), all 128 rows of a
// tile belong to one parallel, so the thread sums its four rows before the slab reduction.  gb_cov_finish adds the slots.
struct LonEpilogue {
    static constexpr bool whole_tile = true;
    const double* trig;      // [kpad][nlp]
    double* var;             // [nrows][4 hmt][nlon]
    int hmt, kpad, nlp, nlon;
    __device__ __forceinline__ int prepare(long long, int, int) const { return 0; }
    __device__ __forceinline__ void tile(int, long long row_base, int col_base, double (&acc)[4][5][2]) const {
        const int mt = (int)(row_base >> 7);
        const int i = mt / hmt;
        const int kb = (mt % hmt) * 128 + (int)(row_base & 127);
#pragma unroll
        for (int ni = 0; ni < 5; ++ni) {
            const int col = col_base + ni * 8;
            double p0 = 0.0, p1 = 0.0;
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                const int k = kb + mi * 8;
                if (k < kpad) {
                    const double* t = trig + (size_t)k * nlp + col;
                    if (col < nlp) p0 = fma(acc[mi][ni][0], t[0], p0);
                    if (col + 1 < nlp) p1 = fma(acc[mi][ni][1], t[1], p1);
                }
            }
            p0 = slab_sum(p0);
            p1 = slab_sum(p1);
            if ((threadIdx.x & 31) < 4) {
                const int slot = (mt % hmt) * 4 + (int)((row_base & 127) >> 5);
                double* o = var + ((size_t)i * (4 * hmt) + slot) * nlon + col;
                if (col < nlon) o[0] = p0;
                if (col + 1 < nlon) o[1] = p1;
            }
        }
    }
};

__global__ void gb_cov_finish(const double* __restrict__ part, double* __restrict__ out, int nslots, int nlon, long long n,
                              int take_sqrt) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const long long i = idx / nlon;
    const int j = (int)(idx - i * nlon);
    double v = 0.0;
    for (int s = 0; s < nslots; ++s) v += part[((size_t)i * nslots + s) * nlon + j];
    out[idx] = take_sqrt ? sqrt(v) : v;
}

// Index tables of one (min_degree, parallel count): order-wise layouts, permutations, tile offsets.  Built on the host
// and uploaded once, kept with the plan: a repeated propagation launches its kernels without host work or copies.
struct CovLayout {
    int nmin = -1, nrows = -1;
    int Kp8 = 0, Kp4 = 0, Kg = 8, n_atiles = 0, rows_a = 0, nti = 0, n_ct = 0, hmt = 0, n_padrows = 0, n_pieces = 0;
    int* d_all = nullptr;     // one allocation; the pointers below point into it
    int *d_perm8 = nullptr, *d_perm4 = nullptr, *d_rowgroup = nullptr, *d_goff8 = nullptr, *d_koff = nullptr,
        *d_klen = nullptr, *d_goff4 = nullptr, *d_padrows = nullptr, *d_first_nt = nullptr, *d_piece_of = nullptr,
        *d_pstart = nullptr, *d_first_group = nullptr;
};

int build_layout(gb_plan* p, int nmin, int nrows, CovLayout** out) {
    CovLayout* c = static_cast<CovLayout*>(p->cov_layout);
    if (c && c->nmin == nmin && c->nrows == nrows) {
        *out = c;
        return GB_OK;
    }
    gb_cov_layout_free(p);
    c = new CovLayout();
    const int L = p->L, kpad = p->kpad;
    // order-wise layouts: group k = 2m + cs holds degrees n0(m)..nmax
    std::vector<int> gcnt(kpad, 0), goff8(kpad, 0), goff4(kpad, 0);
    int Kp8 = 0, Kp4 = 0, Kg = 8;   // Kg: rows per U tile (covers the 8-padded groups read by the epilogue)
    for (int k = 0; k < kpad; ++k) {
        const int m = k >> 1, cs = k & 1;
        int cnt = 0;
        if (m < L && !(m == 0 && cs == 1)) cnt = L - (m > nmin ? m : nmin);
        gcnt[k] = cnt;
        goff8[k] = Kp8;
        goff4[k] = Kp4;
        Kp8 += (cnt + 7) / 8 * 8;
        Kp4 += (cnt + 3) / 4 * 4;
        if ((cnt + 7) / 8 * 8 > Kg) Kg = (cnt + 7) / 8 * 8;
    }
    const int n_atiles = (Kp8 + GB_TM - 1) / GB_TM;
    const int rows_a = n_atiles * GB_TM;
    std::vector<int> perm8(rows_a, -1), perm4(Kp4, -1), rowgroup(rows_a / 8, -1);
    for (int k = 0; k < kpad; ++k) {
        const int m = k >> 1, cs = k & 1;
        const int n0 = (m > nmin ? m : nmin);
        for (int r = 0; r < gcnt[k]; ++r) {
            const int n = n0 + r;
            const long long idx = (long long)n * n + (m == 0 ? 0 : 2 * m - 1 + cs) - (long long)nmin * nmin;
            perm8[goff8[k] + r] = (int)idx;
            perm4[goff4[k] + r] = (int)idx;
        }
        for (int r = 0; r < (gcnt[k] + 7) / 8; ++r) rowgroup[goff8[k] / 8 + r] = k;
    }
    const int nti = (nrows + GB_S2_TN - 1) / GB_S2_TN;          // parallel tiles
    const int n_ct = kpad * nti;                                // column tiles of the quadratic-form GEMM
    std::vector<int> nt_koff(n_ct), nt_klen(n_ct);
    for (int k = 0; k < kpad; ++k)
        for (int it = 0; it < nti; ++it) {
            nt_koff[k * nti + it] = goff4[k];
            nt_klen[k * nti + it] = (gcnt[k] + 3) / 4 * 4;
        }
    std::vector<int> padrows;
    for (int b = 0; b < Kp4; ++b)
        if (perm4[b] < 0) padrows.push_back(b);
    // symmetric Sigma: H_i is symmetric, row tile mt needs the column groups k' >= its first group only
    std::vector<int> first_nt(n_atiles, 0);
    for (int t = 0; t < n_atiles; ++t) {
        int kmin = kpad;
        for (int sl = t * (GB_TM / 8); sl < (t + 1) * (GB_TM / 8); ++sl)
            if (rowgroup[sl] >= 0 && rowgroup[sl] < kmin) kmin = rowgroup[sl];
        first_nt[t] = kmin * nti;
    }
    // pieces: runs of slabs of one group inside a warp's 32 rows (4 slabs); the pieces of a group are consecutive
    std::vector<int> piece_of(rows_a / 8, -1), pstart(kpad + 1, 0);
    int n_pieces = 0;
    {
        std::vector<int> count(kpad, 0);
        for (int w = 0; w < rows_a / 32; ++w)
            for (int sl = 0; sl < 4; ++sl) {
                const int k = rowgroup[w * 4 + sl];
                if (k < 0 || (sl > 0 && rowgroup[w * 4 + sl - 1] == k)) continue;
                piece_of[w * 4 + sl] = n_pieces++;
                ++count[k];
            }
        for (int k = 0; k < kpad; ++k) pstart[k + 1] = pstart[k] + count[k];   // groups appear in increasing order of a'
    }
    c->n_pieces = n_pieces;
    std::vector<int> first_group(n_atiles, kpad);
    for (int t = 0; t < n_atiles; ++t) first_group[t] = first_nt[t] / nti;
    c->nmin = nmin; c->nrows = nrows;
    c->Kp8 = Kp8; c->Kp4 = Kp4; c->Kg = Kg; c->n_atiles = n_atiles; c->rows_a = rows_a; c->nti = nti; c->n_ct = n_ct;
    c->hmt = (kpad + GB_TM - 1) / GB_TM;
    c->n_padrows = (int)padrows.size();
    constexpr int NT = 12;
    const std::vector<int>* parts[NT] = {&perm8, &perm4, &rowgroup, &goff8, &nt_koff, &nt_klen, &goff4, &padrows, &first_nt,
                                         &piece_of, &pstart, &first_group};
    int** slots[NT] = {&c->d_perm8, &c->d_perm4, &c->d_rowgroup, &c->d_goff8, &c->d_koff, &c->d_klen, &c->d_goff4,
                       &c->d_padrows, &c->d_first_nt, &c->d_piece_of, &c->d_pstart, &c->d_first_group};
    std::vector<int> all;
    size_t offs[NT];
    for (int i = 0; i < NT; ++i) {
        offs[i] = all.size();
        all.insert(all.end(), parts[i]->begin(), parts[i]->end());
        all.resize((all.size() + 3) / 4 * 4);                   // keep every table 16-byte aligned
    }
    if (cudaMalloc(reinterpret_cast<void**>(&c->d_all), (all.size() ? all.size() : 4) * sizeof(int)) != cudaSuccess) {
        delete c;
        return gb_set_error(GB_ERR_CUDA, "gb_covariance_propagation: cannot allocate the index tables");
    }
    if (cudaMemcpy(c->d_all, all.data(), all.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(c->d_all);
        delete c;
        return gb_set_error(GB_ERR_CUDA, "gb_covariance_propagation: cannot upload the index tables");
    }
    for (int i = 0; i < NT; ++i) *slots[i] = c->d_all + offs[i];
    p->cov_layout = c;
    *out = c;
    return GB_OK;
}

}  // namespace

void gb_cov_layout_free(gb_plan* p) {
    CovLayout* c = static_cast<CovLayout*>(p->cov_layout);
    if (!c) return;
    cudaFree(c->d_all);
    delete c;
    p->cov_layout = nullptr;
}

extern "C" int gb_covariance_propagation(gb_plan* plan, const double* d_sigma, int nmin, int row0, int nrows,
                                         double* d_out, int flags, void* stream) {
    return gb_covariance_propagation_filtered(plan, d_sigma, nmin, row0, nrows, d_out, flags, nullptr, nullptr, 0, nullptr,
                                              stream);
}

extern "C" int gb_covariance_propagation_filtered(gb_plan* plan, const double* d_sigma, int nmin, int row0, int nrows,
                                                  double* d_out, int flags, const double* d_blocks,
                                                  const int64_t* block_offsets, int nf, const double* d_wn, void* stream) {
    const int take_sqrt = flags & GB_COV_SQRT;
    const int symmetric = (flags & GB_COV_SYMMETRIC) ? 1 : 0;
    GB_REQUIRE(plan != nullptr, "gb_covariance_propagation: plan is NULL");
    gb_plan* p = plan;
    GB_REQUIRE(nmin >= 0 && nmin <= p->nmax, "gb_covariance_propagation: min_degree=%d outside [0, %d]", nmin, p->nmax);
    GB_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= p->nlat,
               "gb_covariance_propagation: parallels [%d, %d) outside the grid (%d parallels)", row0, row0 + nrows, p->nlat);
    if (nrows == 0) return GB_OK;
    GB_REQUIRE(d_sigma && d_out, "gb_covariance_propagation: NULL device pointer");
    GB_REQUIRE(!d_blocks || (block_offsets && nf >= p->nmax),
               "gb_covariance_propagation_filtered: the order-wise filter (degree %d) does not reach degree %d", nf, p->nmax);
    GB_CUDA(cudaSetDevice(p->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        int rca = gb_plan_acquire(p, st);   // the cached index tables are rebuilt when the shape changes
        if (rca) return rca;
    }
    gb_scratch scratch(st);           // frees every temporary of this call on all return paths
    const int L = p->L, kpad = p->kpad;
    const long long K = (long long)L * L - (long long)nmin * nmin;

    CovLayout* lay = nullptr;
    {
        int rc0 = build_layout(p, nmin, nrows, &lay);
        if (rc0) return rc0;
    }
    const int Kp4 = lay->Kp4, Kg = lay->Kg, n_atiles = lay->n_atiles, rows_a = lay->rows_a, nti = lay->nti,
              n_ct = lay->n_ct, hmt = lay->hmt;
    GB_REQUIRE(Kp4 <= 65535, "gb_covariance_propagation: degree %d is too large for this path", p->nmax);
    GB_REQUIRE((long long)n_ct * Kg * GB_S2_LDB < (1LL << 31),
               "gb_covariance_propagation: %d parallels at degree %d exceed the 32-bit tile index; pass row blocks", nrows, p->nmax);
    const size_t permute_smem = (size_t)CP_ROWS * (2 * p->nmax + 1) * sizeof(double);
    const bool by_degree = permute_smem <= 200 * 1024;
    int *d_perm8 = lay->d_perm8, *d_perm4 = lay->d_perm4, *d_rowgroup = lay->d_rowgroup, *d_goff8 = lay->d_goff8,
        *d_koff = lay->d_koff, *d_klen = lay->d_klen, *d_goff4 = lay->d_goff4, *d_padrows = lay->d_padrows,
        *d_first_nt = lay->d_first_nt;
    const int n_pieces = lay->n_pieces, nslots = 4 * hmt;
    // GB_COV_SLICE_MB=x re-tiles Sigma slice by slice (row tiles [t0, t1) of the order-wise operand, x MB each) into one
    // reused buffer that the GEMM of the slice reads straight away, so that the re-tiled copy lives in L2.  Measured on
    // config 4: 48 MB slices 4.0 ms, 96 MB 3.8 ms, one slice 3.76 ms -- the per-launch tail of 20 short GEMMs costs more
    // than the HBM round trip it saves, so the default is one slice.
    const size_t tile_bytes = (size_t)Kp4 * GB_LDA * sizeof(double);
    const char* env_mb = getenv("GB_COV_SLICE_MB");
    const double slice_mb = env_mb ? atof(env_mb) : 0.0;
    int slice_tiles = slice_mb > 0 ? (int)(slice_mb * 1048576.0 / (double)tile_bytes) : n_atiles;
    if (slice_tiles < 1) slice_tiles = 1;
    if (slice_tiles > n_atiles) slice_tiles = n_atiles;
    double *d_st = nullptr, *d_ut = nullptr, *d_ht = nullptr, *d_hpart = nullptr, *d_vpart = nullptr;
    int rc = GB_OK;
    const size_t st_elems = (size_t)slice_tiles * Kp4 * GB_LDA;
    const size_t ut_elems = (size_t)n_ct * Kg * GB_S2_LDB;
    const size_t ht_elems = (size_t)nrows * hmt * kpad * GB_LDA;
    const size_t hp_elems = (size_t)nrows * kpad * n_pieces;
    const size_t vp_elems = (size_t)nrows * nslots * p->nlon;
    GB_CUDA(scratch.alloc(&d_st, st_elems));
    GB_CUDA(scratch.alloc(&d_ut, ut_elems));
    GB_CUDA(scratch.alloc(&d_ht, ht_elems));
    GB_CUDA(scratch.alloc(&d_hpart, hp_elems));
    GB_CUDA(scratch.alloc(&d_vpart, vp_elems));
    GB_CUDA(cudaMemsetAsync(d_st, 0, st_elems * sizeof(double), st));      // padding rows of the groups stay zero
    GB_CUDA(cudaMemsetAsync(d_ut, 0, ut_elems * sizeof(double), st));
    if (permute_smem > 48 * 1024 && by_degree)
        GB_CUDA(cudaFuncSetAttribute(gb_cov_permute_degree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)permute_smem));
    {
        dim3 grid((nrows + 127) / 128, L);
        gb_cov_legendre<<<grid, 128, 0, st>>>(d_ut, p->d_ct, p->d_kn, p->d_pmm, p->d_ra, p->d_rb, p->d_rc, L, nmin,
                                              row0, nrows, nti, Kg, d_wn);
        GB_LAUNCH_CHECK();
    }
    long long* d_boff = nullptr;
    if (d_blocks) {
        // U <- F' U group by group (see gb_cov_apply_blocks); the GEMMs below then run on A F
        const int nblocks = 2 * nf + 1;
        GB_CUDA(scratch.alloc(&d_boff, (size_t)nblocks + 1));
        GB_CUDA(cudaMemcpyAsync(d_boff, block_offsets, (nblocks + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
        double* d_ut2 = nullptr;
        GB_CUDA(scratch.alloc(&d_ut2, ut_elems));
        GB_CUDA(cudaMemsetAsync(d_ut2, 0, ut_elems * sizeof(double), st));
        dim3 grid(kpad, nti, (Kg + CF_B - 1) / CF_B);
        gb_cov_apply_blocks<<<grid, 128, 0, st>>>(d_ut, d_ut2, d_blocks, d_boff, nf, L, nmin, nti, Kg);
        GB_LAUNCH_CHECK();
        d_ut = d_ut2;
    }
    for (int t0 = 0; t0 < n_atiles; t0 += slice_tiles) {
        const int t1 = (t0 + slice_tiles < n_atiles) ? t0 + slice_tiles : n_atiles;
        if (by_degree) {
            dim3 grid((t1 - t0) * GB_TM / CP_ROWS, L - nmin);
            gb_cov_permute_degree<<<grid, 256, permute_smem, st>>>(d_sigma, d_st, d_perm8, d_goff4, Kp4, K, nmin,
                                                                   t0 * GB_TM, symmetric ? lay->d_first_group : nullptr);
            GB_LAUNCH_CHECK();
        } else {
            dim3 grid(((t1 - t0) * GB_TM + 255) / 256, Kp4);
            gb_cov_permute<<<grid, 256, 0, st>>>(d_sigma, d_st, d_perm8 + (size_t)t0 * GB_TM, d_perm4, (t1 - t0) * GB_TM, Kp4, K);
            GB_LAUNCH_CHECK();
        }
        gbgemm::Shape sh;
        sh.A_t = d_st;
        sh.a_rows = Kp4;
        sh.a_koff_mul = 0;
        sh.tiles_per_group = 1;
        sh.B_t = d_ut;
        sh.b_rows = Kg;
        sh.klen = 0;
        sh.n_mtiles = t1 - t0;
        sh.n_ntiles = n_ct;
        sh.nt_koff = d_koff;
        sh.nt_klen = d_klen;
        sh.mt_first_nt = symmetric ? d_first_nt + t0 : nullptr;
        QuadEpilogue epi{d_ut, d_hpart, d_rowgroup, d_goff8, lay->d_piece_of, nti, Kg, kpad, n_pieces, nrows, symmetric,
                         t0 * GB_TM};
        if ((rc = gbgemm::launch(sh, epi, p->sm_count, st))) return rc;
    }
    {
        dim3 grid(nrows, kpad);
        gb_cov_reduce_h<<<grid, 256, 0, st>>>(d_hpart, d_ht, lay->d_pstart, kpad, n_pieces, hmt, symmetric);
        GB_LAUNCH_CHECK();
    }
    {
        gbgemm::Shape sh;
        sh.A_t = d_ht;
        sh.a_rows = kpad;
        sh.a_koff_mul = 0;
        sh.tiles_per_group = 1;
        sh.B_t = p->d_trig_t;
        sh.b_rows = kpad;
        sh.klen = kpad;
        sh.n_mtiles = nrows * hmt;
        sh.n_ntiles = p->n_ntiles;
        if (symmetric) {            // H is upper triangular then: spectral rows k >= 128 j only meet columns k' >= 128 j
            sh.mt_kstart_mod = hmt;
            sh.mt_kstart_mul = GB_TM;
        }
        LonEpilogue epi{p->d_trig, d_vpart, hmt, kpad, p->nlp, p->nlon};
        if ((rc = gbgemm::launch(sh, epi, p->sm_count, st))) return rc;
    }
    {
        const long long n = (long long)nrows * p->nlon;
        gb_cov_finish<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_vpart, d_out, nslots, p->nlon, n, take_sqrt ? 1 : 0);
        GB_LAUNCH_CHECK();
    }
    return GB_OK;
}
