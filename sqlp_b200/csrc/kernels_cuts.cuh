// kernels_cuts.cuh -- the epigraph's cut list on the device:
// evaluation of the piecewise approximation, the incumbent test and the master's cut rows.
//
// Reference:
//   evaluate_epigraph      src/sd_algorithm/epigraph.jl:177-220
//       best = lb;  for each cut: discount = weight_mark / total_weight,
//       val = discount * (alpha + dot(beta, x)) + (1 - discount) * lb, kept if val > best;
//       the incumbent cut enters undiscounted; the result is multiplied by the epigraph's weight.
//   add_cut_to_master!     src/sd_algorithm/epigraph.jl:101-117 (called by sync_cuts!, cell.jl:163-202)
//       row = (discount * alpha + (1 - discount) * lb,  discount * beta); incumbent: discount = 1.
//   check_improvement      src/sd_algorithm/improvement.jl:19-49
//
// A cut is a row (alpha, beta[n1], weight_mark).  The work is O(cuts * n1) per call -- tens of
// kilobytes -- so these kernels are latency bound by construction; they exist so that the cuts the
// reduction just produced never have to leave the device before the incumbent test and so that
// the master receives ONE dense block per epigraph.  dot(beta, x) is a fixed shuffle tree over
// lane partials (the reference's BLAS ddot has no defined order either); max is order independent.
#pragma once
#include "common.cuh"

namespace sqlp {

struct CutList {
    const double *cuts;   // [n][n1 + 2]
    int n;
    const double *inc;    // [n1 + 2] or null
};

__device__ __forceinline__ double warp_dot(const double *__restrict__ a, const double *__restrict__ b, int n, int lane)
{
    double s = 0.0;
    for (int j = lane; j < n; j += 32) s = fma(a[j], b[j], s);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    return s;
}

// out[l * NX + x] = weight * max(lb, discounted cuts at x, incumbent cut at x) for lists l = 0, 1
// (current approximation and the snapshot f_{k-1}); one block, warps over cuts.
__global__ void __launch_bounds__(256) k_cuts_evaluate(CutList cur, CutList last, int nlists,
                                                       const double *__restrict__ x, int NX, int n1,
                                                       double total_weight, double lb, double weight,
                                                       double *__restrict__ out)
{
    griddep_sync();
    __shared__ double red[8][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int RS = n1 + 2;
    for (int l = 0; l < nlists; ++l) {
        const CutList L = l ? last : cur;
        double best[2] = {lb, lb};
        for (int j = warp; j < L.n + (L.inc ? 1 : 0); j += 8) {
            const bool is_inc = (j == L.n);
            const double *row = is_inc ? L.inc : L.cuts + (long long)j * RS;
            const double discount = row[n1 + 1] / total_weight;
            for (int p = 0; p < NX; ++p) {
                const double d = warp_dot(row + 1, x + (long long)p * n1, n1, lane);
                const double v0 = __dadd_rn(row[0], d);   // no contraction: the reference does not fuse
                const double val = is_inc ? v0                                            // epigraph.jl:195
                                          : __dadd_rn(__dmul_rn(discount, v0),
                                                      __dmul_rn(__dsub_rn(1.0, discount), lb));   // :184-185
                if (val > best[p]) best[p] = val;
            }
        }
        if (lane == 0) { red[warp][l * 2] = best[0]; red[warp][l * 2 + 1] = best[1]; }
    }
    __syncthreads();
    if (threadIdx.x < nlists * NX) {
        const int l = threadIdx.x / NX, p = threadIdx.x % NX;
        double b = lb;
        for (int w = 0; w < 8; ++w) if (red[w][l * 2 + p] > b) b = red[w][l * 2 + p];
        out[l * NX + p] = __dmul_rn(weight, b);   // epigraph.jl:210
    }
}

// The master rows of one epigraph: n discounted cuts, then the undiscounted incumbent cut.
__global__ void k_cuts_master_rows(CutList L, int n1, double total_weight, double lb, double *__restrict__ rows)
{
    griddep_sync();
    const int RS = n1 + 2, RW = n1 + 1;
    const int total = (L.n + (L.inc ? 1 : 0)) * RW;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total; q += gridDim.x * blockDim.x) {
        const int j = q / RW, c = q % RW;
        const bool is_inc = (j == L.n);
        const double *row = is_inc ? L.inc : L.cuts + (long long)j * RS;
        const double discount = is_inc ? 1.0 : row[n1 + 1] / total_weight;
        rows[q] = (c == 0) ? __dadd_rn(__dmul_rn(discount, row[0]), __dmul_rn(__dsub_rn(1.0, discount), lb))
                           : __dmul_rn(discount, row[c]);
    }
}

// (alpha, beta, anything) -> (alpha, beta, weight_mark): a reduction result becomes a stored cut.
__global__ void k_cut_store(const double *__restrict__ src, double weight_mark, int n1, double *__restrict__ dst)
{
    griddep_sync();
    for (int q = threadIdx.x; q < n1 + 2; q += blockDim.x) dst[q] = (q == n1 + 1) ? weight_mark : src[q];
}

// deleteat!: gather the kept rows (ascending) into tmp.
__global__ void k_cuts_gather(const double *__restrict__ cuts, const int *__restrict__ keep, int n_keep, int RS,
                              double *__restrict__ tmp)
{
    griddep_sync();
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n_keep * RS; q += gridDim.x * blockDim.x)
        tmp[q] = cuts[(long long)keep[q / RS] * RS + q % RS];
}

// check_improvement: est[e] = {cur@cand, cur@inc, last@cand, last@inc} per epigraph, cost . x added
// once; sums over epigraphs in order (improvement.jl:27-36).  out = {candidate_estimation,
// incumbent_estimation, required_improvement, is_improved}.
__global__ void k_improvement(const double *__restrict__ est, int n_epi, const double *__restrict__ cost,
                              const double *__restrict__ x2, int n1, double q_factor, double *__restrict__ out)
{
    griddep_sync();
    const int lane = threadIdx.x & 31;
    if (threadIdx.x >= 32) return;
    const double f_cand = warp_dot(cost, x2, n1, lane), f_inc = warp_dot(cost, x2 + n1, n1, lane);
    if (lane) return;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (int e = 0; e < n_epi; ++e)
        for (int k = 0; k < 4; ++k) s[k] += est[e * 4 + k];
    const double cand = __dadd_rn(s[0], f_cand), inc = __dadd_rn(s[1], f_inc);
    const double required = __dmul_rn(q_factor, __dsub_rn(__dadd_rn(s[2], f_cand), __dadd_rn(s[3], f_inc)));
    out[0] = cand;
    out[1] = inc;
    out[2] = required;
    out[3] = (cand < __dadd_rn(inc, required)) ? 1.0 : 0.0;
}

}  // namespace sqlp
