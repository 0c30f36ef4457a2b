// kernels_delta.cuh -- SMPS random-element deltas built on the device.
//
// Reference: delta_coefficients, src/sd_algorithm/subprob.jl:104-121 (called from
// add_scenario!, epigraph.jl:92):  delta_rhs[row] = val - rbar[row] for RHS elements,
// delta_T[row, col] = val - Tbar[row, col] otherwise.
//
// HBM layout of the scenario store (per epigraph, per rank):
//   D[tile][j][128]   x-independent part  d = delta_rhs restricted to the stochastic rows
//                     S (j indexes S, padded with zero rows to s_pad), scenario i at
//                     (tile = i / 128, column i % 128).  This is exactly the operand tile
//                     the contraction streams, so it is written once and never reshaped.
//   dT[i][n_T]        row-major delta_T values (only when some element perturbs Tbar).
//   w[i]              scenario weights.
// One CUDA block per 128-aligned block of global scenario ordinals: realised values are
// read coalesced along the element axis, transposed through shared memory and written
// coalesced along the scenario axis.
#pragma once
#include "common.cuh"

namespace sqlp {

struct DeltaTables {
    int s;                    // random elements
    int n_T;                  // of which perturb Tbar
    const int *elem_j;        // [s] row slot in S of element e
    const int *elem_t;        // [s] slot in dT of element e, -1 for RHS elements
    const double *elem_base;  // [s] rbar[row] or Tbar[row, col] of element e
    // optional outcome tables for device-side sampling
    const double *out_vals;   // [s][mo]
    const double *out_cdf;    // [s][mo]
    const int *out_cnt;       // [s]
    int mo;
};

#define SQLP_DELTA_SLAB 32

// SAMPLE = false: values[i_batch][e] given.  SAMPLE = true: drawn from the outcome tables.
template <bool SAMPLE>
__global__ void __launch_bounds__(256)
k_delta_build(DeltaTables tb, const double *__restrict__ values, long long g0, long long n_new,
              int rank, int world, int s_pad, double *__restrict__ D, double *__restrict__ dT,
              double *__restrict__ w, const double *__restrict__ w_batch, unsigned long long seed,
              unsigned long long wseed)
{
    __shared__ double sh[SQLP_DELTA_SLAB][SQLP_TILE + 1];
    const long long gblock = g0 / SQLP_TILE + blockIdx.x;   // global 128-block
    if ((int)(gblock % world) != rank) return;
    const long long gb0 = gblock * SQLP_TILE;
    const long long lo = max(gb0, g0), hi = min(gb0 + SQLP_TILE, g0 + n_new);
    const int c0 = (int)(lo - gb0), c1 = (int)(hi - gb0);   // columns [c0, c1) of the tile
    const long long ltile = gblock / world;                  // local tile index
    double *Dt = D + ltile * (long long)s_pad * SQLP_TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;

    // weights
    for (int c = c0 + threadIdx.x; c < c1; c += blockDim.x) {
        long long g = gb0 + c;
        double wt = 1.0;
        if (SAMPLE) {
            if (wseed) wt = 0.5 + u01(wseed, (unsigned long long)g);
        } else if (w_batch) {
            wt = w_batch[g - g0];
        }
        w[ltile * SQLP_TILE + c] = wt;
    }

    for (int e0 = 0; e0 < tb.s; e0 += SQLP_DELTA_SLAB) {
        const int ne = min(SQLP_DELTA_SLAB, tb.s - e0);
        // phase 1: one warp per scenario, lanes along the element axis (coalesced reads)
        for (int c = c0 + warp; c < c1; c += nwarp) {
            if (lane < ne) {
                const int e = e0 + lane;
                const long long g = gb0 + c;
                double val;
                if (SAMPLE) {
                    double u = u01(seed, (unsigned long long)g * tb.s + e);
                    const double *cdf = tb.out_cdf + (long long)e * tb.mo;
                    int idx = 0;
                    for (int q = 0; q < tb.mo; ++q) idx += (u >= cdf[q]) ? 1 : 0;
                    idx = min(idx, max(tb.out_cnt[e] - 1, 0));
                    val = tb.out_vals[(long long)e * tb.mo + idx];
                } else {
                    val = values[(g - g0) * tb.s + e];
                }
                sh[lane][c] = __dsub_rn(val, tb.elem_base[e]);   // :114 / :117
            }
        }
        __syncthreads();
        // phase 2: one warp per element, lanes along the scenario axis (coalesced writes)
        for (int q = warp; q < ne; q += nwarp) {
            const int e = e0 + q;
            const int t = tb.elem_t[e];
            for (int c = c0 + lane; c < c1; c += 32) {
                if (t < 0)
                    Dt[(long long)tb.elem_j[e] * SQLP_TILE + c] = sh[q][c];
                else
                    dT[(ltile * SQLP_TILE + c) * (long long)tb.n_T + t] = sh[q][c];
            }
        }
        __syncthreads();
    }
}

// d(x) = delta_rhs - delta_T * x on the stochastic rows (subprob.jl:149), needed only when
// some random element perturbs Tbar.  T elements are pre-sorted by (row slot, column).
// Dx starts as a copy of D; one thread per scenario walks the sorted element list.
struct TransferList {
    int n_T;
    const int *t_j;      // [n_T] row slot in S, ascending
    const int *t_col;    // [n_T] first-stage column
    const int *t_slot;   // [n_T] slot in dT
};

__global__ void k_delta_x(TransferList tl, const double *__restrict__ x, long long n_local,
                          int s_pad, const double *__restrict__ D, const double *__restrict__ dT,
                          double *__restrict__ Dx)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_local) return;
    const long long tile = i >> 7;
    const int c = (int)(i & 127);
    const double *Dt = D + tile * (long long)s_pad * SQLP_TILE + c;
    double *Xt = Dx + tile * (long long)s_pad * SQLP_TILE + c;
    const double *row = dT + i * (long long)tl.n_T;
    int q = 0;
    while (q < tl.n_T) {
        const int j = tl.t_j[q];
        double acc = 0.0;
        while (q < tl.n_T && tl.t_j[q] == j) {
            acc = __dadd_rn(acc, __dmul_rn(row[tl.t_slot[q]], x[tl.t_col[q]]));
            ++q;
        }
        Xt[(long long)j * SQLP_TILE] = __dsub_rn(Dt[(long long)j * SQLP_TILE], acc);
    }
}

}  // namespace sqlp
