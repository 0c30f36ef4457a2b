// kernels_delta.cuh -- SMPS random-element deltas built on the device.
//
// Reference: delta_coefficients, src/sd_algorithm/subprob.jl:104-121 (called from
// add_scenario!, epigraph.jl:92):  delta_rhs[row] = val - rbar[row] for RHS elements,
// delta_T[row, col] = val - Tbar[row, col] otherwise.
//
// HBM layout of the scenario store (per epigraph, per rank):
//   D[tile]   x-independent part  d = delta_rhs restricted to the stochastic rows S (slot j
//             indexes S, padded with zero slots to s_pad); scenario i is column i % 128 of
//             tile i / 128, stored in the fragment-major order of common.cuh tile_off.
//             This is exactly the operand tile the contraction streams, so it is written
//             once and never reshaped.
//   dT[i][n_T]  row-major delta_T values (only when some element perturbs Tbar).
//   w[i]        scenario weights.
// One CUDA block per 128-aligned block of global scenario ordinals: realised values are
// read coalesced along the element axis, staged in shared memory, and written out as the
// contiguous 4 KB (4 slots x 128 columns) cells of the tile layout.
#pragma once
#include "common.cuh"

namespace sqlp {

struct DeltaTables {
    int s;                    // random elements
    int n_T;                  // of which perturb Tbar
    int n_rows;               // stochastic row slots
    const int *slot_elem;     // [n_rows] RHS element of slot j, -1 if only T elements touch it
    const int *t_elem;        // [n_T] element of dT slot t
    const double *elem_base;  // [s] rbar[row] or Tbar[row, col] of element e
    // optional outcome tables for device-side sampling
    const double *out_vals;   // [s][mo]
    const double *out_cdf;    // [s][mo]
    const int *out_cnt;       // [s]
    int mo;
};

#define SQLP_DELTA_SLAB 32   // slots staged per pass (8 k-groups)

template <bool SAMPLE>
__device__ __forceinline__ double realised_value(const DeltaTables &tb, const double *values,
                                                 long long g, long long g0, int e,
                                                 unsigned long long seed)
{
    if (SAMPLE) {
        double u = u01(seed, (unsigned long long)g * tb.s + e);
        const double *cdf = tb.out_cdf + (long long)e * tb.mo;
        int idx = 0;
        for (int q = 0; q < tb.mo; ++q) idx += (u >= cdf[q]) ? 1 : 0;
        idx = min(idx, max(tb.out_cnt[e] - 1, 0));
        return tb.out_vals[(long long)e * tb.mo + idx];
    }
    return values[(g - g0) * tb.s + e];
}

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

    // delta_T values, row-major per scenario (rare path)
    for (int q = threadIdx.x; q < (c1 - c0) * tb.n_T; q += blockDim.x) {
        const int c = c0 + q / tb.n_T, t = q % tb.n_T, e = tb.t_elem[t];
        const double val = realised_value<SAMPLE>(tb, values, gb0 + c, g0, e, seed);
        dT[(ltile * SQLP_TILE + c) * (long long)tb.n_T + t] = __dsub_rn(val, tb.elem_base[e]);   // :117
    }

    for (int j0 = 0; j0 < tb.n_rows; j0 += SQLP_DELTA_SLAB) {
        // phase 1: one warp per scenario, lanes along the slot axis (coalesced reads)
        for (int c = c0 + warp; c < c1; c += nwarp) {
            const int j = j0 + lane;
            const int e = (j < tb.n_rows) ? tb.slot_elem[j] : -1;
            double d = 0.0;
            if (e >= 0)
                d = __dsub_rn(realised_value<SAMPLE>(tb, values, gb0 + c, g0, e, seed),
                              tb.elem_base[e]);                                                   // :114
            sh[lane][c] = d;
        }
        __syncthreads();
        // phase 2: each k-group (4 slots x 128 columns) is 512 contiguous doubles of the tile
        const int ngroups = min(SQLP_DELTA_SLAB / 4, (s_pad - j0) / 4);
        for (int q = threadIdx.x; q < ngroups * 512; q += blockDim.x) {
            const int gq = q >> 9, o = q & 511;
            const int P = o >> 6, t = (o & 63) >> 1, h = o & 1;
            const int c = (2 * P + h) * 8 + (t >> 2);
            if (c >= c0 && c < c1)
                Dt[(long long)(j0 / 4 + gq) * 512 + o] = sh[gq * 4 + (t & 3)][c];
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
    const long long tbase = (i >> 7) * (long long)s_pad * SQLP_TILE;
    const int c = (int)(i & 127);
    const double *row = dT + i * (long long)tl.n_T;
    int q = 0;
    while (q < tl.n_T) {
        const int j = tl.t_j[q];
        double acc = 0.0;
        while (q < tl.n_T && tl.t_j[q] == j) {
            acc = __dadd_rn(acc, __dmul_rn(row[tl.t_slot[q]], x[tl.t_col[q]]));
            ++q;
        }
        const long long o = tbase + tile_off(c, j);
        Dx[o] = __dsub_rn(D[o], acc);
    }
}

// One scenario's column of D gathered into a dense [n_rows] vector (delta readback).
__global__ void k_gather_column(const double *__restrict__ Dtile, int c, int n_rows,
                                double *__restrict__ out)
{
    for (int j = threadIdx.x; j < n_rows; j += blockDim.x) out[j] = Dtile[tile_off(c, j)];
}

}  // namespace sqlp
