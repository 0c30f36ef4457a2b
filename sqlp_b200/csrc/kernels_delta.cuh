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
// One CUDA block per 64-scenario half of a 128-aligned block of global scenario ordinals:
// realised values are read coalesced along the element axis, staged in shared memory, and
// written out as the contiguous 2 KB (4 slots x 64 columns) half cells of the tile layout.
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
    // optional continuous elements: kind 0 DISCRETE (tables above), 1 NORMAL(mean, variance),
    // 2 UNIFORM(left, right) -- smps_sto.jl:113-127; null = every element is discrete
    const int *kind;          // [s]
    const double *par_a;      // [s] mean | left
    const double *par_b;      // [s] variance | right
};

#define SQLP_DELTA_COLS 64      // scenarios per block (half a tile)
#define SQLP_DELTA_SLAB 128     // row slots staged per pass
#define SQLP_DELTA_THREADS 512

// Row stride (doubles) of the staging buffer sh[column][slot]: >= slab and = 4 (mod 16), so that
// phase 1 (lanes along the slot axis) and phase 2 (lanes over 4 slots x 8 columns that are 0..3
// and 8..11 apart) both touch 16 distinct 8-byte bank pairs per half-warp.
__host__ __device__ __forceinline__ int delta_stride(int slab)
{
    return ((slab + 11) / 16) * 16 + 4;
}

template <bool SAMPLE>
__device__ __forceinline__ double realised_value(const DeltaTables &tb, const double *values,
                                                 long long g, long long g0, int e,
                                                 unsigned long long seed)
{
    if (SAMPLE) {
        const int kind = tb.kind ? tb.kind[e] : 0;
        if (kind != 0) {
            // open-interval uniform (z + 1/2) / 2^53 so that the normal quantile stays finite
            const double uo = u01_open(seed, (unsigned long long)g * tb.s + e);
            if (kind == 1) return fma(sqrt(tb.par_b[e]), normcdfinv(uo), tb.par_a[e]);
            return fma(tb.par_b[e] - tb.par_a[e], uo, tb.par_a[e]);
        }
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
// One block per 64-scenario half of a 128-aligned block of global ordinals.  Phase 1: one warp
// per scenario, lanes along the slot axis, so the realised values (contiguous per scenario) are
// read in 256-byte runs and all of a block's loads are in flight at once.  Phase 2: the half
// tile's part of every k-group (4 slots x 64 columns = 2 KB contiguous) is written in order.
template <bool SAMPLE>
__global__ void __launch_bounds__(SQLP_DELTA_THREADS, 2)
k_delta_build(DeltaTables tb, const double *__restrict__ values, long long g0, long long n_new,
              int rank, int world, int s_pad, double *__restrict__ D, double *__restrict__ dT,
              double *__restrict__ w, const double *__restrict__ w_batch, unsigned long long seed,
              unsigned long long wseed, double *__restrict__ DR)
{
    griddep_sync();
    extern __shared__ __align__(16) double sh[];   // [SQLP_DELTA_COLS][stride]
    const long long gblock = g0 / SQLP_TILE + (blockIdx.x >> 1);   // global 128-block
    if ((int)(gblock % world) != rank) return;
    const int half = blockIdx.x & 1;
    const long long gb0 = gblock * SQLP_TILE;
    const long long lo = max(gb0 + half * SQLP_DELTA_COLS, g0);
    const long long hi = min(gb0 + (half + 1) * SQLP_DELTA_COLS, g0 + n_new);
    if (lo >= hi) return;
    const int c0 = (int)(lo - gb0), c1 = (int)(hi - gb0);   // columns [c0, c1) of the tile
    const int cb = half * SQLP_DELTA_COLS;
    const long long ltile = gblock / world;                  // local tile index
    double *Dt = D + ltile * (long long)s_pad * SQLP_TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slab = min(s_pad, SQLP_DELTA_SLAB), stride = delta_stride(slab);

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

    for (int j0 = 0; j0 < s_pad; j0 += slab) {
        const int jn = min(slab, s_pad - j0);   // slots of this pass (a multiple of 4)
        // phase 1: the element and template value of a lane's slots do not depend on the scenario
        int el[SQLP_DELTA_SLAB / 32];
        double eb[SQLP_DELTA_SLAB / 32];
#pragma unroll
        for (int k = 0; k < SQLP_DELTA_SLAB / 32; ++k) {
            const int j = j0 + lane + 32 * k;
            el[k] = (lane + 32 * k < jn && j < tb.n_rows) ? tb.slot_elem[j] : -1;
            eb[k] = (el[k] >= 0) ? tb.elem_base[el[k]] : 0.0;
        }
        // every load of the block (64 scenarios x up to 128 slots over 16 warps) is issued before
        // the first use: 16 independent 8-byte loads per lane
        constexpr int CPW = SQLP_DELTA_COLS / (SQLP_DELTA_THREADS / 32);   // scenarios per warp
        if (SAMPLE) {
            // drawn values need no loads in flight: one scenario at a time keeps the registers (and
            // with them the resident blocks per SM) for the generator
#pragma unroll 1
            for (int i = 0; i < CPW; ++i) {
                const int c = cb + warp + i * (SQLP_DELTA_THREADS / 32);
                double *row = sh + (warp + i * (SQLP_DELTA_THREADS / 32)) * stride;
#pragma unroll
                for (int k = 0; k < SQLP_DELTA_SLAB / 32; ++k) {
                    if (lane + 32 * k >= jn) continue;
                    double d = 0.0;
                    if (el[k] >= 0 && c >= c0 && c < c1)
                        d = __dsub_rn(realised_value<true>(tb, values, gb0 + c, g0, el[k], seed), eb[k]);   // :114
                    row[lane + 32 * k] = d;
                }
            }
        } else {
        double v[CPW][SQLP_DELTA_SLAB / 32];
#pragma unroll
        for (int i = 0; i < CPW; ++i) {
            const int c = cb + warp + i * (SQLP_DELTA_THREADS / 32);
#pragma unroll
            for (int k = 0; k < SQLP_DELTA_SLAB / 32; ++k)
                v[i][k] = (el[k] >= 0 && c >= c0 && c < c1)
                              ? realised_value<SAMPLE>(tb, values, gb0 + c, g0, el[k], seed) : 0.0;
        }
#pragma unroll
        for (int i = 0; i < CPW; ++i) {
            double *row = sh + (warp + i * (SQLP_DELTA_THREADS / 32)) * stride;
#pragma unroll
            for (int k = 0; k < SQLP_DELTA_SLAB / 32; ++k)
                if (lane + 32 * k < jn) row[lane + 32 * k] = (el[k] >= 0) ? __dsub_rn(v[i][k], eb[k]) : 0.0;   // :114
        }
        }
        __syncthreads();
        // phase 2: offset o of a k-group's 512 doubles is cell P = o / 64, lane t = (o % 64) / 2,
        // h = o % 2, i.e. column (2 P + h) * 8 + t / 4 and slot t % 4 (common.cuh tile_off); the two
        // columns of a pair (h = 0, 1) are one 16-byte store
        const bool whole = (c0 == cb && c1 == cb + SQLP_DELTA_COLS);
        for (int q = threadIdx.x; q < (jn / 4) * 128; q += blockDim.x) {
            const int gq = q >> 7, o = cb * 4 + 2 * (q & 127);
            const int P = o >> 6, t = (o & 63) >> 1;
            const int c = (2 * P) * 8 + (t >> 2);            // h = 0 column; h = 1 is c + 8
            const double *src = sh + (c - cb) * stride + gq * 4 + (t & 3);
            double *dst = Dt + (long long)(j0 / 4 + gq) * 512 + o;
            if (whole) {
                *reinterpret_cast<double2 *>(dst) = make_double2(src[0], src[8 * stride]);
            } else {
                if (c >= c0 && c < c1) dst[0] = src[0];
                if (c + 8 >= c0 && c + 8 < c1) dst[1] = src[8 * stride];
            }
        }
        // the same values row-major [scenario][s_pad]: what the exact decision of the screening pass gathers
        if (DR) {
            for (int q = threadIdx.x; q < (c1 - c0) * jn; q += blockDim.x) {
                const int c = c0 + q / jn, jj = q - (c - c0) * jn;
                DR[(ltile * SQLP_TILE + c) * (long long)s_pad + j0 + jj] = sh[(c - cb) * stride + jj];
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
    griddep_sync();
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
    griddep_sync();
    for (int j = threadIdx.x; j < n_rows; j += blockDim.x) out[j] = Dtile[tile_off(c, j)];
}

}  // namespace sqlp
