// kernels_pool.cuh -- the device-resident dual-vertex pool (sdDualVertexSet).
//
// Reference behaviour reproduced bit for bit (src/sd_algorithm/dual_set.jl):
//   hash_dual_vector :46-53   sequential 1-norm, rounded to 16 significant bits, as UInt64
//   isequal          :24-40   same hash AND every element equal after rounding (fp !=)
//   push!            :84-93   first match in insertion order wins, else append
//
// HBM layout: pi[cap][m2] row-major full vertices, hash[cap] uint64.  The pool size K
// lives in device memory so pushes can be enqueued back to back without a host round trip.
#pragma once
#include "common.cuh"

namespace sqlp {

struct PushResult {
    int64_t index;
    int32_t inserted;
    int32_t pad;
};

// Scratch of the push kernel (one per pool).
struct PushScratch {
    unsigned long long hash;   // hash of the vector being pushed
    int match;                 // lowest stored slot equal to it, INT_MAX if none
    unsigned int done;         // blocks finished (last-block-commits pattern)
};

// hash_dual_vector alone (sqlp_pool_hash, debug / test): hash of the vector and its rounded copy.
__global__ void k_pool_prepare(const double *__restrict__ v, int m2, double *__restrict__ vr,
                               PushScratch *__restrict__ sc)
{
    griddep_sync();
    extern __shared__ double sh[];
    for (int j = threadIdx.x; j < m2; j += blockDim.x) {
        double x = v[j];
        sh[j] = fabs(x);
        vr[j] = round_sig16(x);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double mysum = 0.0;
        for (int j = 0; j < m2; ++j) mysum = __dadd_rn(mysum, sh[j]);
        sc->hash = (unsigned long long)__double_as_longlong(round_sig16(mysum));
    }
}

// One push in one launch: every block computes the hash of the new vector (the sequential 1-norm is one
// thread's chain of m2 additions -- the same few microseconds whether one block does it or all of them do it
// side by side, and a launch cheaper than handing it over from a kernel of its own), then the blocks scan
// the stored hashes, one per thread (grid-stride), and the last block to finish commits.
// `sc` must hold {match = INT_MAX, done = 0} on entry; the committing block restores that for the next push.
#define SQLP_PUSH_SMEM_DOUBLES 4096
__global__ void __launch_bounds__(256) k_pool_push(double *__restrict__ pi, unsigned long long *__restrict__ hash,
                                                   long long *__restrict__ d_K, int m2, const double *__restrict__ v,
                                                   PushScratch *__restrict__ sc, PushResult *__restrict__ result)
{
    griddep_sync();
    extern __shared__ double sh_abs[];
    __shared__ unsigned long long h_sh;
    __shared__ bool is_last;
    const bool in_smem = m2 <= SQLP_PUSH_SMEM_DOUBLES;
    if (in_smem)
        for (int j = threadIdx.x; j < m2; j += blockDim.x) sh_abs[j] = fabs(v[j]);
    __syncthreads();
    if (threadIdx.x == 0) {
        double mysum = 0.0;                               // :49-51 sequential, index order
        if (in_smem)
            for (int j = 0; j < m2; ++j) mysum = __dadd_rn(mysum, sh_abs[j]);
        else
            for (int j = 0; j < m2; ++j) mysum = __dadd_rn(mysum, fabs(v[j]));
        h_sh = (unsigned long long)__double_as_longlong(round_sig16(mysum));
    }
    __syncthreads();
    const long long K = *d_K;
    const unsigned long long h = h_sh;
    const int lane = threadIdx.x & 31;
    // the hash gate (:26), one stored hash per thread (coalesced); the rare hit is compared element by
    // element by the whole warp.  Few blocks: the scan is 8 K bytes, and every block costs one atomic on
    // the shared counter below.
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long kb = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); kb < K; kb += stride) {
        const long long mine = kb + lane;
        unsigned hits = __ballot_sync(0xffffffffu, mine < K && hash[mine] == h);
        while (hits) {
            const long long k = kb + (__ffs(hits) - 1);
            hits &= hits - 1;
            const double *row = pi + k * (long long)m2;
            bool same = true;
            for (int j = lane; j < m2; j += 32) {
                double r1 = round_sig16(v[j]);
                double r2 = round_sig16(row[j]);
                if (r1 != r2) same = false;               // :34  NaN != NaN
            }
            same = __all_sync(0xffffffffu, same);
            if (same && lane == 0) atomicMin(&sc->match, (int)k);
        }
    }

    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int t = atomicAdd(&sc->done, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int match = *((volatile int *)&sc->match);
    __syncthreads();                                      // everyone has read the match before it is reset
    if (threadIdx.x == 0) {
        sc->hash = h;
        sc->match = 0x7fffffff;
        sc->done = 0u;
    }
    if (match != 0x7fffffff) {
        if (threadIdx.x == 0) {
            result->index = match;
            result->inserted = 0;
        }
        return;
    }
    double *dst = pi + K * (long long)m2;                 // :91 append
    for (int j = threadIdx.x; j < m2; j += blockDim.x) dst[j] = v[j];
    if (threadIdx.x == 0) {
        hash[K] = h;
        result->index = K;
        result->inserted = 1;
        __threadfence();
        *d_K = K + 1;
    }
}

// Stochastic-row view of the pool in the contraction's fragment-major tile layout
// (common.cuh tile_off): vertex k is column k % 128 of tile k / 128, slot j < s_pad.
// Idempotent; run over [k_lo, *d_K) after pushes.
__global__ void k_view_sync(const double *__restrict__ pi, int m2, const int *__restrict__ s_rows,
                            int n_rows, int s_pad, double *__restrict__ piS, long long k_lo,
                            const long long *__restrict__ d_K)
{
    griddep_sync();
    const long long K = *d_K;
    const long long total = (K - k_lo) * n_rows;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        long long k = k_lo + t / n_rows;
        int j = (int)(t % n_rows);
        piS[(k >> 7) * (long long)s_pad * SQLP_TILE + tile_off((int)(k & 127), j)] = pi[k * m2 + s_rows[j]];
    }
}

// Per-epigraph vertex tables: rt[k][0] = rho_k = pi_k . rbar (index order),
// rt[k][1 + c] = tau_kc = sum over column c of Tbar (rows ascending) of T * pi_k[row]
// -- the per-column gather of `(transfer)' * dual`, epigraph.jl:141.
__global__ void k_epi_tables(const double *__restrict__ pi, int m2, const int *__restrict__ r_idx,
                             const double *__restrict__ r_val, int r_nnz,
                             const long long *__restrict__ T_colptr, const int *__restrict__ T_rowval,
                             const double *__restrict__ T_nzval, int n1, double *__restrict__ rt,
                             long long k_lo, const long long *__restrict__ d_K, int stage_rho)
{
    griddep_sync();
    extern __shared__ double prod[];                 // [r_nnz] when stage_rho
    const long long K = *d_K;
    const int RT = n1 + 1;
    for (long long k = k_lo + blockIdx.x; k < K; k += gridDim.x) {
        const double *row = pi + k * (long long)m2;
        if (stage_rho) {   // the products of rho, fetched by every thread side by side; the ordered chain below
            for (int q = threadIdx.x; q < r_nnz; q += blockDim.x) prod[q] = __dmul_rn(row[r_idx[q]], r_val[q]);
            __syncthreads();
        }
        for (int c = threadIdx.x; c < RT; c += blockDim.x) {
            double acc = 0.0;
            if (c == 0) {   // the non-zeros of rbar in index order: one ordered chain
                if (stage_rho)
                    for (int q = 0; q < r_nnz; ++q) acc = __dadd_rn(acc, prod[q]);
                else
                    for (int q = 0; q < r_nnz; ++q) acc = __dadd_rn(acc, __dmul_rn(row[r_idx[q]], r_val[q]));
            } else {
                for (long long q = T_colptr[c - 1]; q < T_colptr[c]; ++q)
                    acc = __dadd_rn(acc, __dmul_rn(T_nzval[q], row[T_rowval[q]]));
            }
            rt[k * RT + c] = acc;
        }
        if (stage_rho) __syncthreads();              // prod is rewritten for the block's next vertex
    }
}

}  // namespace sqlp
