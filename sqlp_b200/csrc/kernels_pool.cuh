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
#define SQLP_PUSH_BATCH 8              // vectors per launch
struct PushScratch {
    unsigned long long hash;           // hash of the (last) vector pushed, for sqlp_pool_hash
    int match[SQLP_PUSH_BATCH];        // lowest stored slot equal to vector i, INT_MAX if none
    unsigned int done;                 // blocks finished (last-block-commits pattern)
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

// Up to SQLP_PUSH_BATCH pushes in one launch, with the result of pushing them one after the other.
//   1. every block computes the hashes of all n vectors: the reference's hash is a SEQUENTIAL 1-norm, one
//      chain of m2 dependent additions per vector, so lane i of warp 0 walks the chain of vector i and the n
//      chains cost the time of one;
//   2. the blocks scan the stored hashes, one per thread (grid-stride), against all n hashes; the rare hit is
//      compared element by element by the whole warp (dual_set.jl:26-39) and the lowest equal slot kept;
//   3. the last block to finish resolves the pushes IN ORDER: vector i is a duplicate of the lowest equal
//      stored slot if there is one (stored slots are lower than any slot this batch appends, so first-match
//      order is kept), else of the first EARLIER vector of this batch that was appended and equals it, else it
//      is appended (dual_set.jl:84-93).
// `sc` must hold {match[] = INT_MAX, done = 0} on entry; the committing block restores that.
// Dynamic shared memory: n * m2 doubles when n * m2 <= SQLP_PUSH_SMEM_DOUBLES (the host batches accordingly),
// none for a single very long vector.
#define SQLP_PUSH_SMEM_DOUBLES 5632   // 44 KB: with the static arrays below 48 KB, no opt-in needed
__device__ __forceinline__ bool rounded_equal_block(const double *__restrict__ a, const double *__restrict__ b, int m2)
{
    bool same = true;
    for (int j = threadIdx.x; j < m2; j += blockDim.x)
        if (round_sig16(a[j]) != round_sig16(b[j])) same = false;   // :34  NaN != NaN
    return __syncthreads_and(same);
}

__global__ void __launch_bounds__(256) k_pool_push(double *__restrict__ pi, unsigned long long *__restrict__ hash,
                                                   long long *__restrict__ d_K, int m2, const double *__restrict__ v,
                                                   int n, PushScratch *__restrict__ sc, PushResult *__restrict__ result)
{
    griddep_sync();
    extern __shared__ double sh_abs[];
    __shared__ unsigned long long h_sh[SQLP_PUSH_BATCH];
    __shared__ long long slot_sh[SQLP_PUSH_BATCH];        // last block: slot a vector of this batch was appended to, or -1
    __shared__ bool is_last;
    const bool in_smem = (long long)n * m2 <= SQLP_PUSH_SMEM_DOUBLES;
    if (in_smem)
        for (int j = threadIdx.x; j < n * m2; j += blockDim.x) sh_abs[j] = fabs(v[j]);
    __syncthreads();
    if (threadIdx.x < n) {
        const int i = threadIdx.x;
        double mysum = 0.0;                               // :49-51 sequential, index order
        if (in_smem)
            for (int j = 0; j < m2; ++j) mysum = __dadd_rn(mysum, sh_abs[i * m2 + j]);
        else
            for (int j = 0; j < m2; ++j) mysum = __dadd_rn(mysum, fabs(v[(long long)i * m2 + j]));
        h_sh[i] = (unsigned long long)__double_as_longlong(round_sig16(mysum));
    }
    __syncthreads();
    const long long K = *d_K;
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long kb = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); kb < K; kb += stride) {
        const long long mine = kb + lane;
        const unsigned long long hk = mine < K ? hash[mine] : 0ull;
        for (int i = 0; i < n; ++i) {
            unsigned hits = __ballot_sync(0xffffffffu, mine < K && hk == h_sh[i]);   // :26 hash gate
            while (hits) {
                const long long k = kb + (__ffs(hits) - 1);
                hits &= hits - 1;
                const double *row = pi + k * (long long)m2, *vi = v + (long long)i * m2;
                bool same = true;
                for (int j = lane; j < m2; j += 32)
                    if (round_sig16(vi[j]) != round_sig16(row[j])) same = false;      // :34  NaN != NaN
                same = __all_sync(0xffffffffu, same);
                if (same && lane == 0) atomicMin(&sc->match[i], (int)k);
            }
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
    long long Kcur = K;
    for (int i = 0; i < n; ++i) {                         // the pushes, in order
        long long m = *((volatile int *)&sc->match[i]);
        if (m == 0x7fffffff) {
            m = -1;
            for (int e = 0; e < i && m < 0; ++e)          // earlier vectors of this batch that were appended
                if (slot_sh[e] >= 0 && h_sh[e] == h_sh[i] &&
                    rounded_equal_block(v + (long long)i * m2, v + (long long)e * m2, m2))
                    m = slot_sh[e];
        }
        __syncthreads();
        if (m >= 0) {
            if (threadIdx.x == 0) {
                slot_sh[i] = -1;
                result[i].index = m;
                result[i].inserted = 0;
            }
        } else {
            double *dst = pi + Kcur * (long long)m2;      // :91 append
            for (int j = threadIdx.x; j < m2; j += blockDim.x) dst[j] = v[(long long)i * m2 + j];
            if (threadIdx.x == 0) {
                hash[Kcur] = h_sh[i];
                slot_sh[i] = Kcur;
                result[i].index = Kcur;
                result[i].inserted = 1;
            }
            ++Kcur;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        sc->hash = h_sh[n - 1];
        for (int i = 0; i < SQLP_PUSH_BATCH; ++i) sc->match[i] = 0x7fffffff;
        sc->done = 0u;
        __threadfence();
        *d_K = Kcur;
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
