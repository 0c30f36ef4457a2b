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
                            int n_rows, int s_pad, double *__restrict__ piS, double *__restrict__ piR, long long k_lo,
                            const long long *__restrict__ d_K)
{
    griddep_sync();
    const long long K = *d_K;
    const long long total = (K - k_lo) * n_rows;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        long long k = k_lo + t / n_rows;
        int j = (int)(t % n_rows);
        const double val = pi[k * m2 + s_rows[j]];
        piS[(k >> 7) * (long long)s_pad * SQLP_TILE + tile_off((int)(k & 127), j)] = val;
        piR[k * (long long)s_pad + j] = val;             // row-major copy for the exact decision's gathers
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

// ---------------------------------------------------------------------------------------------------
// Score-equivalent vertices ("twins").  A vertex enters a score only through the rows where
//   rbar != 0, or Tbar has an entry, or a random element lives            (the RELEVANT rows):
//   score_x[k, i] = pi_k . (rbar - Tbar x) + pi_k|_S . d_i,   rho_k = pi_k . rbar,   tau_k = Tbar' pi_k.
// On every other row (rbar - Tbar x) is exactly +0.0 for every x.  Degenerate stage-2 LPs return many optimal
// duals that differ ONLY there (storm: 16 384 harvested vertices are 3 599 classes): distinct for the
// reference's dedup rule (dual_set.jl:24-40 compares all rows), identical -- bit for bit -- in every score
// and every cut coefficient.  argmax_procedure keeps the FIRST maximum (subprob.jl:156), so a later twin can
// never be selected: leaving it out of the sweep changes no index, no value and no cut.
// The pool view therefore holds one column per CLASS: act[v] = pool index of the first vertex of class v,
// ascending, so "first view slot" = "first pool index".  A vertex with a non-finite entry anywhere is kept
// apart (never a representative, never shadowed): 0 * Inf is NaN, not zero.
// Everything is decided on the device (pushes are enqueued, not awaited) and deterministically: the
// representative of a class is its LOWEST pool index (atomicMin), whatever the thread order.
struct TwinState {
    long long synced;      // pool vertices [0, synced) have been classified
    long long Kv;          // classes = columns of the view
    long long Kv_prev;     // Kv before the last classification (the fill kernels work on [Kv_prev, Kv))
};
#define SQLP_TWIN_EMPTY 0xFFFFFFFFFFFFFFFFull

__device__ __forceinline__ unsigned long long twin_mix(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// One warp per new vertex: hash of its relevant rows (order-independent sum of position-keyed element hashes),
// finiteness of ALL its rows, and its entry in the table (lowest pool index per hash value).
__global__ void k_twin_hash(const double *__restrict__ pi, int m2, const int *__restrict__ rel, int n_rel,
                            const long long *__restrict__ d_K, const TwinState *__restrict__ st,
                            unsigned long long *__restrict__ hk, unsigned long long *__restrict__ tkey,
                            int *__restrict__ trep, unsigned int tmask)
{
    griddep_sync();
    const long long K = *d_K, lo = st->synced;
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long k = lo + (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < K; k += nw) {
        const double *row = pi + k * (long long)m2;
        bool fin = true;
        for (int j = lane; j < m2; j += 32) fin = fin && isfinite(row[j]);
        unsigned long long h = 0;
        for (int q = lane; q < n_rel; q += 32)
            h += twin_mix((unsigned long long)__double_as_longlong(row[rel[q]]) ^ ((unsigned long long)(q + 1) * 0xD6E8FEB86659FD93ULL));
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) h += __shfl_xor_sync(0xffffffffu, h, off);
        fin = __all_sync(0xffffffffu, fin);
        h &= ~1ull;                                     // bit 0 of the stored word says "finite"
        if (h == (SQLP_TWIN_EMPTY & ~1ull)) h = 2;
        if (lane == 0) {
            hk[k] = h | (fin ? 1ull : 0ull);
            if (fin) {
                unsigned int p = (unsigned int)(h >> 17) & tmask;
                for (;;) {
                    const unsigned long long old = atomicCAS(tkey + p, SQLP_TWIN_EMPTY, h);
                    if (old == SQLP_TWIN_EMPTY || old == h) { atomicMin(trep + p, (int)k); break; }
                    p = (p + 1) & tmask;
                }
            }
        }
    }
}

// One warp per new vertex: is it the first of its class?  flag[k - synced] = 1 keeps it.
__global__ void k_twin_mark(const double *__restrict__ pi, int m2, const int *__restrict__ rel, int n_rel,
                            const long long *__restrict__ d_K, const TwinState *__restrict__ st,
                            const unsigned long long *__restrict__ hk, const unsigned long long *__restrict__ tkey,
                            const int *__restrict__ trep, unsigned int tmask, unsigned char *__restrict__ flag)
{
    griddep_sync();
    const long long K = *d_K, lo = st->synced;
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long k = lo + (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < K; k += nw) {
        const unsigned long long w = hk[k];
        bool keep = true;
        if (w & 1ull) {
            const unsigned long long h = w & ~1ull;
            unsigned int p = (unsigned int)(h >> 17) & tmask;
            while (tkey[p] != h) p = (p + 1) & tmask;    // inserted by k_twin_hash: always found
            const int r = trep[p];
            if (r != (int)k) {
                // same hash, lower index: a twin iff the relevant rows agree bit for bit (else a hash collision)
                const double *a = pi + k * (long long)m2, *b = pi + (long long)r * m2;
                bool same = true;
                for (int q = lane; q < n_rel; q += 32)
                    same = same && (__double_as_longlong(a[rel[q]]) == __double_as_longlong(b[rel[q]]));
                keep = !__all_sync(0xffffffffu, same);
            }
        }
        if (lane == 0) flag[k - lo] = keep ? 1 : 0;
    }
}

// One block: the kept vertices, in pool order, get the next view slots.
__global__ void __launch_bounds__(1024) k_twin_compact(const long long *__restrict__ d_K, TwinState *__restrict__ st,
                                                       const unsigned char *__restrict__ flag, int *__restrict__ act)
{
    griddep_sync();
    __shared__ int wsum[32], wexcl[32];
    __shared__ int total_sh;
    __shared__ long long base_sh;
    const long long K = *d_K, lo = st->synced;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base_sh = st->Kv;
    __syncthreads();
    for (long long k0 = lo; k0 < K; k0 += blockDim.x) {
        const long long k = k0 + threadIdx.x;
        const int f = (k < K) ? (int)flag[k - lo] : 0;
        int incl = f;                                    // inclusive scan inside the warp
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {                                 // scan of the 32 warp totals
            const int v = wsum[lane];
            int ws = v;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, ws, off);
                if (lane >= off) ws += t;
            }
            wexcl[lane] = ws - v;
            if (lane == 31) total_sh = ws;
        }
        __syncthreads();
        const long long base = base_sh;
        if (f) act[base + wexcl[warp] + incl - 1] = (int)k;
        __syncthreads();
        if (threadIdx.x == 0) base_sh = base + total_sh;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        st->Kv_prev = st->Kv;
        st->Kv = base_sh;
        st->synced = K;
    }
}

// View columns [Kv_prev, Kv) from the pool rows of their representatives (same tile layout as k_view_sync).
__global__ void k_view_fill(const double *__restrict__ pi, int m2, const int *__restrict__ s_rows, int n_rows, int s_pad,
                            double *__restrict__ piS, double *__restrict__ piR, const TwinState *__restrict__ st,
                            const int *__restrict__ act)
{
    griddep_sync();
    const long long v0 = st->Kv_prev, v1 = st->Kv;
    const long long total = (v1 - v0) * n_rows;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long v = v0 + t / n_rows;
        const int j = (int)(t % n_rows);
        const double val = pi[(long long)act[v] * m2 + s_rows[j]];
        piS[(v >> 7) * (long long)s_pad * SQLP_TILE + tile_off((int)(v & 127), j)] = val;
        piR[v * (long long)s_pad + j] = val;
    }
}

// argmax_procedure's indices back from view slots to pool slots.
__global__ void k_unmap_idx(int *__restrict__ idx, long long n, const int *__restrict__ act)
{
    griddep_sync();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int v = idx[i];
        if (v >= 0) idx[i] = act[v];
    }
}

}  // namespace sqlp
