// kernels_reduce.cuh -- bias vectors, the deterministic weighted cut reduction, eval_dual.
//
// Reference: build_sasa_cut, src/sd_algorithm/epigraph.jl:134-143
//   p      = w_i / W
//   alpha += p * dot(pi*, rbar + delta_rhs_i)
//   beta  += -p * (Tbar + delta_T_i)' * pi*
//   val   += p * max_val_i
// Restated per scenario with the per-vertex tables rho_k = pi_k . rbar and
// tau_k = Tbar' pi_k (kernels_pool.cuh):
//   alpha_i = rho_k* + sum_{j in S} PiS[k*, j] * delta_rhs_i[j]
//   beta_i  = tau_k* + sum_{T elements e} delta_T_ie * pi*[row_e]  (added to column col_e)
// Accumulation is deterministic: one block per 128-scenario tile sums its scenarios in
// index order, block partials are summed in block order by a fixed two-level pass, rank
// partials (after the all-gather) in rank order.  No floating-point atomics anywhere.
#pragma once
#include "common.cuh"

namespace sqlp {

// base_x = rbar - Tbar * x   (subprob.jl:147).  SparseArrays' CSC product adds the entries of a
// row in column order starting from zero; a CSR copy of Tbar (columns ascending within a row)
// lets one thread per row do exactly that sum, so every row is bit-identical to the reference
// and no thread walks the whole matrix.  blockIdx.y selects the point.
__global__ void k_base(const double *__restrict__ rbar, int m2, int n1,
                       const int *__restrict__ R_ptr, const int *__restrict__ R_col,
                       const double *__restrict__ R_val, const double *__restrict__ x2,
                       double *__restrict__ base, int *__restrict__ flags)
{
    griddep_sync();
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *flags = 0;   // this call's "no argmax" bit
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m2) return;
    const double *x = x2 + (long long)blockIdx.y * n1;
    double y = 0.0;
    for (int q = R_ptr[j]; q < R_ptr[j + 1]; ++q) y = __dadd_rn(y, __dmul_rn(R_val[q], x[R_col[q]]));
    base[(long long)blockIdx.y * m2 + j] = __dsub_rn(rbar[j], y);
}

// bias_x[k] = dot(pi_k, base_x)  (the first dot of subprob.jl:155), one warp per vertex,
// lanes stride the row, fixed shuffle tree.  Slots K..kpad-1 get -inf so padded vertices
// of the last chunk can never win.
// With twins (kernels_pool.cuh) k runs over the view's columns and act[k] is the pool slot of the class's first vertex.
// FUSED: every block first builds base_x = rbar - Tbar x for its NX points in shared memory (one thread per row, the
// row's entries in column order -- the arithmetic of k_base, which is then not launched) and block 0 clears the
// call's "no argmax" bit; one launch less per cut formation, which is what small shapes are made of.
struct BaseArgs {
    const double *rbar;       // [m2]
    const int *R_ptr, *R_col; // CSR copy of Tbar, columns ascending within a row
    const double *R_val;
    const double *x2;         // [NX][n1]
    int n1;
    int *flags;
};

template <int NX, bool FUSED>
__global__ void k_bias(const double *__restrict__ pi, int m2, const double *__restrict__ base,
                       const long long *__restrict__ d_K, long long kpad, double *__restrict__ bias,
                       long long bias_stride, const int *__restrict__ act, BaseArgs b)
{
    griddep_sync();
    extern __shared__ double sbase[];            // [NX][m2] when FUSED
    if (FUSED) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *b.flags = 0;
        for (int q = threadIdx.x; q < NX * m2; q += blockDim.x) {
            const int x = q / m2, j = q % m2;
            const double *xx = b.x2 + (long long)x * b.n1;
            double y = 0.0;
            for (int t = b.R_ptr[j]; t < b.R_ptr[j + 1]; ++t) y = __dadd_rn(y, __dmul_rn(b.R_val[t], xx[b.R_col[t]]));
            sbase[q] = __dsub_rn(b.rbar[j], y);
        }
        __syncthreads();
        base = sbase;
    }
    const long long K = *d_K;
    const int lane = threadIdx.x & 31;
    const long long k = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= kpad) return;
    if (k >= K) {
        if (lane < NX) bias[lane * bias_stride + k] = -INFINITY;
        return;
    }
    const double *row = pi + (act ? (long long)act[k] : k) * (long long)m2;
    double s[NX];
#pragma unroll
    for (int x = 0; x < NX; ++x) s[x] = 0.0;
    for (int j = lane; j < m2; j += 32) {
        const double p = row[j];
#pragma unroll
        for (int x = 0; x < NX; ++x) s[x] = fma(p, base[(long long)x * m2 + j], s[x]);
    }
#pragma unroll
    for (int x = 0; x < NX; ++x) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) s[x] += __shfl_xor_sync(0xffffffffu, s[x], off);
        if (lane == 0) bias[x * bias_stride + k] = s[x];
    }
}

struct ReduceArgs {
    const double *D;          // [ntiles] fragment-major tiles: delta_rhs on S
    const double *dT;         // [n_local][n_T] or null
    const double *w;          // [n_local]
    const double *PiS;        // [nchunks] fragment-major tiles
    const double *rt;         // [K][n1 + 1]  (rho, tau), by POOL slot
    const int *act;           // view column -> pool slot (twins), or null: the identity
    const double *bias;       // [NX][bias_stride] of the contraction that produced best_*, or null
    long long bias_stride;    //   (null: some element perturbs Tbar, the winning dot is recomputed from D)
    const double *best_val;   // [NX][out_stride]
    const int *best_idx;      // [NX][out_stride]
    long long out_stride;
    long long n_local;
    int s_pad, n_rows, n1;
    double total_weight;
    // delta_T scatter: elements sorted by first-stage column
    int n_T;
    const int *tc_col;        // [n_T] column, ascending
    const int *tc_j;          // [n_T] row slot in S
    const int *tc_slot;       // [n_T] slot in dT
    double *partial;          // [ntiles][NX][n1 + 2]   (alpha, beta[n1], val)
    int nsub;                 // sub-ranges a tile is split into (1, 2, 4 or 8; > 1 needs nsub * NX * (n1 + 2) doubles of dynamic smem)
    int *flags;               // bit 0: some scenario had no argmax
};

template <int NX>
__global__ void __launch_bounds__(256) k_cut_partial(ReduceArgs a)
{
    griddep_sync();
    __shared__ double a_i[NX][SQLP_TILE];   // PiS[k*, :] . delta_rhs_i
    __shared__ double p_i[SQLP_TILE];
    __shared__ int k_i[NX][SQLP_TILE];        // winning view column (bias, pool view)
    __shared__ int kp_i[NX][SQLP_TILE];       // its pool slot (rho, tau table)
    const long long tile = blockIdx.x;
    const long long i0 = tile * SQLP_TILE;
    const int cnt = (int)min((long long)SQLP_TILE, a.n_local - i0);
    const double *Dt = a.D + tile * (long long)a.s_pad * SQLP_TILE;

    // phase A: thread per (x, scenario)
    for (int q = threadIdx.x; q < NX * SQLP_TILE; q += blockDim.x) {
        const int x = q / SQLP_TILE, c = q % SQLP_TILE;
        int k = -1;
        double acc = 0.0;
        bool need_dot = false;
        if (c < cnt) {
            k = a.best_idx[x * a.out_stride + i0 + c];
            if (k < 0) {
                atomicOr(a.flags, 1);
            } else if (a.bias) {
                // delta_T == 0: the contraction's winning score is bias_x[k] + PiS[k] . delta_rhs_i with
                // exactly this bias, so the dot is the score minus the bias (off by at most half an
                // ulp of the score) and neither D nor the pool view is read again
                const double sc = a.best_val[x * a.out_stride + i0 + c];
                acc = __dsub_rn(sc, a.bias[x * a.bias_stride + k]);
                // ... unless the score is so much larger than what alpha is made of (rho_k + dot) that half an
                // ulp of it shows at the 1e-10 level: |tau_k . x| >> |rho_k|, |dot| (x far from the data's
                // scale).  Then the dot is recomputed from D like in the delta_T != 0 case below.
                need_dot = fabs(sc) > 8192.0 * fmax(fabs(a.rt[(long long)(a.act ? a.act[k] : k) * (a.n1 + 1)]), fabs(acc));
            }
            if (k >= 0 && (!a.bias || need_dot)) {
                // slots in order; a k-group (4 slots) is 512 doubles further in both tiles and its
                // four slots sit 2 doubles apart, so the walk needs no index arithmetic.  Pad slots
                // are zero in both operands and add nothing.
                const double *P = a.PiS + ((long long)(k >> 7) * a.s_pad) * SQLP_TILE + tile_off(k & 127, 0);
                const double *Dc = Dt + tile_off(c, 0);
                const int ng = a.s_pad / 4;   // even: s_pad is a multiple of 8
                acc = 0.0;
                for (int g = 0; g < ng; g += 2, P += 1024, Dc += 1024) {
                    double pv[8], dv[8];   // sixteen loads in flight, then the ordered chain
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        pv[u] = P[(u >> 2) * 512 + (u & 3) * 2];
                        dv[u] = Dc[(u >> 2) * 512 + (u & 3) * 2];
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) acc = fma(pv[u], dv[u], acc);
                }
            }
            if (x == 0) p_i[c] = a.w[i0 + c] / a.total_weight;   // epigraph.jl:138
        }
        a_i[x][c] = acc;
        k_i[x][c] = k;
        kp_i[x][c] = (k >= 0 && a.act) ? a.act[k] : k;
    }
    __syncthreads();

    // phase B: a work item is (sub-range of the tile's scenarios, point x, output column); a.nsub sub-ranges
    // of 128 / nsub consecutive scenarios each.  Inside an item the scenarios are added in index order, the
    // table rows of eight scenarios being fetched before they are used so the gathers overlap; the nsub
    // sub-sums are then added in sub-range order.  The order is fixed by (N, nsub) alone, so the result is
    // deterministic; splitting the tile only shortens the chain of dependent gathers a thread walks (16
    // batches -> 2 at nsub = 8), which is what bounds this kernel at small N.  The loop body is branch
    // free: a scenario without argmax contributes fma(0, 0, sum) = sum.
    extern __shared__ double sub_sum[];              // [nsub][NX * NC] when nsub > 1
    const int NC = a.n1 + 2;
    const long long RT = a.n1 + 1;
    const int ncols = NX * NC, nsub = a.nsub, per = SQLP_TILE / nsub;
    for (int item = threadIdx.x; item < nsub * ncols; item += blockDim.x) {
        const int sub = item / ncols, q = item % ncols;
        const int x = q / NC, col = q % NC;
        const int c0 = sub * per, c1 = c0 + per;
        const bool is_alpha = (col == 0), is_val = (col == NC - 1);
        double sum = 0.0;
        if (!is_alpha && !is_val && a.n_T) {
            // beta column with delta_T elements landing in it (no shipped instance has any)
            int t0 = 0, t1 = 0;
            while (t0 < a.n_T && a.tc_col[t0] < col - 1) ++t0;
            t1 = t0;
            while (t1 < a.n_T && a.tc_col[t1] == col - 1) ++t1;
            for (int c = c0; c < min(c1, cnt); ++c) {
                const int k = k_i[x][c];
                if (k < 0) continue;
                double term = a.rt[(long long)kp_i[x][c] * RT + col];                  // :141
                for (int t = t0; t < t1; ++t) {
                    const double piv =
                        a.PiS[(long long)(k >> 7) * a.s_pad * SQLP_TILE + tile_off(k & 127, a.tc_j[t])];
                    term = fma(a.dT[(i0 + c) * (long long)a.n_T + a.tc_slot[t]], piv, term);
                }
                sum = fma(-p_i[c], term, sum);
            }
        } else {
            const double *src = is_val ? a.best_val + x * a.out_stride + i0 : a.rt + col;
            const long long stride = is_val ? 0 : RT;      // table row stride; the value column is per scenario
            const double sign = (is_alpha || is_val) ? 1.0 : -1.0;                     // :140-142
            for (int cb = c0; cb < c1; cb += 8) {          // k_i is -1 beyond cnt
                double tv[8], pw[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int c = cb + u, k = kp_i[x][c];
                    const bool ok = (k >= 0);
                    const double v = ok ? (is_val ? src[c] : src[(long long)k * stride]) : 0.0;
                    tv[u] = is_alpha ? v + (ok ? a_i[x][c] : 0.0) : v;
                    pw[u] = ok ? sign * p_i[c] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) sum = fma(pw[u], tv[u], sum);
            }
        }
        if (nsub == 1) a.partial[tile * ncols + q] = sum;
        else sub_sum[sub * ncols + q] = sum;
    }
    if (nsub == 1) return;
    __syncthreads();
    for (int q = threadIdx.x; q < ncols; q += blockDim.x) {
        double sum = sub_sum[q];
        for (int sub = 1; sub < nsub; ++sub) sum += sub_sum[sub * ncols + q];
        a.partial[tile * ncols + q] = sum;
    }
}

// Two ordered levels in one launch.  Level 1: block g adds the partial rows [g*group, min(n,(g+1)*group))
// -- nsub sub-ranges of group / nsub consecutive rows, each added in row order by its own thread (loads
// batched so they overlap), the sub-sums then added in sub-range order -- into out[g][q].  Level 2: the last
// block to finish adds the group sums in group order into fin[q] and re-arms the counter.  Which block runs
// level 2 depends on timing; what it computes does not.
__global__ void __launch_bounds__(256) k_sum_groups(const double *__restrict__ in, long long n, int group, int nsub,
                                                    int width, double *__restrict__ out,
                                                    unsigned int *__restrict__ counter, double *__restrict__ fin,
                                                    const int *__restrict__ flags_in, bool flag_behind_row)
{
    griddep_sync();
    extern __shared__ double sub_sum[];              // [nsub][width] when nsub > 1
    __shared__ bool is_last;
    const long long g = blockIdx.x;
    const int per = group / nsub;
    for (int item = threadIdx.x; item < nsub * width; item += blockDim.x) {
        const int sub = item / width, q = item % width;
        const long long p0 = g * group + (long long)sub * per, p1 = min(n, p0 + per);
        double s = 0.0;
        for (long long p = p0; p < p1; p += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (p + u < p1) ? in[(p + u) * width + q] : 0.0;
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];       // + 0.0 leaves a partial sum unchanged (never -0 here: s starts at +0)
        }
        if (nsub == 1) out[g * width + q] = s;
        else sub_sum[item] = s;
    }
    if (nsub > 1) {
        __syncthreads();
        for (int q = threadIdx.x; q < width; q += blockDim.x) {
            double s = sub_sum[q];
            for (int sub = 1; sub < nsub; ++sub) s += sub_sum[sub * width + q];
            out[g * width + q] = s;
        }
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int q = threadIdx.x; q < width; q += blockDim.x) {
        double s = 0.0;
        for (long long gg = 0; gg < gridDim.x; gg += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (gg + u < gridDim.x) ? __ldcg(out + (gg + u) * width + q) : 0.0;
#pragma unroll
            for (int u = 0; u < 8; ++u) s += v[u];
        }
        fin[q] = s;
    }
    // sharded job: the "no argmax" bit travels with the row (one more double), so that after the all-gather every
    // rank reaches the same verdict (k_cut_partial, the previous kernel of the stream, has set it)
    if (threadIdx.x == 0 && flag_behind_row) fin[width] = (double)(*flags_in & 1);
    if (threadIdx.x == 0) *counter = 0u;
}

// ---------------------------------------------------------------------------------------------------
// Cut reduction through per-vertex weight sums (delta_T == 0; k_cut_partial above stays for delta_T != 0 and
// for very small N).  build_sasa_cut (epigraph.jl:134-143) regrouped by the selected vertex:
//   c_k    = sum_{i : k*(i) = k} p_i
//   alpha  = sum_k c_k rho_k  + sum_i p_i (PiS[k*(i)] . delta_rhs_i)
//   beta   = - sum_k c_k tau_k
//   val    = sum_i p_i max_val_i
// k_cut_partial gathers one (rho, tau) row per scenario and point -- 2 N (n1 + 1) loads, 488 MB of L2 traffic
// at the bench shape for 44 MB of algorithmic bytes.  Here the table is read ONCE (K (n1 + 1) doubles).
// Deterministic without floating-point atomics:
//   k_cut_hist  block b owns a contiguous segment of scenarios and a private c[] in shared memory; warp w owns
//               the columns k = w (mod 32); every warp walks the segment in index order and adds, through ONE
//               lane, the weights of the scenarios whose winner it owns -- so c_b[k] is the index-ordered sum
//               over the segment, whatever the timing.  The segment's two scalar sums are a fixed tree.
//   k_cut_fold  c[k] = sum_b c_b[k] in block order; chunk of 256 columns -> sum_k c_k (rho_k, tau_k) in column
//               order; chunk 0 adds the scalar sums in block order.  k_sum_groups (above) adds the chunks.
#define SQLP_HIST_THREADS 1024
#define SQLP_HIST_SUB 2048          // scenarios staged in shared memory at a time
#define SQLP_FOLD_COLS 64

struct HistArgs {
    const double *w;          // [n_local]
    const double *rt;         // [K][n1 + 1] by pool slot
    const int *act;           // view column -> pool slot, or null
    const double *bias;       // [NX][bias_stride]
    long long bias_stride;
    const double *best_val;   // [NX][out_stride]
    const int *best_idx;      // [NX][out_stride] view columns
    long long out_stride;
    long long n_local;
    long long seg;            // scenarios per block (multiple of 128)
    const long long *d_Kv;    // columns of the view
    int kc;                   // columns the shared-memory histogram holds (>= *d_Kv, host upper bound)
    int n1;
    double total_weight;
    // recomputing the winning dot when the score - bias shortcut is too coarse (see k_cut_partial)
    const double *D, *PiS;
    int s_pad;
    double *cpart;            // [nblk][NX][kc]
    double *spart;            // [nblk][NX][2]   (sum p dot, sum p score)
    int *flags;
    // k_cut_hist_fx: the weight sums as integers (see there)
    long long *hfx;           // [NX][kc][3] limb sums, zero before the launch; null: cpart holds the sums
    double sc1, inv1;         // 2^(60 - eb) and its inverse
};

template <int NX>
__global__ void __launch_bounds__(SQLP_HIST_THREADS, 1) k_cut_hist(HistArgs a)
{
    griddep_sync();
    extern __shared__ __align__(16) unsigned char hist_raw[];
    double *c = reinterpret_cast<double *>(hist_raw);                    // [kc]
    double *ps = c + a.kc;                                               // [SUB] weights p_i
    int *ks = reinterpret_cast<int *>(ps + SQLP_HIST_SUB);               // [SUB] winning columns
    __shared__ double red[2][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long i0 = (long long)blockIdx.x * a.seg, i1 = min(a.n_local, i0 + a.seg);
    const long long RT = a.n1 + 1;
    for (int x = 0; x < NX; ++x) {
        for (int k = tid; k < a.kc; k += blockDim.x) c[k] = 0.0;
        double sdot = 0.0, sval = 0.0;                                   // this thread's scenarios, index order
        __syncthreads();
        for (long long b0 = i0; b0 < i1; b0 += SQLP_HIST_SUB) {
            const int cnt = (int)min((long long)SQLP_HIST_SUB, i1 - b0);
            for (int q = tid; q < cnt; q += blockDim.x) {
                const long long i = b0 + q;
                const int k = a.best_idx[x * a.out_stride + i];
                double p = 0.0;
                if (k < 0) {
                    atomicOr(a.flags, 1);
                } else {
                    p = a.w[i] / a.total_weight;                         // epigraph.jl:138
                    const double sc = a.best_val[x * a.out_stride + i];
                    double acc = __dsub_rn(sc, a.bias[x * a.bias_stride + k]);
                    if (fabs(sc) > 8192.0 * fmax(fabs(a.rt[(long long)(a.act ? a.act[k] : k) * RT]), fabs(acc))) {
                        const double *P = a.PiS + ((long long)(k >> 7) * a.s_pad) * SQLP_TILE + tile_off(k & 127, 0);
                        const double *Dc = a.D + (i >> 7) * (long long)a.s_pad * SQLP_TILE + tile_off((int)(i & 127), 0);
                        acc = 0.0;
                        for (int g = 0; g < a.s_pad / 4; ++g, P += 512, Dc += 512)
#pragma unroll
                            for (int u = 0; u < 4; ++u) acc = fma(P[u * 2], Dc[u * 2], acc);
                    }
                    sdot = fma(p, acc, sdot);                            // :140 (the part that is not rho)
                    sval = fma(p, sc, sval);                             // :142
                }
                ks[q] = k;
                ps[q] = p;
            }
            __syncthreads();
            // warp w adds the weights of the winners k = w (mod 32), in scenario order, through lane 0
            for (int g = 0; g < cnt; g += 32) {
                const int q = g + lane;
                const int k = q < cnt ? ks[q] : -1;
                const double p = q < cnt ? ps[q] : 0.0;
                const bool mine = k >= 0 && (k & 31) == warp;
                unsigned own = __ballot_sync(0xffffffffu, mine);
                while (own) {
                    // the owned scenarios of this group that chose the same column are added up first (in lane
                    // order, in a register) and reach shared memory as ONE update: when the winners concentrate on
                    // a few vertices, a chain of dependent shared-memory updates would serialise the block
                    const int kk = __shfl_sync(0xffffffffu, k, __ffs(own) - 1);
                    unsigned same = __ballot_sync(0xffffffffu, mine && k == kk);
                    own &= ~same;
                    double sum = 0.0;
                    if (__popc(same) <= 4) {                 // a few: one after the other, in lane order
                        while (same) {
                            sum += __shfl_sync(0xffffffffu, p, __ffs(same) - 1);
                            same &= same - 1;
                        }
                    } else {                                 // many: the fixed butterfly over the 32 lanes (non-members add +0)
                        sum = ((same >> lane) & 1u) ? p : 0.0;
#pragma unroll
                        for (int off = 16; off >= 1; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
                    }
                    if (lane == 0) c[kk] += sum;
                }
            }
            __syncthreads();
        }
        // the block's scalar sums: lanes, then warps, in a fixed tree
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            sdot += __shfl_xor_sync(0xffffffffu, sdot, off);
            sval += __shfl_xor_sync(0xffffffffu, sval, off);
        }
        if (lane == 0) { red[0][warp] = sdot; red[1][warp] = sval; }
        __syncthreads();
        if (warp == 0) {
            double u = red[0][lane], v = red[1][lane];
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                u += __shfl_xor_sync(0xffffffffu, u, off);
                v += __shfl_xor_sync(0xffffffffu, v, off);
            }
            if (lane == 0) {
                a.spart[((long long)blockIdx.x * NX + x) * 2] = u;
                a.spart[((long long)blockIdx.x * NX + x) * 2 + 1] = v;
            }
        }
        double *dst = a.cpart + ((long long)blockIdx.x * NX + x) * a.kc;
        for (int k = tid; k < a.kc; k += blockDim.x) dst[k] = c[k];
        __syncthreads();
    }
}

// The weight sums in fixed point.  k_cut_hist keeps the sums deterministic by giving every column to one warp that
// adds its scenarios in order -- every warp reads every staged scenario, 32 x the work, and the kernel ran at 4 % of
// what its bytes allow.  Integers add in any order to the same bits, so here every thread handles its own scenario
// and the sums are atomic integer adds.  With |p_i| <= max|w| / |total| < 2^eb (both known on the host) and
// f_i = rint(p_i 2^(60 - eb)), |f_i| < 2^60, a weight is three signed limbs of 20 bits, f = a 2^40 + b 2^20 + c: a
// block of at most 2 047 scenarios adds limbs into 32-bit words of its shared-memory table (native atomics, no
// overflow: 2 047 x 2^20 < 2^31), its non-zero words go to 64-bit global sums, and the fold forms
// c_k = (A_k 2^40 + B_k 2^20 + C_k) 2^(eb - 60): every weight to 2^-60 of the largest one (the reference's sequential
// FP64 sum carries 2^-53 of the running sum), the same bits whatever the block partition and the order of the adds.
// The table holds the columns the view really has (*d_Kv): both points at once if they fit the launch's shared
// memory, else one point per pass.  The scalar sums stay as in k_cut_hist (fixed tree per block, blocks in order).
#define SQLP_HISTFX_SEG 1792          // scenarios per block: 14 tiles, < 2 048
template <int NX>
__global__ void __launch_bounds__(SQLP_HIST_THREADS, 1) k_cut_hist_fx(HistArgs a, int smem_bytes)
{
    griddep_sync();
    extern __shared__ __align__(16) unsigned char hist_raw[];
    int *H = reinterpret_cast<int *>(hist_raw);                                   // [xp][Kv][3]
    __shared__ double red[2][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long i0 = (long long)blockIdx.x * a.seg, i1 = min(a.n_local, i0 + a.seg);
    const long long RT = a.n1 + 1;
    const int Kv = (int)min(*a.d_Kv, (long long)a.kc);
    const int xp = ((long long)NX * Kv * 12 <= smem_bytes) ? NX : 1;               // points per pass
    for (int x0 = 0; x0 < NX; x0 += xp) {
        for (int t = tid; t < xp * Kv * 3; t += blockDim.x) H[t] = 0;
        double sdot[NX], sval[NX];
#pragma unroll
        for (int x = 0; x < NX; ++x) { sdot[x] = 0.0; sval[x] = 0.0; }
        __syncthreads();
        for (long long i = i0 + tid; i < i1; i += blockDim.x) {              // this thread's scenarios, index order
            const double p = a.w[i] / a.total_weight;                            // epigraph.jl:138
            const long long f = __double2ll_rn(p * a.sc1);
            const long long m = f < 0 ? -f : f;
            const int sg = f < 0 ? -1 : 1;
            const int la = sg * (int)(m >> 40), lb = sg * (int)((m >> 20) & 0xFFFFF), lc = sg * (int)(m & 0xFFFFF);
#pragma unroll
            for (int x = 0; x < NX; ++x) {
                if (x < x0 || x >= x0 + xp) continue;
                const int k = a.best_idx[x * a.out_stride + i];
                if (k < 0 || k >= Kv) {
                    atomicOr(a.flags, 1);
                    continue;
                }
                const double sc = a.best_val[x * a.out_stride + i];
                double acc = __dsub_rn(sc, a.bias[x * a.bias_stride + k]);
                if (fabs(sc) > 8192.0 * fmax(fabs(a.rt[(long long)(a.act ? a.act[k] : k) * RT]), fabs(acc))) {
                    const double *P = a.PiS + ((long long)(k >> 7) * a.s_pad) * SQLP_TILE + tile_off(k & 127, 0);
                    const double *Dc = a.D + (i >> 7) * (long long)a.s_pad * SQLP_TILE + tile_off((int)(i & 127), 0);
                    acc = 0.0;
                    for (int g = 0; g < a.s_pad / 4; ++g, P += 512, Dc += 512)
#pragma unroll
                        for (int u = 0; u < 4; ++u) acc = fma(P[u * 2], Dc[u * 2], acc);
                }
                sdot[x] = fma(p, acc, sdot[x]);                                  // :140 (the part that is not rho)
                sval[x] = fma(p, sc, sval[x]);                                   // :142
                int *h = H + ((size_t)(x - x0) * Kv + k) * 3;
                if (la) atomicAdd(h, la);
                if (lb) atomicAdd(h + 1, lb);
                if (lc) atomicAdd(h + 2, lc);
            }
        }
        __syncthreads();
        for (int t = tid; t < xp * Kv * 3; t += blockDim.x) {
            const int v = H[t];
            if (v) {
                const int xx = t / (Kv * 3), rem = t - xx * (Kv * 3);
                atomicAdd(reinterpret_cast<unsigned long long *>(a.hfx) + ((size_t)(x0 + xx) * a.kc) * 3 + rem,
                          (unsigned long long)(long long)v);
            }
        }
        // the block's scalar sums: lanes, then warps, in a fixed tree
#pragma unroll
        for (int x = 0; x < NX; ++x) {
            if (x < x0 || x >= x0 + xp) continue;
            double u = sdot[x], v = sval[x];
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                u += __shfl_xor_sync(0xffffffffu, u, off);
                v += __shfl_xor_sync(0xffffffffu, v, off);
            }
            if (lane == 0) { red[0][warp] = u; red[1][warp] = v; }
            __syncthreads();
            if (warp == 0) {
                u = red[0][lane];
                v = red[1][lane];
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    u += __shfl_xor_sync(0xffffffffu, u, off);
                    v += __shfl_xor_sync(0xffffffffu, v, off);
                }
                if (lane == 0) {
                    a.spart[((long long)blockIdx.x * NX + x) * 2] = u;
                    a.spart[((long long)blockIdx.x * NX + x) * 2 + 1] = v;
                }
            }
            __syncthreads();
        }
    }
}

// part[chunk][x][col]: col 0 = sum_k c_k rho_k (+ the blocks' sum p dot on chunk 0), 1..n1 = -sum_k c_k tau_kj,
// n1 + 1 = the blocks' sum p score (chunk 0 only).  One block per chunk of 256 columns.
template <int NX>
__global__ void __launch_bounds__(256) k_cut_fold(HistArgs a, int nblk, double *__restrict__ part)
{
    griddep_sync();
    __shared__ double ck[NX][SQLP_FOLD_COLS];
    __shared__ int kp[SQLP_FOLD_COLS];
    const long long Kv = *a.d_Kv;
    const int k0 = blockIdx.x * SQLP_FOLD_COLS;
    const int NC = a.n1 + 2;
    const long long RT = a.n1 + 1;
    for (int q = threadIdx.x; q < NX * SQLP_FOLD_COLS; q += blockDim.x) {
        const int x = q / SQLP_FOLD_COLS, k = k0 + q % SQLP_FOLD_COLS;
        double s = 0.0;
        if (k < Kv && k < a.kc && a.hfx) {
            const long long *hh = a.hfx + ((long long)x * a.kc + k) * 3;       // limb sums: exact as doubles (< 2^53)
            s = ((double)hh[0] * 0x1p40 + (double)hh[1] * 0x1p20 + (double)hh[2]) * a.inv1;
        } else if (k < Kv && k < a.kc) {
            const double *src = a.cpart + (long long)x * a.kc + k;
            const long long bs = (long long)NX * a.kc;
            for (int b = 0; b < nblk; b += 8) {              // block order; eight loads in flight
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = b + u < nblk ? src[(b + u) * bs] : 0.0;
#pragma unroll
                for (int u = 0; u < 8; ++u) s += v[u];       // + 0.0 changes nothing (s starts at +0)
            }
        }
        ck[x][q % SQLP_FOLD_COLS] = s;
        if (x == 0) kp[q % SQLP_FOLD_COLS] = k < Kv ? (a.act ? a.act[k] : k) : -1;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < NX * NC; q += blockDim.x) {
        const int x = q / NC, col = q % NC;
        double s = 0.0;
        if (col <= a.n1) {
            // column order.  The table row of a vertex nobody selected is loaded (from a valid address) but not
            // used: its entries may be Inf / NaN, and 0 * Inf is not 0.
            for (int u0 = 0; u0 < SQLP_FOLD_COLS; u0 += 8) {
                double tv[8], cw[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int k = kp[u0 + u];
                    cw[u] = ck[x][u0 + u];
                    const double t = a.rt[(long long)max(k, 0) * RT + col];
                    tv[u] = (k >= 0 && cw[u] != 0.0) ? t : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) s = fma(cw[u], tv[u], s);
            }
            if (col > 0) s = -s;                                                                  // :141
        }
        if (blockIdx.x == 0 && (col == 0 || col == NC - 1)) {
            double t = 0.0;
            for (int b = 0; b < nblk; ++b) t += a.spart[((long long)b * NX + x) * 2 + (col == 0 ? 0 : 1)];
            s += t;
        }
        part[(long long)blockIdx.x * NX * NC + q] = s;
    }
}

// eval_dual, subprob.jl:128-131, in the reference's operation order (single thread):
//   dot(dual, (rbar + delta_rhs) - (Tbar + delta_T) * x)
struct EvalArgs {
    const double *pi_row;     // [m2]
    const double *rbar;       // [m2]
    const long long *T_colptr;
    const int *T_rowval;
    const double *T_nzval;
    int m2, n1, s_pad, n_rows, n_T;
    const int *s_rows;        // [n_rows] stage-2 row of slot j
    const double *Dtile;      // base of the scenario's D tile
    int dcol;                 // its column in the tile
    const double *dTrow;      // [n_T] or null
    // delta_T elements sorted by (col, row): position in CSC order
    const int *mc_col;
    const int *mc_row;
    const int *mc_slot;
    const double *x;
    double *out;
    double *scratch;          // [2 * m2]
};

__global__ void k_eval_dual(EvalArgs a)
{
    griddep_sync();
    if (threadIdx.x || blockIdx.x) return;
    double *y = a.scratch, *r = a.scratch + a.m2;
    for (int j = 0; j < a.m2; ++j) { y[j] = 0.0; r[j] = a.rbar[j]; }
    for (int j = 0; j < a.n_rows; ++j)
        r[a.s_rows[j]] = __dadd_rn(a.rbar[a.s_rows[j]], a.Dtile[tile_off(a.dcol, j)]);
    int t = 0;
    for (int c = 0; c < a.n1; ++c) {   // merged (Tbar + delta_T) * x, rows ascending per column
        const double xc = a.x[c];
        long long q = a.T_colptr[c], qe = a.T_colptr[c + 1];
        while (t < a.n_T && a.mc_col[t] < c) ++t;
        while (q < qe || (t < a.n_T && a.mc_col[t] == c)) {
            const int rq = (q < qe) ? a.T_rowval[q] : 0x7fffffff;
            const int rt = (t < a.n_T && a.mc_col[t] == c) ? a.mc_row[t] : 0x7fffffff;
            if (rq < rt) {
                y[rq] = __dadd_rn(y[rq], __dmul_rn(a.T_nzval[q], xc)); ++q;
            } else if (rt < rq) {
                y[rt] = __dadd_rn(y[rt], __dmul_rn(a.dTrow[a.mc_slot[t]], xc)); ++t;
            } else {
                y[rq] = __dadd_rn(y[rq], __dmul_rn(__dadd_rn(a.T_nzval[q], a.dTrow[a.mc_slot[t]]), xc));
                ++q; ++t;
            }
        }
    }
    double s = 0.0;
    for (int j = 0; j < a.m2; ++j) s = __dadd_rn(s, __dmul_rn(a.pi_row[j], __dsub_rn(r[j], y[j])));
    *a.out = s;
}

// Rank-ordered sum of all-gathered partials: out[q] = sum_r in[r * stride + q], r = 0..world-1, q < width; the
// double behind each rank's row is its "no argmax" bit, OR-ed into flags_out so that every rank raises (or not)
// together.
__global__ void k_rank_sum(const double *__restrict__ in, int world, long long stride, int width,
                           double *__restrict__ out, int *__restrict__ flags_out)
{
    griddep_sync();
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q > width) return;
    double s = 0.0;
    for (int r = 0; r < world; ++r) s += in[(long long)r * stride + q];
    if (q < width) out[q] = s;
    else *flags_out = s > 0.0 ? 1 : 0;
}

}  // namespace sqlp
