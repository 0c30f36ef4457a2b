// kernels_contract.cuh -- the fused fp64 contraction + argmax (the hot kernel).
//
// Reference: the double loop of argmax_procedure, src/sd_algorithm/subprob.jl:148-166:
//   for each scenario i, for each pool vertex k (insertion order):
//       v = dot(pi_k, base) + dot(pi_k, delta_i);  strict '>' keeps the FIRST maximum.
// Restated as  score[k, i] = bias_x[k] + sum_{j in S} PiS[k, j] * d_i[j]  with
// bias_x[k] = pi_k . (rbar - Tbar x) over all m2 rows (kernels_reduce.cuh) and the second
// dot restricted to the stochastic rows S, the only rows where delta_i is non-zero.
// When no random element perturbs Tbar, d_i does not depend on x and the candidate and
// incumbent points (NX = 2) share ONE contraction and differ only in the bias.
//
// Shape: a [N x s] by [s x K] GEMM in fp64 whose [N x K] result never leaves registers.
//   CTA tile 128 scenarios x 128 vertices, 256 threads, 8 x 8 accumulators per thread,
//   operands streamed L2 -> shared memory in 8-row slabs by a 4-stage cp.async pipeline
//   that runs continuously across slabs, vertex chunks and scenario tiles.  Both operands
//   are stored in HBM already in [tile][j][128] order (kernels_delta.cuh, kernels_pool.cuh),
//   so every slab is one contiguous 8 KB block and every shared-memory read in the inner
//   loop is a conflict-free, broadcast LDS.128.
//   After the last slab of a vertex chunk the epilogue adds the bias and folds the 64
//   scores of each thread into per-thread running (max, argmax); after the last chunk of a
//   scenario tile the 16 threads that share a scenario row merge with warp shuffles and
//   one shared-memory pass, comparing on (value desc, index asc) so the first index wins.
//   fp64 has no tcgen05 path (the 5th-gen tensor cores stop at tf32); the roofline is the
//   DFMA pipe: 2 * s flop per (scenario, vertex) evaluation.
#pragma once
#include "common.cuh"

namespace sqlp {

#define SQLP_CT_THREADS 256
#define SQLP_CT_STAGES 4

template <int NX>
struct ContractSmem {
    static constexpr int kStageDoubles = 2 * SQLP_BK * SQLP_TILE + NX * SQLP_TILE;
    static constexpr int kRedDoubles = 4 * SQLP_TILE * NX;  // merge buffers (value)
    static constexpr size_t bytes()
    {
        return sizeof(double) * (SQLP_CT_STAGES * kStageDoubles + kRedDoubles) +
               sizeof(int) * (4 * SQLP_TILE * NX);
    }
};

struct ContractArgs {
    const double *D;        // [ntiles][s_pad][128] scenario deltas on the stochastic rows
    const double *PiS;      // [nchunks][s_pad][128] pool restricted to the stochastic rows
    const double *bias;     // [NX][bias_stride]; -inf for k >= K
    long long bias_stride;
    const long long *d_K;   // pool size (device resident)
    int s_pad;              // multiple of SQLP_BK
    int ntiles;
    long long n_local;      // scenarios held by this rank
    double *best_val;       // [NX][out_stride]
    int *best_idx;          // [NX][out_stride]
    long long out_stride;
};

__device__ __forceinline__ bool better(double ov, int oi, double v, int i)
{
    return (ov > v) || (ov == v && (unsigned)oi < (unsigned)i);
}

template <int NX>
__global__ void __launch_bounds__(SQLP_CT_THREADS, 1) k_contract_argmax(ContractArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *stages = reinterpret_cast<double *>(smem_raw);
    double *red_val = stages + SQLP_CT_STAGES * ContractSmem<NX>::kStageDoubles;
    int *red_idx = reinterpret_cast<int *>(red_val + ContractSmem<NX>::kRedDoubles);

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wy = warp >> 2, wx = warp & 3;   // warp grid 2 (scenarios) x 4 (vertices)
    const int ly = lane >> 2, lx = lane & 3;   // lane grid 8 x 4
    const int a_off = wy * 64 + ly * 2;        // + n * 16, n = 0..3  (two scenarios each)
    const int b_off = wx * 32 + lx * 2;        // + m * 8,  m = 0..3  (two vertices each)

    const long long K = *a.d_K;
    const int nchunks = (int)((K + SQLP_TILE - 1) / SQLP_TILE);
    const int nslab = a.s_pad / SQLP_BK;
    const int my_tiles = (a.ntiles > (int)blockIdx.x)
                             ? (a.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x
                             : 0;

    if (nchunks == 0) {  // empty pool: nothing beats -Inf (subprob.jl:151)
        for (int t = 0; t < my_tiles; ++t) {
            long long tile = blockIdx.x + (long long)t * gridDim.x;
            for (int q = tid; q < SQLP_TILE * NX; q += SQLP_CT_THREADS) {
                long long i = tile * SQLP_TILE + (q % SQLP_TILE);
                if (i < a.n_local) {
                    a.best_val[(q / SQLP_TILE) * a.out_stride + i] = -INFINITY;
                    a.best_idx[(q / SQLP_TILE) * a.out_stride + i] = -1;
                }
            }
        }
        return;
    }

    const long long total = (long long)my_tiles * nchunks * nslab;
    const size_t slab_doubles = (size_t)SQLP_BK * SQLP_TILE;
    const size_t tile_doubles = (size_t)a.s_pad * SQLP_TILE;

    // producer cursor
    int ld_slab = 0, ld_chunk = 0, ld_t = 0;
    long long ld_it = 0;
    auto issue_load = [&]() {
        if (ld_it < total) {
            double *st = stages + (ld_it % SQLP_CT_STAGES) * ContractSmem<NX>::kStageDoubles;
            const long long tile = blockIdx.x + (long long)ld_t * gridDim.x;
            const double *gA = a.D + tile * tile_doubles + ld_slab * slab_doubles;
            const double *gB = a.PiS + (size_t)ld_chunk * tile_doubles + ld_slab * slab_doubles;
#pragma unroll
            for (int q = 0; q < 2; ++q) {   // 512 16-byte pieces per operand, 256 threads
                int piece = tid + q * SQLP_CT_THREADS;
                cp_async16(st + piece * 2, gA + piece * 2);
                cp_async16(st + slab_doubles + piece * 2, gB + piece * 2);
            }
            if (ld_slab == nslab - 1 && tid < NX * SQLP_TILE / 2) {
                int x = tid / (SQLP_TILE / 2), p = tid % (SQLP_TILE / 2);
                cp_async16(st + 2 * slab_doubles + x * SQLP_TILE + p * 2,
                           a.bias + x * a.bias_stride + (size_t)ld_chunk * SQLP_TILE + p * 2);
            }
            if (++ld_slab == nslab) {
                ld_slab = 0;
                if (++ld_chunk == nchunks) { ld_chunk = 0; ++ld_t; }
            }
        }
        ++ld_it;
        cp_async_commit();
    };

    double acc[8][8];
    double best[NX][8];
    int bidx[NX][8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.0;
#pragma unroll
        for (int x = 0; x < NX; ++x) { best[x][r] = -INFINITY; bidx[x][r] = -1; }
    }

#pragma unroll
    for (int p = 0; p < SQLP_CT_STAGES - 1; ++p) issue_load();

    int slab = 0, chunk = 0, t = 0;
    for (long long it = 0; it < total; ++it) {
        cp_async_wait<SQLP_CT_STAGES - 2>();
        __syncthreads();
        issue_load();

        const double *st = stages + (it % SQLP_CT_STAGES) * ContractSmem<NX>::kStageDoubles;
        const double *As = st + a_off;
        const double *Bs = st + slab_doubles + b_off;
#pragma unroll
        for (int j = 0; j < SQLP_BK; ++j) {
            double2 av[4], bv[4];
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                av[n] = *reinterpret_cast<const double2 *>(As + j * SQLP_TILE + n * 16);
                bv[n] = *reinterpret_cast<const double2 *>(Bs + j * SQLP_TILE + n * 8);
            }
#pragma unroll
            for (int n = 0; n < 4; ++n) {
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    acc[2 * n][2 * m] = fma(av[n].x, bv[m].x, acc[2 * n][2 * m]);
                    acc[2 * n][2 * m + 1] = fma(av[n].x, bv[m].y, acc[2 * n][2 * m + 1]);
                    acc[2 * n + 1][2 * m] = fma(av[n].y, bv[m].x, acc[2 * n + 1][2 * m]);
                    acc[2 * n + 1][2 * m + 1] = fma(av[n].y, bv[m].y, acc[2 * n + 1][2 * m + 1]);
                }
            }
        }

        if (slab == nslab - 1) {
            // ---- chunk epilogue: bias add + running argmax (vertex index ascending) ----
            const double *bs = st + 2 * slab_doubles + b_off;
            const int kbase = chunk * SQLP_TILE + b_off;
#pragma unroll
            for (int x = 0; x < NX; ++x) {
                double bb[8];
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    double2 q = *reinterpret_cast<const double2 *>(bs + x * SQLP_TILE + m * 8);
                    bb[2 * m] = q.x;
                    bb[2 * m + 1] = q.y;
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        double v = acc[r][c] + bb[c];
                        if (v > best[x][r]) {   // strict: first maximum wins (subprob.jl:156)
                            best[x][r] = v;
                            bidx[x][r] = kbase + (c >> 1) * 8 + (c & 1);
                        }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][c] = 0.0;

            if (chunk == nchunks - 1) {
                // ---- tile epilogue: merge the 16 threads sharing each scenario row ----
#pragma unroll
                for (int x = 0; x < NX; ++x) {
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        double v = best[x][r];
                        int i = bidx[x][r];
#pragma unroll
                        for (int off = 1; off <= 2; off <<= 1) {
                            double ov = __shfl_xor_sync(0xffffffffu, v, off);
                            int oi = __shfl_xor_sync(0xffffffffu, i, off);
                            if (better(ov, oi, v, i)) { v = ov; i = oi; }
                        }
                        if (lx == 0) {
                            int row = a_off + (r >> 1) * 16 + (r & 1);
                            red_val[(x * 4 + wx) * SQLP_TILE + row] = v;
                            red_idx[(x * 4 + wx) * SQLP_TILE + row] = i;
                        }
                        best[x][r] = -INFINITY;
                        bidx[x][r] = -1;
                    }
                }
                __syncthreads();
                const long long tile = blockIdx.x + (long long)t * gridDim.x;
                for (int q = tid; q < SQLP_TILE * NX; q += SQLP_CT_THREADS) {
                    const int x = q / SQLP_TILE, row = q % SQLP_TILE;
                    double v = red_val[(x * 4) * SQLP_TILE + row];
                    int i = red_idx[(x * 4) * SQLP_TILE + row];
#pragma unroll
                    for (int w = 1; w < 4; ++w) {
                        double ov = red_val[(x * 4 + w) * SQLP_TILE + row];
                        int oi = red_idx[(x * 4 + w) * SQLP_TILE + row];
                        if (better(ov, oi, v, i)) { v = ov; i = oi; }
                    }
                    const long long sc = tile * SQLP_TILE + row;
                    if (sc < a.n_local) {
                        a.best_val[x * a.out_stride + sc] = v;
                        a.best_idx[x * a.out_stride + sc] = i;
                    }
                }
                // red_* is next written after at least one more __syncthreads (top of loop)
            }
        }
        if (++slab == nslab) {
            slab = 0;
            if (++chunk == nchunks) { chunk = 0; ++t; }
        }
    }
    cp_async_wait<0>();
}

}  // namespace sqlp
