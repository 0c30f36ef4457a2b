// kernels_contract.cuh -- the fused fp64 contraction + argmax, STREAMING variant: both operands
// flow through the shared-memory ring, so it works for any number of stochastic rows.  It was the
// hot kernel up to v6 and is now the last fallback behind
// kernels_contract_ws.cuh and kernels_contract_res.cuh; ContractArgs and the (value, index)
// order `better` defined here are shared by all three.
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
//   fp64 has no tcgen05 path (the 5th-gen tensor cores stop at tf32), so the math runs on
//   the FP64 tensor pipe through mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4).  Measured on this
//   pool's B200 (profiles/fp64_peak.json): DMMA 37.0 TFLOP/s, DFMA chain 36.7, but an 8x8
//   register-tiled DFMA loop tops out near 31 (three 64-bit sources per DFMA exceed the
//   register-file read bandwidth), and the first version of this kernel, built on DFMA,
//   reached 21.4.  DMMA reads 4 register pairs per 256 MACs instead of 3 per MAC.
//   CTA tile 64 scenarios x 128 vertices, 128 threads = 4 warps side by side along the
//   vertex axis, warp tile 64 x 32 = 8 x 4 m8n8 accumulator blocks (128 registers); two
//   CTAs per SM, i.e. two warps from DIFFERENT CTAs on every sub-partition.  Both operands live in HBM in the
//   fragment-major tile layout of common.cuh, so (a) every 8-slot pipeline slab is one
//   contiguous 8 KB block, moved L2 -> shared memory by ONE bulk async copy (TMA 1-D,
//   cp.async.bulk, SASS UBLKCP) per operand, and (b) one conflict-free LDS.128 hands each
//   lane two operand fragments: 6 LDS.128 per 32 DMMA.
//   Pipeline: SQLP_CT_STAGES stages guarded by full/empty mbarriers, running continuously
//   across slabs, vertex chunks and scenario half-tiles; no CTA-wide barrier in the steady
//   state.  ptxas spaces a warp's DMMAs 16 cycles apart, exactly the pipe's rate, so ONE
//   warp can saturate its sub-partition's FP64 tensor pipe; the second warp there belongs to
//   the other resident CTA, whose phase is unrelated, so it feeds the pipe while the first is
//   in its epilogue or waiting on a load (two warps of one CTA share the pipe fairly and
//   would reach their epilogues together).  The copy for item L is issued by lane 0 of warp
//   L % 4 once every warp has released the stage's previous contents.
//   After the last slab of a vertex chunk the epilogue adds the bias and folds the 64
//   scores of each thread into per-thread running (max, argmax); after the last chunk of a
//   scenario tile the 16 threads that share a scenario row merge with warp shuffles and
//   one shared-memory pass, comparing on (value desc, index asc) so the first index wins.
//   Roofline: 2 * s flop per (scenario, vertex) evaluation against the measured FP64 peak.
#pragma once
#include "common.cuh"

namespace sqlp {

#define SQLP_CT_THREADS 128

// Compile-time shape of one contraction variant.
//   MI      m8n8 blocks per warp along the scenario axis: the CTA tile is (8 MI) scenarios x
//           128 vertices (4 warps side by side along the vertex axis, warp tile 8 MI x 32)
//   STAGES  pipeline stages; PREFETCH  how many items ahead the copies run.  The fastest warp
//           may lead the slowest by STAGES - PREFETCH - 1 items before a copy has to wait.
//   CTAS    resident CTAs per SM the register budget is set for
//   KG      k-groups (of 4 row slots) per pipeline item; the last item of a chunk may be
//           shorter (s_pad is a multiple of 8, so every item holds an even number of groups)
template <int NX_, int MI_, int STAGES_, int PREFETCH_, int CTAS_, int KG_>
struct ContractCfg {
    static constexpr int NX = NX_, MI = MI_, STAGES = STAGES_, PREFETCH = PREFETCH_, CTAS = CTAS_, KG = KG_;
    static constexpr int ROWS = 8 * MI;                         // scenarios per work unit
    static constexpr int UNITS_PER_TILE = SQLP_TILE / ROWS;
    static constexpr int kAGroup = MI * 32;                     // doubles of A per k-group
    static constexpr int kADoubles = KG * kAGroup;
    static constexpr int kBDoubles = KG * 512;
    static constexpr int kStageDoubles = kADoubles + kBDoubles + NX * SQLP_TILE;
    static constexpr int kRedDoubles = 4 * ROWS * NX;
    static constexpr size_t smem_bytes()
    {
        return sizeof(double) * (STAGES * kStageDoubles + kRedDoubles) + sizeof(int) * (4 * ROWS * NX) +
               sizeof(unsigned long long) * 2 * STAGES;
    }
};

struct ContractArgs {
    const double *D;        // [ntiles] fragment-major tiles: scenario deltas on the stochastic rows
    const double *PiS;      // [nchunks] fragment-major tiles: pool restricted to those rows
    const double *bias;     // [NX][bias_stride]; -inf for k >= K
    long long bias_stride;
    const long long *d_K;   // pool size (device resident)
    int s_pad;              // multiple of SQLP_BK
    int ntiles;
    long long n_local;      // scenarios held by this rank
    double *best_val;       // [NX][out_stride]
    int *best_idx;          // [NX][out_stride]
    long long out_stride;
    // resident variant only (kernels_contract_res.cuh)
    int nstages, prefetch;  // pool-ring depth and how many items ahead the copies run
    int lag_ns;             // warp-specialised variant: start-up delay of the second row group
    double *piece_val;      // [grid][2][NX][ROWS] partial results of units cut by a span boundary
    int *piece_idx;
    const ScreenCtl *gate;  // non-null: a screening pass ran first; the sweep runs only if that fell back
};

__device__ __forceinline__ bool sweep_gated_off(const ContractArgs &a)
{
    return a.gate != nullptr && !screen_falls_back(a.gate);
}

__device__ __forceinline__ bool better(double ov, int oi, double v, int i)
{
    return (ov > v) || (ov == v && (unsigned)oi < (unsigned)i);
}

template <class C>
__global__ void __launch_bounds__(SQLP_CT_THREADS, C::CTAS) k_contract_argmax(ContractArgs a)
{
    griddep_sync();
    if (sweep_gated_off(a)) return;
    constexpr int NX = C::NX, MI = C::MI, S = C::STAGES, ROWS = C::ROWS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *stages = reinterpret_cast<double *>(smem_raw);
    double *red_val = stages + S * C::kStageDoubles;
    int *red_idx = reinterpret_cast<int *>(red_val + C::kRedDoubles);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(red_idx + 4 * ROWS * NX);
    const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + S);

    const int tid = threadIdx.x;
    const int lane = tid & 31, wx = tid >> 5;  // 4 warps along the vertex axis
    const int ly = lane >> 2, lx = lane & 3;   // m8n8k4 C fragment: row ly, columns 2 lx + {0, 1}
    // fragment-major cell [g][P][lane][h]: the staged A sub-tile holds MI / 2 cells per k-group
    // (scenario blocks mi = 0..MI-1), the B cells of this warp are P = 2 wx, 2 wx + 1 (ni = 0..3)
    const int a_off = lane * 2;
    const int b_off = (wx * 2) * 64 + lane * 2;

    const long long K = *a.d_K;
    const int nchunks = (int)((K + SQLP_TILE - 1) / SQLP_TILE);
    const int ngroups = a.s_pad / 4;
    const int nslab = (ngroups + C::KG - 1) / C::KG;
    const long long nunits = (long long)C::UNITS_PER_TILE * a.ntiles;
    const long long my_units = (nunits > blockIdx.x) ? (nunits - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (nchunks == 0) {  // empty pool: nothing beats -Inf (subprob.jl:151)
        for (long long t = 0; t < my_units; ++t) {
            const long long unit = blockIdx.x + t * gridDim.x;
            for (int q = tid; q < ROWS * NX; q += SQLP_CT_THREADS) {
                const long long i = unit * ROWS + (q % ROWS);
                if (i < a.n_local) {
                    a.best_val[(q / ROWS) * a.out_stride + i] = -INFINITY;
                    a.best_idx[(q / ROWS) * a.out_stride + i] = -1;
                }
            }
        }
        return;
    }

    const long long total = my_units * nchunks * nslab;
    const size_t tile_doubles = (size_t)a.s_pad * SQLP_TILE;

    if (tid == 0) {
        for (int q = 0; q < S; ++q) {
            mbar_init(full0 + 8 * q, 1);                       // one arrive.expect_tx per fill
            mbar_init(empty0 + 8 * q, SQLP_CT_THREADS / 32);   // one arrive per consumer warp
        }
        mbar_fence_init();
    }
    __syncthreads();

    // ---- producer cursor: item pL lands in stage p_stage; every thread advances it the same
    // way (no divisions in the loop), lane 0 of warp pL % 4 issues the copies.
    long long pL = 0, p_unit = blockIdx.x;
    int p_stage = 0, p_slab = 0, p_chunk = 0, p_duty = 0;
    unsigned p_par = 0;   // parity of pL / S
    auto produce = [&]() {
        if (pL >= total) return;
        if (wx == p_duty && lane == 0) {
            if (pL >= S) mbar_wait(empty0 + 8 * p_stage, p_par ^ 1u);   // previous contents released
            const long long tile = p_unit / C::UNITS_PER_TILE;
            const int sub = (int)(p_unit % C::UNITS_PER_TILE);
            double *st = stages + p_stage * C::kStageDoubles;
            const unsigned bar = full0 + 8 * p_stage;
            const bool last = (p_slab == nslab - 1);
            const int pg = min(C::KG, ngroups - p_slab * C::KG);   // k-groups in this item
            mbar_arrive_expect_tx(bar, (pg * (C::kAGroup + 512) + (last ? NX * SQLP_TILE : 0)) * 8);
            const double *gA = a.D + tile * tile_doubles + (size_t)p_slab * (C::KG * 512) + sub * C::kAGroup;
            for (int g = 0; g < pg; ++g)   // per k-group, this unit's cells of the D tile
                bulk_g2s(smem_u32(st + g * C::kAGroup), gA + g * 512, C::kAGroup * 8, bar);
            bulk_g2s(smem_u32(st + C::kADoubles),
                     a.PiS + (size_t)p_chunk * tile_doubles + (size_t)p_slab * (C::KG * 512), pg * 512 * 8, bar);
            if (last) {
#pragma unroll
                for (int x = 0; x < NX; ++x)
                    bulk_g2s(smem_u32(st + C::kADoubles + C::kBDoubles + x * SQLP_TILE),
                             a.bias + x * a.bias_stride + (size_t)p_chunk * SQLP_TILE, SQLP_TILE * 8, bar);
            }
        }
        if (++p_slab == nslab) {
            p_slab = 0;
            if (++p_chunk == nchunks) { p_chunk = 0; p_unit += gridDim.x; }
        }
        if (++p_stage == S) { p_stage = 0; p_par ^= 1u; }
        p_duty = (p_duty + 1) & 3;
        ++pL;
    };

    double acc[MI][4][2];   // [mi][ni][h]: scenario mi*8 + ly of the unit, vertex wx*32 + ni*8 + 2 lx + h
    double best[NX][MI];
    int bidx[NX][MI];
#pragma unroll
    for (int r = 0; r < MI; ++r) {
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;
#pragma unroll
        for (int x = 0; x < NX; ++x) { best[x][r] = -INFINITY; bidx[x][r] = -1; }
    }

#pragma unroll 1
    for (int q = 0; q < C::PREFETCH; ++q) produce();

    int slab = 0, chunk = 0, stage = 0;
    unsigned par = 0;
    long long unit = blockIdx.x;
    bool ready = false;   // item `it` already known to have landed (probed during item it - 1)
#pragma unroll 1
    for (long long it = 0; it < total; ++it) {
        produce();
        if (!ready) mbar_wait(full0 + 8 * stage, par);
        // probe the next item now; the answer comes back while this item's DMMAs run
        const int nstage = (stage + 1 == S) ? 0 : stage + 1;
        const unsigned npar = (stage + 1 == S) ? par ^ 1u : par;
        ready = mbar_test(full0 + 8 * nstage, npar);

        const double *st = stages + stage * C::kStageDoubles;
        const double *As = st + a_off;
        const double *Bs = st + C::kADoubles + b_off;
        const int ng = min(C::KG, ngroups - slab * C::KG);
#pragma unroll 1
        for (int g0 = 0; g0 < ng; g0 += 2) {
#pragma unroll
        for (int gg = 0; gg < 2; ++gg) {
            const int g = g0 + gg;
            double2 av[MI / 2], bv[2];
#pragma unroll
            for (int q = 0; q < MI / 2; ++q)
                av[q] = *reinterpret_cast<const double2 *>(As + g * C::kAGroup + q * 64);
#pragma unroll
            for (int q = 0; q < 2; ++q)
                bv[q] = *reinterpret_cast<const double2 *>(Bs + g * 512 + q * 64);
#pragma unroll
            for (int mi = 0; mi < MI; ++mi) {
                const double af = (mi & 1) ? av[mi >> 1].y : av[mi >> 1].x;
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    const double bf = (ni & 1) ? bv[ni >> 1].y : bv[ni >> 1].x;
                    asm volatile(
                        "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                        : "+d"(acc[mi][ni][0]), "+d"(acc[mi][ni][1])
                        : "d"(af), "d"(bf));
                }
            }
        }
        }

        if (slab == nslab - 1) {
            // ---- chunk epilogue: bias add + running argmax (vertex index ascending) ----
            const double *bs = st + C::kADoubles + C::kBDoubles + wx * 32 + lx * 2;
            const int kbase = chunk * SQLP_TILE + wx * 32 + lx * 2;
#ifndef SQLP_EXPERIMENT_NO_EPILOGUE
#pragma unroll
            for (int x = 0; x < NX; ++x) {
                double2 bb[4];
#pragma unroll
                for (int ni = 0; ni < 4; ++ni)
                    bb[ni] = *reinterpret_cast<const double2 *>(bs + x * SQLP_TILE + ni * 8);
#pragma unroll
                for (int r = 0; r < MI; ++r) {
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) {   // vertex index ascending in (ni, h)
                        double v0 = acc[r][ni][0] + bb[ni].x;
                        if (v0 > best[x][r]) {   // strict: first maximum wins (subprob.jl:156)
                            best[x][r] = v0;
                            bidx[x][r] = kbase + ni * 8;
                        }
                        double v1 = acc[r][ni][1] + bb[ni].y;
                        if (v1 > best[x][r]) {
                            best[x][r] = v1;
                            bidx[x][r] = kbase + ni * 8 + 1;
                        }
                    }
                }
            }
#else
            if (chunk == nchunks - 1) {   // experiment: epilogue cost ceiling (results are wrong)
#pragma unroll
                for (int r = 0; r < MI; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
#pragma unroll
                        for (int x = 0; x < NX; ++x)
                            if (acc[r][c][0] + acc[r][c][1] > best[x][r]) { best[x][r] = acc[r][c][0]; bidx[x][r] = kbase + c; }
                    }
            }
#endif
#pragma unroll
            for (int r = 0; r < MI; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;

            if (chunk == nchunks - 1) {
                // ---- unit epilogue: merge the 16 threads sharing each scenario row ----
#pragma unroll
                for (int x = 0; x < NX; ++x) {
#pragma unroll
                    for (int r = 0; r < MI; ++r) {
                        double v = best[x][r];
                        int i = bidx[x][r];
#pragma unroll
                        for (int off = 1; off <= 2; off <<= 1) {
                            double ov = __shfl_xor_sync(0xffffffffu, v, off);
                            int oi = __shfl_xor_sync(0xffffffffu, i, off);
                            if (better(ov, oi, v, i)) { v = ov; i = oi; }
                        }
                        if (lx == 0) {
                            const int row = r * 8 + ly;
                            red_val[(x * 4 + wx) * ROWS + row] = v;
                            red_idx[(x * 4 + wx) * ROWS + row] = i;
                        }
                        best[x][r] = -INFINITY;
                        bidx[x][r] = -1;
                    }
                }
                __syncthreads();
                for (int q = tid; q < ROWS * NX; q += SQLP_CT_THREADS) {
                    const int x = q / ROWS, row = q % ROWS;
                    double v = red_val[(x * 4) * ROWS + row];
                    int i = red_idx[(x * 4) * ROWS + row];
#pragma unroll
                    for (int w = 1; w < 4; ++w) {
                        double ov = red_val[(x * 4 + w) * ROWS + row];
                        int oi = red_idx[(x * 4 + w) * ROWS + row];
                        if (better(ov, oi, v, i)) { v = ov; i = oi; }
                    }
                    const long long sc = unit * ROWS + row;
                    if (sc < a.n_local) {
                        a.best_val[x * a.out_stride + sc] = v;
                        a.best_idx[x * a.out_stride + sc] = i;
                    }
                }
                __syncthreads();   // red_* may be rewritten at the next unit's end
            }
        }
        // this warp is done with the stage (operands and, on a chunk's last slab, the bias)
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * stage);
        if (++stage == S) { stage = 0; par ^= 1u; }
        if (++slab == nslab) {
            slab = 0;
            if (++chunk == nchunks) { chunk = 0; unit += gridDim.x; }
        }
    }
}

}  // namespace sqlp
