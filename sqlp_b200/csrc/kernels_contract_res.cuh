// kernels_contract_res.cuh -- fused fp64 contraction + argmax, resident-scenario variant
// with an even ("stream-K") split of the work over the persistent grid.
//
// Same mathematics and same reference lines as kernels_contract.cuh (argmax_procedure,
// src/sd_algorithm/subprob.jl:148-166).  Role: the FIRST FALLBACK of the warp-specialised
// production kernel (kernels_contract_ws.cuh, which includes this file for the work split and
// the fix-up kernel): it needs only two ring stages next to the resident scenarios; the
// streaming kernel of kernels_contract.cuh takes over for very wide stochastic row sets.
//
// What changed against the streaming kernel, and why (profiles/r01d_contract_source_regions.txt):
//   * the per-item bookkeeping (cursor arithmetic + 5..7 bulk copies per pipeline item) took
//     14 % of every warp's time.  Here the D sub-tile of a unit (ROWS scenarios x s_pad slots,
//     61 KB at storm) is loaded ONCE per unit and stays resident while every vertex chunk
//     streams past it, so a pipeline item is a single contiguous bulk copy of the pool view
//     (+ the bias on a chunk's last slab).  L2 -> SM traffic of D drops by the number of
//     chunks (129 x at storm).
//   * work is the linear sequence of (unit, chunk) pairs; CTA b takes the contiguous span
//     [total * b / grid, total * (b + 1) / grid).  Every CTA gets the same number of chunks
//     (+-1) whatever N and K are, so there is no partial last wave (2 % at the bench shape,
//     12-17 % at ssn K=5k x N=1e5 or at a 1/8 shard of storm).  A unit cut by a span boundary
//     is finished by k_argmax_fixup: each CTA stores the (max, argmax) of its part in a
//     piece buffer and the fix-up merges the parts on (value desc, index asc), the same
//     total order the kernel uses everywhere, so the result does not depend on the split.
#pragma once
#include "kernels_contract.cuh"

namespace sqlp {

#define SQLP_RES_MAX_STAGES 8
#ifndef SQLP_RES_UNROLL   // k-groups per trip of the DMMA loop (fragment loads are pipelined inside a trip)
#define SQLP_RES_UNROLL 2
#endif


//   WR   warp rows: the CTA is WR x 4 warps, warp tile (8 MI) scenarios x 32 vertices
//   MI   m8n8 blocks per warp along the scenario axis (even)
//   KG   k-groups (of 4 row slots) per pipeline item
//   CTAS resident CTAs per SM the register budget is set for
template <int NX_, int WR_, int MI_, int KG_, int CTAS_>
struct ResidentCfg {
    static constexpr int NX = NX_, WR = WR_, MI = MI_, KG = KG_, CTAS = CTAS_;
    static constexpr int WARPS = 4 * WR, THREADS = 32 * WARPS;
    static constexpr int ROWS = 8 * MI * WR;                    // scenarios per unit
    static constexpr int UNITS_PER_TILE = SQLP_TILE / ROWS;
    static constexpr int kAGroup = ROWS * 4;                    // doubles of D per k-group
    static constexpr int kStageDoubles = KG * 512 + NX * SQLP_TILE;
    static constexpr int kRedDoubles = 4 * ROWS * NX;
    static_assert(MI % 2 == 0 && SQLP_TILE % ROWS == 0, "unit must be whole 16-column cells");
    static size_t fixed_bytes(int s_pad)
    {
        return sizeof(double) * ((size_t)(s_pad / 4) * kAGroup + kRedDoubles) + sizeof(int) * (4 * ROWS * NX) +
               sizeof(unsigned long long) * (2 * SQLP_RES_MAX_STAGES + 2);
    }
    static constexpr size_t stage_bytes() { return sizeof(double) * kStageDoubles; }
};

// span boundary b of the even split of `total` chunk-units over `grid` CTAs
__host__ __device__ __forceinline__ long long span_at(long long total, int b, int grid)
{
    return (long long)(((unsigned long long)total * (unsigned long long)b) / (unsigned long long)grid);
}

template <class C>
__global__ void __launch_bounds__(C::THREADS, C::CTAS) k_contract_resident(ContractArgs a)
{
    griddep_sync();
    if (sweep_gated_off(a)) return;
    constexpr int NX = C::NX, MI = C::MI, ROWS = C::ROWS, WARPS = C::WARPS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ngroups = a.s_pad / 4;
    const int S = a.nstages;
    double *Ares = reinterpret_cast<double *>(smem_raw);
    double *stages = Ares + (size_t)ngroups * C::kAGroup;
    double *red_val = stages + (size_t)S * C::kStageDoubles;
    int *red_idx = reinterpret_cast<int *>(red_val + C::kRedDoubles);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(red_idx + 4 * ROWS * NX);
    const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + SQLP_RES_MAX_STAGES),
                   afull = smem_u32(bars + 2 * SQLP_RES_MAX_STAGES);

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wx = warp & 3, wy = warp >> 2;    // 4 warps along the vertex axis, WR along scenarios
    const int ly = lane >> 2, lx = lane & 3;    // m8n8k4 C fragment: row ly, columns 2 lx + {0, 1}
    const int a_off = wy * (MI / 2) * 64 + lane * 2;
    const int b_off = (wx * 2) * 64 + lane * 2;

    const long long K = *a.d_K;
    const int nchunks = (int)((K + SQLP_TILE - 1) / SQLP_TILE);
    const int nslab = (ngroups + C::KG - 1) / C::KG;
    const long long nunits = (long long)C::UNITS_PER_TILE * a.ntiles;

    if (nchunks == 0) {  // empty pool: nothing beats -Inf (subprob.jl:151)
        for (long long unit = blockIdx.x; unit < nunits; unit += gridDim.x)
            for (int q = tid; q < ROWS * NX; q += C::THREADS) {
                const long long i = unit * ROWS + (q % ROWS);
                if (i < a.n_local) {
                    a.best_val[(q / ROWS) * a.out_stride + i] = -INFINITY;
                    a.best_idx[(q / ROWS) * a.out_stride + i] = -1;
                }
            }
        return;
    }

    const long long total_cu = nunits * nchunks;
    const long long L0 = span_at(total_cu, blockIdx.x, gridDim.x);
    const long long L1 = span_at(total_cu, blockIdx.x + 1, gridDim.x);
    if (L0 >= L1) return;
    const size_t tile_doubles = (size_t)a.s_pad * SQLP_TILE;

    if (tid == 0) {
        for (int q = 0; q < S; ++q) {
            mbar_init(full0 + 8 * q, 1);        // one arrive.expect_tx per fill
            mbar_init(empty0 + 8 * q, WARPS);   // one arrive per consumer warp
        }
        mbar_init(afull, 1);
        mbar_fence_init();
    }
    __syncthreads();

    // ---- pool-view producer: items are (chunk, slab) in span order; the address depends on
    // the chunk and slab only, so the cursor ignores unit boundaries except for the wrap.
    int p_left = (int)(L1 - L0) * nslab;   // items of this span (< 2^31 for any supported shape)
    int p_chunk = (int)(L0 % nchunks), p_slab = 0, p_stage = 0, p_duty = 0;
    unsigned p_par = 0;
    bool p_wrapped = false;
    auto produce = [&]() {
        if (p_left == 0) return;
        if (warp == p_duty && lane == 0) {
            if (p_wrapped) mbar_wait(empty0 + 8 * p_stage, p_par ^ 1u);   // previous contents released
            double *st = stages + (size_t)p_stage * C::kStageDoubles;
            const unsigned bar = full0 + 8 * p_stage;
            const bool last = (p_slab == nslab - 1);
            const int pg = min(C::KG, ngroups - p_slab * C::KG);
            mbar_arrive_expect_tx(bar, (pg * 512 + (last ? NX * SQLP_TILE : 0)) * 8);
            bulk_g2s(smem_u32(st), a.PiS + (size_t)p_chunk * tile_doubles + (size_t)p_slab * (C::KG * 512),
                     pg * 512 * 8, bar);
            if (last) {
#pragma unroll
                for (int x = 0; x < NX; ++x)
                    bulk_g2s(smem_u32(st + C::KG * 512 + x * SQLP_TILE),
                             a.bias + x * a.bias_stride + (size_t)p_chunk * SQLP_TILE, SQLP_TILE * 8, bar);
            }
        }
        if (++p_slab == nslab) {
            p_slab = 0;
            if (++p_chunk == nchunks) p_chunk = 0;
        }
        if (++p_stage == S) { p_stage = 0; p_par ^= 1u; p_wrapped = true; }
        if (++p_duty == WARPS) p_duty = 0;
        --p_left;
    };

    double acc[MI][4][2];   // [mi][ni][h]: scenario (wy MI + mi) 8 + ly of the unit, vertex wx 32 + ni 8 + 2 lx + h
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
    for (int q = 0; q < a.prefetch; ++q) produce();

    auto mma_group = [&](const double *Ag, const double *Bg) {
        double2 av[MI / 2], bv[2];
#pragma unroll
        for (int q = 0; q < MI / 2; ++q) av[q] = *reinterpret_cast<const double2 *>(Ag + q * 64);
#pragma unroll
        for (int q = 0; q < 2; ++q) bv[q] = *reinterpret_cast<const double2 *>(Bg + q * 64);
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) {
            const double af = (mi & 1) ? av[mi >> 1].y : av[mi >> 1].x;
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const double bf = (ni & 1) ? bv[ni >> 1].y : bv[ni >> 1].x;
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(acc[mi][ni][0]), "+d"(acc[mi][ni][1])
                             : "d"(af), "d"(bf));
            }
        }
    };

    // Bias add + running argmax of one 8-scenario row block (vertex index ascending in (ni, h)).
    auto epilogue_row = [&](int r, const double *bs, int kbase) {
#pragma unroll
        for (int x = 0; x < NX; ++x) {
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const double2 bb = *reinterpret_cast<const double2 *>(bs + x * SQLP_TILE + ni * 8);
                double v0 = acc[r][ni][0] + bb.x;
                if (v0 > best[x][r]) {   // strict: first maximum wins (subprob.jl:156)
                    best[x][r] = v0;
                    bidx[x][r] = kbase + ni * 8;
                }
                double v1 = acc[r][ni][1] + bb.y;
                if (v1 > best[x][r]) {
                    best[x][r] = v1;
                    bidx[x][r] = kbase + ni * 8 + 1;
                }
            }
        }
    };
    // One 2..4 KB bulk copy per k-group brings a unit's scenarios into shared memory.
    auto load_unit = [&](long long u) {
        if (warp == 0) {
            if (lane == 0) mbar_arrive_expect_tx(afull, (unsigned)(ngroups * C::kAGroup * 8));
            __syncwarp();
            const double *src = a.D + (size_t)(u / C::UNITS_PER_TILE) * tile_doubles +
                                (size_t)(u % C::UNITS_PER_TILE) * C::kAGroup;
            for (int g = lane; g < ngroups; g += 32)
                bulk_g2s(smem_u32(Ares + (size_t)g * C::kAGroup), src + (size_t)g * 512, C::kAGroup * 8, afull);
        }
    };

    int stage = 0;
    unsigned par = 0, uphase = 0;
    bool ready = false;   // the current item is already known to have landed
    long long unit = L0 / nchunks;
    int c_begin = (int)(L0 - unit * nchunks);
    int rem = (int)(L1 - L0);   // chunk-units left in the span
    load_unit(unit);
#pragma unroll 1
    while (rem > 0) {
        const int c_end = min(nchunks, c_begin + rem);

        // ---- the unit's scenarios become resident (the copies were issued by load_unit below:
        // before the loop for the first unit, during the previous unit's epilogue otherwise)
        mbar_wait(afull, uphase);
        uphase ^= 1u;

#pragma unroll 1
        for (int chunk = c_begin; chunk < c_end; ++chunk) {
#pragma unroll 1
            for (int slab = 0; slab < nslab; ++slab) {
                produce();
                if (!ready) mbar_wait(full0 + 8 * stage, par);
                // probe the next item now; the answer comes back while this item's DMMAs run
                const int nstage = (stage + 1 == S) ? 0 : stage + 1;
                const unsigned npar = (stage + 1 == S) ? par ^ 1u : par;
                ready = mbar_test(full0 + 8 * nstage, npar);

                const double *st = stages + (size_t)stage * C::kStageDoubles;
                const double *As = Ares + (size_t)(slab * C::KG) * C::kAGroup + a_off;
                const double *Bs = st + b_off;
                const int ng = min(C::KG, ngroups - slab * C::KG);
                int g = 0;
#pragma unroll 1
                for (; g + SQLP_RES_UNROLL <= ng; g += SQLP_RES_UNROLL) {
#pragma unroll
                    for (int u = 0; u < SQLP_RES_UNROLL; ++u)
                        mma_group(As + (g + u) * C::kAGroup, Bs + (g + u) * 512);
                }
#pragma unroll 1
                for (; g < ng; ++g) mma_group(As + g * C::kAGroup, Bs + g * 512);

                if (slab == nslab - 1) {
                    // ---- chunk epilogue: bias add + running argmax (vertex index ascending) ----
                    const double *bs = st + C::KG * 512 + wx * 32 + lx * 2;
                    const int kbase = chunk * SQLP_TILE + wx * 32 + lx * 2;
#pragma unroll
                    for (int r = 0; r < MI; ++r) epilogue_row(r, bs, kbase);
#pragma unroll
                    for (int r = 0; r < MI; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;
                }
                // this warp is done with the stage (operands and, on a chunk's last slab, the bias)
                __syncwarp();
                if (lane == 0) mbar_arrive(empty0 + 8 * stage);
                if (++stage == S) { stage = 0; par ^= 1u; }
            }
        }

        // ---- unit epilogue: merge the 16 threads (4 lanes x 4 warps) sharing each scenario row
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
                    const int row = (wy * MI + r) * 8 + ly;
                    red_val[(x * 4 + wx) * ROWS + row] = v;
                    red_idx[(x * 4 + wx) * ROWS + row] = i;
                }
                best[x][r] = -INFINITY;
                bidx[x][r] = -1;
            }
        }
        __syncthreads();
        // every warp is past its last read of the resident scenarios: the next unit's copies run
        // under the merge below
        if (rem > c_end - c_begin) load_unit(unit + 1);
        const bool complete = (c_begin == 0 && c_end == nchunks);
        for (int q = tid; q < ROWS * NX; q += C::THREADS) {
            const int x = q / ROWS, row = q % ROWS;
            double v = red_val[(x * 4) * ROWS + row];
            int i = red_idx[(x * 4) * ROWS + row];
#pragma unroll
            for (int w = 1; w < 4; ++w) {
                double ov = red_val[(x * 4 + w) * ROWS + row];
                int oi = red_idx[(x * 4 + w) * ROWS + row];
                if (better(ov, oi, v, i)) { v = ov; i = oi; }
            }
            if (complete) {
                const long long sc = unit * ROWS + row;
                if (sc < a.n_local) {
                    a.best_val[x * a.out_stride + sc] = v;
                    a.best_idx[x * a.out_stride + sc] = i;
                }
            } else {   // a part of the unit: slot 0 = starts after chunk 0, slot 1 = starts at chunk 0
                const size_t o = (((size_t)blockIdx.x * 2 + (c_begin > 0 ? 0 : 1)) * NX + x) * ROWS + row;
                a.piece_val[o] = v;
                a.piece_idx[o] = i;
            }
        }
        __syncthreads();   // red_* and the resident scenarios are rewritten for the next unit
        rem -= c_end - c_begin;
        c_begin = 0;
        ++unit;
    }
}

// Units cut by a span boundary: merge the parts.  Block b finishes the unit in which span b
// ends, provided span b also holds that unit's first chunk (so exactly one block per unit).
template <int ROWS, int NX>
__global__ void k_argmax_fixup(ContractArgs a, int main_grid)
{
    griddep_sync();
    if (sweep_gated_off(a)) return;
    const long long K = *a.d_K;
    const int nchunks = (int)((K + SQLP_TILE - 1) / SQLP_TILE);
    if (nchunks == 0) return;
    const long long nunits = (long long)(SQLP_TILE / ROWS) * a.ntiles;
    const long long total_cu = nunits * nchunks;
    const int b = blockIdx.x;
    const long long L0 = span_at(total_cu, b, main_grid), L1 = span_at(total_cu, b + 1, main_grid);
    if (L0 >= L1 || L1 >= total_cu || L1 % nchunks == 0) return;
    const long long unit = L1 / nchunks;
    const long long ustart = unit * nchunks, uend = ustart + nchunks;
    if (L0 > ustart) return;
    for (int q = threadIdx.x; q < ROWS * NX; q += blockDim.x) {
        const int x = q / ROWS, row = q % ROWS;
        size_t o = (((size_t)b * 2 + 1) * NX + x) * ROWS + row;
        double v = a.piece_val[o];
        int i = a.piece_idx[o];
        // the blocks holding the rest of the unit: b + 1 .. last, found by span arithmetic alone, so their
        // pieces are fetched four at a time instead of one dependent load after another
        int last = b + 1;
        while (last < main_grid - 1 && span_at(total_cu, last + 1, main_grid) < uend) ++last;
        for (int b0 = b + 1; b0 <= last; b0 += 4) {
            double ov[4];
            int oi[4];
            bool use[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int bb = b0 + u;
                use[u] = bb <= last && span_at(total_cu, bb, main_grid) < span_at(total_cu, bb + 1, main_grid);
                o = (((size_t)min(bb, last) * 2) * NX + x) * ROWS + row;
                ov[u] = a.piece_val[o];
                oi[u] = a.piece_idx[o];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (use[u] && better(ov[u], oi[u], v, i)) { v = ov[u]; i = oi[u]; }
        }
        const long long sc = unit * ROWS + row;
        if (sc < a.n_local) {
            a.best_val[x * a.out_stride + sc] = v;
            a.best_idx[x * a.out_stride + sc] = i;
        }
    }
}

}  // namespace sqlp
