// kernels_contract_ws.cuh -- fused fp64 contraction + argmax, warp-specialised variant.
//
// Same mathematics, operand layouts, even (unit, chunk) split and fix-up as the resident kernel of
// kernels_contract_res.cuh (reference: argmax_procedure, src/sd_algorithm/subprob.jl:148-166).
// What differs is who does what inside the CTA (one per SM, 384 threads = three warpgroups):
//
//   warpgroup 0, 1   consumer ROW GROUPS: 4 warps each, 64 scenarios x 128 vertices per group.  A
//                    group owns its half of the resident scenarios (own full barrier, own named
//                    barrier for the unit epilogue), so the two groups never wait for each other:
//                    they only share the pool ring.  Group 1 starts a fraction of a chunk later and
//                    keeps that lag, so the groups' chunk epilogues and per-item bookkeeping -- the
//                    phases in which a warp issues no DMMA -- fall into the other group's DMMA
//                    stream instead of coinciding on every sub-partition (the resident kernel's
//                    eight warps reach each chunk boundary together: 7 % of the pipe idles there).
//   warpgroup 2      one PRODUCER warp feeds the pool ring (wait for the stage to be released by
//                    all eight consumer warps, expect_tx, one bulk copy per item + the bias on a
//                    chunk's last item); its three other warps retire at once.  The consumers carry
//                    no producer cursor and do no copy bookkeeping.
//   registers        launched at 168 per thread; the producer warpgroup shrinks to 24 and the
//                    consumers grow to 240 (setmaxnreg), the split CUTLASS uses for two MMA
//                    warpgroups.
#pragma once
#include "kernels_contract_res.cuh"

namespace sqlp {

#define SQLP_WS_THREADS 384
#define SQLP_WS_CONSUMER_REGS 240
#define SQLP_WS_PRODUCER_REGS 24

template <int NX_, int KG_>
struct WsCfg {
    static constexpr int NX = NX_, KG = KG_, MI = 8;
    static constexpr int ROWS = 128, GROUP_ROWS = 64;           // scenarios per unit / per row group
    static constexpr int UNITS_PER_TILE = SQLP_TILE / ROWS;
    static constexpr int kAGroup = ROWS * 4;                    // doubles of D per k-group
    static constexpr int kStageDoubles = KG * 512 + NX * SQLP_TILE;
    static constexpr int kRedDoubles = 4 * ROWS * NX;
    static size_t fixed_bytes(int s_pad)
    {
        return sizeof(double) * ((size_t)(s_pad / 4) * kAGroup + kRedDoubles) + sizeof(int) * (4 * ROWS * NX) +
               sizeof(unsigned long long) * (2 * SQLP_RES_MAX_STAGES + 4);
    }
    static constexpr size_t stage_bytes() { return sizeof(double) * kStageDoubles; }
};

__device__ __forceinline__ void group_barrier(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}

template <class C>
__global__ void __launch_bounds__(SQLP_WS_THREADS, 1) k_contract_ws(ContractArgs a)
{
    griddep_sync();
    if (sweep_gated_off(a)) return;
    constexpr int NX = C::NX, MI = C::MI, ROWS = C::ROWS, GR = C::GROUP_ROWS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int ngroups = a.s_pad / 4;
    const int S = a.nstages;
    double *Ares = reinterpret_cast<double *>(smem_raw);
    double *stages = Ares + (size_t)ngroups * C::kAGroup;
    double *red_val = stages + (size_t)S * C::kStageDoubles;
    int *red_idx = reinterpret_cast<int *>(red_val + C::kRedDoubles);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(red_idx + 4 * ROWS * NX);
    const unsigned full0 = smem_u32(bars), empty0 = smem_u32(bars + SQLP_RES_MAX_STAGES),
                   afull0 = smem_u32(bars + 2 * SQLP_RES_MAX_STAGES);   // one per row group

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wg = warp >> 2;                   // 0, 1: consumer row groups; 2: producer
    const int wx = warp & 3;                    // vertex quarter of a consumer warp
    const int ly = lane >> 2, lx = lane & 3;    // m8n8k4 C fragment: row ly, columns 2 lx + {0, 1}

    const long long K = *a.d_K;
    const int nchunks = (int)((K + SQLP_TILE - 1) / SQLP_TILE);
    const int nslab = (ngroups + C::KG - 1) / C::KG;
    const long long nunits = (long long)C::UNITS_PER_TILE * a.ntiles;

    if (nchunks == 0) {  // empty pool: nothing beats -Inf (subprob.jl:151)
        for (long long unit = blockIdx.x; unit < nunits; unit += gridDim.x)
            for (int q = tid; q < ROWS * NX; q += SQLP_WS_THREADS) {
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
            mbar_init(full0 + 8 * q, 1);    // one arrive.expect_tx per fill
            mbar_init(empty0 + 8 * q, 8);   // one arrive per consumer warp
        }
        mbar_init(afull0, 1);
        mbar_init(afull0 + 8, 1);
        mbar_fence_init();
    }
    __syncthreads();

    if (wg == 2) {
        // ------------------------------------------------------------ producer warpgroup ----
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;\n" ::"n"(SQLP_WS_PRODUCER_REGS));
        if (warp != 8 || lane != 0) return;
        int left = (int)(L1 - L0) * nslab;   // items of this span, in (chunk, slab) order
        int chunk = (int)(L0 % nchunks), slab = 0, stage = 0;
        unsigned par = 0;
        bool wrapped = false;
        for (; left > 0; --left) {
            if (wrapped) mbar_wait(empty0 + 8 * stage, par ^ 1u);   // previous contents released
            double *st = stages + (size_t)stage * C::kStageDoubles;
            const unsigned bar = full0 + 8 * stage;
            const bool last = (slab == nslab - 1);
            const int pg = min(C::KG, ngroups - slab * C::KG);
            mbar_arrive_expect_tx(bar, (pg * 512 + (last ? NX * SQLP_TILE : 0)) * 8);
            bulk_g2s(smem_u32(st), a.PiS + (size_t)chunk * tile_doubles + (size_t)slab * (C::KG * 512),
                     pg * 512 * 8, bar);
            if (last) {
#pragma unroll
                for (int x = 0; x < NX; ++x)
                    bulk_g2s(smem_u32(st + C::KG * 512 + x * SQLP_TILE),
                             a.bias + x * a.bias_stride + (size_t)chunk * SQLP_TILE, SQLP_TILE * 8, bar);
            }
            if (++slab == nslab) {
                slab = 0;
                if (++chunk == nchunks) chunk = 0;
            }
            if (++stage == S) { stage = 0; par ^= 1u; wrapped = true; }
        }
        return;
    }

    // ---------------------------------------------------------------- consumer row groups ----
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;\n" ::"n"(SQLP_WS_CONSUMER_REGS));
    const int gtid = tid & 127;                       // thread within the row group
    const unsigned afull = afull0 + 8 * wg;
    const int a_off = wg * (MI / 2) * 64 + lane * 2;  // the group's four 16-row cells of every k-group
    const int b_off = (wx * 2) * 64 + lane * 2;
    double *gred_val = red_val + wg * (4 * GR * NX);   // [x][wx][GR]
    int *gred_idx = red_idx + wg * (4 * GR * NX);

    double acc[MI][4][2];   // [mi][ni][h]: scenario wg 64 + mi 8 + ly of the unit, vertex wx 32 + ni 8 + 2 lx + h
    double best[NX][MI];
    int bidx[NX][MI];
#pragma unroll
    for (int r = 0; r < MI; ++r) {
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c][0] = acc[r][c][1] = 0.0;
#pragma unroll
        for (int x = 0; x < NX; ++x) { best[x][r] = -INFINITY; bidx[x][r] = -1; }
    }

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
    // The group's half of a unit's scenarios: one 2 KB bulk copy per k-group, issued by its warp 0.
    auto load_unit = [&](long long u) {
        if (wx == 0) {
            if (lane == 0) mbar_arrive_expect_tx(afull, (unsigned)(ngroups * (C::kAGroup / 2) * 8));
            __syncwarp();
            const double *src = a.D + (size_t)(u / C::UNITS_PER_TILE) * tile_doubles +
                                (size_t)(u % C::UNITS_PER_TILE) * C::kAGroup + wg * (C::kAGroup / 2);
            for (int g = lane; g < ngroups; g += 32)
                bulk_g2s(smem_u32(Ares + (size_t)g * C::kAGroup + wg * (C::kAGroup / 2)), src + (size_t)g * 512,
                         (C::kAGroup / 2) * 8, afull);
        }
    };

    int stage = 0;
    unsigned par = 0, uphase = 0;
    bool ready = false;   // the current item is already known to have landed
    long long unit = L0 / nchunks;
    int c_begin = (int)(L0 - unit * nchunks);
    int rem = (int)(L1 - L0);   // chunk-units left in the span
    load_unit(unit);
    if (wg == 1 && a.lag_ns > 0) __nanosleep((unsigned)a.lag_ns);   // the lag of row group 1
#pragma unroll 1
    while (rem > 0) {
        const int c_end = min(nchunks, c_begin + rem);
        mbar_wait(afull, uphase);
        uphase ^= 1u;

#pragma unroll 1
        for (int chunk = c_begin; chunk < c_end; ++chunk) {
#pragma unroll 1
            for (int slab = 0; slab < nslab; ++slab) {
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
                for (; g + 1 < ng; g += 2) {
                    mma_group(As + g * C::kAGroup, Bs + g * 512);
                    mma_group(As + (g + 1) * C::kAGroup, Bs + (g + 1) * 512);
                }
                if (g < ng) mma_group(As + g * C::kAGroup, Bs + g * 512);

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

        // ---- unit epilogue of the row group: merge the 16 threads sharing each scenario row
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
                    gred_val[(x * 4 + wx) * GR + row] = v;
                    gred_idx[(x * 4 + wx) * GR + row] = i;
                }
                best[x][r] = -INFINITY;
                bidx[x][r] = -1;
            }
        }
        group_barrier(1 + wg, 128);
        // every warp of the group is past its last read of the resident half: the next unit's
        // copies run under the merge below
        if (rem > c_end - c_begin) load_unit(unit + 1);
        const bool complete = (c_begin == 0 && c_end == nchunks);
        for (int q = gtid; q < GR * NX; q += 128) {
            const int x = q / GR, row = q % GR;
            double v = gred_val[(x * 4) * GR + row];
            int i = gred_idx[(x * 4) * GR + row];
#pragma unroll
            for (int w = 1; w < 4; ++w) {
                double ov = gred_val[(x * 4 + w) * GR + row];
                int oi = gred_idx[(x * 4 + w) * GR + row];
                if (better(ov, oi, v, i)) { v = ov; i = oi; }
            }
            const int urow = wg * GR + row;   // row within the 128-scenario unit
            if (complete) {
                const long long sc = unit * ROWS + urow;
                if (sc < a.n_local) {
                    a.best_val[x * a.out_stride + sc] = v;
                    a.best_idx[x * a.out_stride + sc] = i;
                }
            } else {   // a part of the unit: slot 0 = starts after chunk 0, slot 1 = starts at chunk 0
                const size_t o = (((size_t)blockIdx.x * 2 + (c_begin > 0 ? 0 : 1)) * NX + x) * ROWS + urow;
                a.piece_val[o] = v;
                a.piece_idx[o] = i;
            }
        }
        group_barrier(1 + wg, 128);   // the group's red_* are rewritten at the next unit's end
        rem -= c_end - c_begin;
        c_begin = 0;
        ++unit;
    }
}

}  // namespace sqlp
