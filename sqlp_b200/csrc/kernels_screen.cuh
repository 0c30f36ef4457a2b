// kernels_screen.cuh -- the SCREENING pass of the argmax on the 5th-generation tensor cores, and the
// exact FP64 decision among the vertices it leaves.
//
// Reference: argmax_procedure, src/sd_algorithm/subprob.jl:148-166 -- for every scenario the FIRST pool
// vertex maximising  score[k, i] = bias_x[k] + PiS[k, :] . d_i.  The FP64 sweep of kernels_contract_ws.cuh
// evaluates all K x N scores exactly.  Only the vertices that can still win need that: this file computes
// every score APPROXIMATELY with tcgen05.mma (bf16 operands, fp32 accumulators in tensor memory), bounds
// the error rigorously, keeps per scenario the vertices whose upper bound reaches the best lower bound, and
// evaluates those exactly -- with the same mma.sync.m8n8k4.f64 chain, in the same order, as the FP64 sweep,
// so best_idx AND best_val are bit-identical to it (the true argmax is provably among the candidates, and
// so is every vertex tying with it, hence "first index among the maxima" is preserved).
//
// Approximate score.  Every fp64 operand is split into two bf16 parts, p = ph + pl + rp with
// |rp| <= 2^-16 |p| (two successive round-to-nearest conversions, the second of the exact remainder),
// likewise d = dh + dl + rd.  Three products are kept, dh ph + dl ph + dh pl, each exact in fp32; they are
// accumulated by the tensor core in fp32 as ONE accumulation of 3 * sp terms (sp = row slots padded to 16):
//     per K-step of 16 slots:  D += [dh | dl](A, 128 scenarios) x [ph | ph]    (two instructions)
//                              D += [dh](A)                     x [pl]         (one instruction)
// The operands are CENTRED first (see "Centred operands" below): P'_k = PiS[k] - ctr, d'_i = d_i - dbar, and the
// bias carries P_k . dbar; a term that depends on the scenario alone does not change the argmax.
// Error budget per (k, i), Q = ||P'_k||_2 ||d'_i||_2 >= sum_j |p'_j d'_j| (Cauchy-Schwarz):
//     dropped products            <= 3 * 2^-16 * (1 + 2^-8)       Q
//     fp32 accumulation, any order, truncating adders allowed      <= (3 sp + 8) * 2^-22 * 1.02  Q
//     bias rounded to fp32 after a common shift, final fp32 add    <= 2^-23 (|b32[k]| + 1.02 Q)
//     FP64 operations on the uncentred operands (the sweep's own chain and bias add, P_k . dbar, the centring)
//                                                                  <= eabs  (k_screen_prep)
// E(chunk, i) = coef_q * max_k ||P'_k|| (256 vertices) * ||d'_i|| + coef_b * max |b32| + eabs,
// all norms rounded up, thresholds rounded towards "more candidates".
//
// Chain of a pass (host_epi.cuh:screen_enqueue): operand syncs (k_screen_view_sync, k_screen_scen_sync,
// k_screen_cd: new columns / scenarios only), k_screen_pdbar, k_screen_prep, k_screen_seed (warm start from the
// previous winners), k_screen, k_screen_decide (k_screen_resolve when the sweep was split in K-ranges), and the FP64
// sweep gated on the pass's control block behind it.
//
// Data layouts (tc05.cuh): scenario store DB[unit][part hi,lo][slab][128 scenarios][8] bf16, one contiguous
// block of 512 sp bytes per unit; pool view PiB[chunk of 256 vertices][K-step][hi,lo][slab 2][256][8] bf16,
// one contiguous 16 KB block per (chunk, K-step) = one pipeline stage; shifted fp32 biases
// b32c[chunk][NX * 256 + 4] with the chunk's largest vertex norm in the last slot group.
//
// Kernel k_screen: one persistent CTA per SM, 320 threads:
//   warps 0-7  epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31 (one scenario per thread), columns
//              128 (w / 4) .. +127 of the 128 x 256 accumulator; per score one FADD per point and a share of a
//              3-input max; every 32 columns ONE test of the eight group maxima against the running thresholds;
//              a hit appends (vertex, upper bound) to the thread's own candidate list -- no atomics, no cross-lane
//              traffic.
//   warp 8     producer: bulk async copies (TMA 1-D) of the unit's scenarios (resident for the whole sweep),
//              of the biases of a chunk and of the ring stages.
//   warp 9     one thread issues tcgen05.mma; tcgen05.commit releases ring stages and publishes accumulators.
// Two accumulators of 256 columns (all 512 TMEM columns) alternate, so the epilogue of chunk c runs under
// the MMAs of chunk c + 1.
#pragma once
#include "common.cuh"
#include "tc05.cuh"
#include "kernels_contract.cuh"

namespace sqlp {

#define SCR_THREADS 320
#define SCR_NB 256                 // vertices per screening chunk (N of the MMA)
#define SCR_UNIT 128               // scenarios per unit (M of the MMA)
#define SCR_STAGE_BYTES 16384      // one K-step of a chunk: [hi, lo][2 slabs][256][8] bf16
#define SCR_MAX_STAGES 8
#define SCR_BIAS_BUFS 4
#define SCR_CAP 64                 // candidate list entries per (scenario, point, K-range, column half)
#define SCR_CENTRE_COLS 1024       // the view's centre is the mean of at most this many columns,
#define SCR_CENTRE_SCEN 256        // an epigraph's of at most this many scenarios
#define SCR_DEAD (-3.0e38f)        // shifted bias of a vertex that can never win (or does not exist)

template <int NX>
__host__ __device__ constexpr int scr_bias_floats() { return NX * SCR_NB + 4; }

struct ScreenArgs {
    const __nv_bfloat16 *DB;       // [nunits][2][sp / 8][128][8]
    const __nv_bfloat16 *PiB;      // [nchunks][sp / 16][2][2][256][8]
    const float *b32c;             // [nchunks][NX * 256 + 4]
    const float *dnmax_unit;       // [nunits] largest ||d_i|| of the unit, rounded up
    const float *dn;               // [npad] ||d_i - dbar|| per scenario, rounded up (the epilogue thread owns ONE scenario)
    ScreenCtl *ctl;
    const long long *d_K;
    int sp;                        // row slots padded to a multiple of 16
    int nunits;
    int R;                         // K-ranges a unit's sweep is split into (work items = nunits * R)
    int nstages;
    long long n_local;
    long long npad;                // nunits * 128
    int2 *cand;                    // [((x * R + r) * 2 + h) * npad + i][SCR_CAP]  (vertex, upper bound bits)
    int *cnt;                      // [((x * R + r) * 2 + h) * npad + i]
    float *lfin;                   // same shape: the thread's final lower bound of the best score
    const float *lseed;            // optional [NX][npad]: a lower bound of the scenario's best score known beforehand
    float *dbg;                    // optional [128][256] raw accumulators of the first tile of block 0
    int desc_mode;                 // 0: LBO = K-direction, SBO = row-group direction (tc05.cuh); 1: swapped (probe only)
};

// dynamic shared memory of k_screen
__host__ __device__ inline size_t scr_smem_bytes(int sp, int nstages, int nx)
{
    return 2048 + (size_t)512 * sp + (size_t)nstages * SCR_STAGE_BYTES +
           (size_t)SCR_BIAS_BUFS * (nx * SCR_NB + 4) * 4;   // alignment slack + barrier block + A + ring + biases
}

// Two fp32 additions in one instruction (add.rn.f32x2, SASS FADD2: sm_100 packs a pair of fp32 lanes per register
// pair): round to nearest each, the same bits as two FADDs.
__device__ __forceinline__ void add2(float &r0, float &r1, uint32_t a0, uint32_t a1, float b0, float b1)
{
    unsigned long long x, y, z;
    asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "r"(a0), "r"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b0), "f"(b1));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(z) : "l"(x), "l"(y));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(z));
}

template <int NX, int PK = 1>
__global__ void __launch_bounds__(SCR_THREADS, 1) k_screen(ScreenArgs a)
{
    griddep_sync();
    if (screen_falls_back(a.ctl)) return;
    extern __shared__ __align__(16) unsigned char scr_raw[];
    // carve: [barriers + tmem pointer | A | B ring | bias buffers], 1 KB aligned
    const uint32_t raw0 = smem_u32(scr_raw);
    const uint32_t base = (raw0 + 1023u) & ~1023u;
    unsigned char *basep = scr_raw + (base - raw0);
    const int S = a.nstages;
    const int J = a.sp / 16;
    constexpr int BF = scr_bias_floats<NX>();
    const uint32_t bars = base;                                  // 8-byte mbarriers
    const uint32_t bfull0 = bars, bempty0 = bars + 8 * SCR_MAX_STAGES;
    const uint32_t tfull0 = bars + 16 * SCR_MAX_STAGES, tempty0 = tfull0 + 16;
    const uint32_t afull = tempty0 + 16, aempty = afull + 8;
    const uint32_t biasfull0 = aempty + 8, biasempty0 = biasfull0 + 8 * SCR_BIAS_BUFS;
    const uint32_t tmem_slot = biasempty0 + 8 * SCR_BIAS_BUFS;   // < base + 1024
    const uint32_t A_s = base + 1024;
    const uint32_t B_s = A_s + 512u * a.sp;
    const uint32_t bias_s = B_s + (uint32_t)S * SCR_STAGE_BYTES;
    const float *biasp = reinterpret_cast<const float *>(basep + 1024 + (size_t)512 * a.sp + (size_t)S * SCR_STAGE_BYTES);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long K = *a.d_K;
    const int nch = (int)((K + SCR_NB - 1) / SCR_NB);
    const int cpr = (nch + a.R - 1) / max(a.R, 1);               // chunks per K-range
    const int items = a.nunits * a.R;

    if (tid == 0) {
        for (int q = 0; q < S; ++q) { mbar_init(bfull0 + 8 * q, 1); mbar_init(bempty0 + 8 * q, 1); }
        for (int q = 0; q < 2; ++q) { mbar_init(tfull0 + 8 * q, 1); mbar_init(tempty0 + 8 * q, 8); }
        mbar_init(afull, 1);
        mbar_init(aempty, 1);
        for (int q = 0; q < SCR_BIAS_BUFS; ++q) { mbar_init(biasfull0 + 8 * q, 1); mbar_init(biasempty0 + 8 * q, 8); }
        mbar_fence_init();
    }
    if (warp == 9) tc05::tmem_alloc(tmem_slot, 512);
    tc05::fence_before_sync();
    __syncthreads();
    tc05::fence_after_sync();
    const uint32_t tmem = *reinterpret_cast<const volatile uint32_t *>(basep + (tmem_slot - base));

    if (warp == 8) {
        // ------------------------------------------------------------------ producer ------------
        if (lane == 0) {
            int stage = 0;
            unsigned sph = 0, bb = 0, bph = 0, it = 0;
            for (int w = blockIdx.x; w < items; w += gridDim.x) {
                const int r = w / a.nunits, u = w - r * a.nunits;
                const int c0 = r * cpr, c1 = min(nch, c0 + cpr);
                if (c0 >= c1) continue;
                mbar_wait(aempty, (it & 1u) ^ 1u);               // every MMA of the previous item has read A
                mbar_arrive_expect_tx(afull, 512u * a.sp);
                bulk_g2s(A_s, a.DB + (size_t)u * 256 * a.sp, 512u * a.sp, afull);
                ++it;
                for (int c = c0; c < c1; ++c) {
                    mbar_wait(biasempty0 + 8 * bb, bph ^ 1u);
                    mbar_arrive_expect_tx(biasfull0 + 8 * bb, BF * 4);
                    bulk_g2s(bias_s + bb * BF * 4, a.b32c + (size_t)c * BF, BF * 4, biasfull0 + 8 * bb);
                    if (++bb == SCR_BIAS_BUFS) { bb = 0; bph ^= 1u; }
                    const unsigned char *src = reinterpret_cast<const unsigned char *>(a.PiB) + (size_t)c * J * SCR_STAGE_BYTES;
                    for (int j = 0; j < J; ++j) {
                        mbar_wait(bempty0 + 8 * stage, sph ^ 1u);
                        mbar_arrive_expect_tx(bfull0 + 8 * stage, SCR_STAGE_BYTES);
                        bulk_g2s(B_s + stage * SCR_STAGE_BYTES, src + (size_t)j * SCR_STAGE_BYTES, SCR_STAGE_BYTES,
                                 bfull0 + 8 * stage);
                        if (++stage == S) { stage = 0; sph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 9) {
        // ------------------------------------------------------------------ MMA issuer ----------
        if (lane == 0) {
            const uint32_t idesc = tc05::idesc_bf16_f32(SCR_UNIT, SCR_NB);
            const uint32_t part = (uint32_t)(a.sp / 8) * 2048u;   // bytes of A's hi part (sp / 8 slabs of 128 rows)
            int stage = 0;
            unsigned sph = 0, t = 0, it = 0;
            for (int w = blockIdx.x; w < items; w += gridDim.x) {
                const int r = w / a.nunits;
                const int c0 = r * cpr, c1 = min(nch, c0 + cpr);
                if (c0 >= c1) continue;
                mbar_wait(afull, it & 1u);
                ++it;
                for (int c = c0; c < c1; ++c, ++t) {
                    const unsigned b = t & 1u;
                    mbar_wait(tempty0 + 8 * b, ((t >> 1) & 1u) ^ 1u);   // the epilogue has drained this accumulator
                    tc05::fence_after_sync();
                    const uint32_t d_tmem = tmem + b * SCR_NB;
                    for (int j = 0; j < J; ++j) {
                        mbar_wait(bfull0 + 8 * stage, sph);
                        tc05::fence_after_sync();
                        const uint32_t Ah = A_s + (uint32_t)j * 4096u, Al = Ah + part;
                        const uint32_t Bh = B_s + stage * SCR_STAGE_BYTES, Bl = Bh + 8192u;
                        const uint32_t la = a.desc_mode ? 128u : 2048u, sa = a.desc_mode ? 2048u : 128u;
                        const uint32_t lb = a.desc_mode ? 128u : 4096u, sb = a.desc_mode ? 4096u : 128u;
                        const uint64_t dAh = tc05::smem_desc(Ah, la, sa), dAl = tc05::smem_desc(Al, la, sa);
                        const uint64_t dBh = tc05::smem_desc(Bh, lb, sb), dBl = tc05::smem_desc(Bl, lb, sb);
                        tc05::mma_bf16_ss(d_tmem, dAh, dBh, idesc, j > 0 ? 1u : 0u);   // dh . ph
                        tc05::mma_bf16_ss(d_tmem, dAl, dBh, idesc, 1u);                // dl . ph
                        tc05::mma_bf16_ss(d_tmem, dAh, dBl, idesc, 1u);                // dh . pl
                        tc05::commit(bempty0 + 8 * stage);                             // stage free once these complete
                        if (++stage == S) { stage = 0; sph ^= 1u; }
                    }
                    tc05::commit(tfull0 + 8 * b);
                }
                tc05::commit(aempty);
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue warps 0-7 --
        const int q = warp & 3, h = warp >> 2;
        const int row = 32 * q + lane;
        const float coef_q = a.ctl->coef_q;
        float cb[NX];
#pragma unroll
        for (int x = 0; x < NX; ++x) cb[x] = __fmaf_ru(a.ctl->coef_b, a.ctl->bmax[x], a.ctl->eabs[x]);
        unsigned t = 0, bb = 0, bph = 0;
        unsigned long long emitted = 0;
        for (int w = blockIdx.x; w < items; w += gridDim.x) {
            const int r = w / a.nunits, u = w - r * a.nunits;
            const int c0 = r * cpr, c1 = min(nch, c0 + cpr);
            if (c0 >= c1) continue;
            const long long i = (long long)u * SCR_UNIT + row;
            const bool valid = i < a.n_local;
            // the bound of THIS thread's scenario: its own norm, not the unit's largest (E is per (vertex, scenario))
            const float dnu = a.dn ? (valid ? a.dn[i] : 0.f) : a.dnmax_unit[u];
            const float eq = __fmul_ru(coef_q, dnu);
            float L[NX];
            int n[NX];
            long long slot[NX];
#pragma unroll
            for (int x = 0; x < NX; ++x) {
                L[x] = (a.lseed && valid) ? a.lseed[(long long)x * a.npad + i] : -INFINITY;
                n[x] = 0;
                slot[x] = ((long long)(x * a.R + r) * 2 + h) * a.npad + i;
            }
            for (int c = c0; c < c1; ++c, ++t) {
                const unsigned b = t & 1u;
                mbar_wait(biasfull0 + 8 * bb, bph);
                const float *bs = biasp + bb * BF;
                const float pnmax = bs[NX * SCR_NB];
                float E[NX], E2[NX], thr[NX];
#pragma unroll
                for (int x = 0; x < NX; ++x) {
                    E[x] = __fmaf_ru(eq, pnmax, cb[x]);
                    E2[x] = __fadd_ru(E[x], E[x]);
                    thr[x] = __fadd_rd(L[x], -E[x]);              // candidate iff t >= L - E
                }
                mbar_wait(tfull0 + 8 * b, (t >> 1) & 1u);
                tc05::fence_after_sync();
                const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + b * SCR_NB + h * 128;
#pragma unroll 1
                for (int g = 0; g < 4; ++g) {
                    uint32_t v[32];
                    tc05::ld_32x32b_x32(taddr + g * 32, v);
                    tc05::wait_ld();
                    if (a.dbg && blockIdx.x == 0 && t == 0) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) a.dbg[row * SCR_NB + h * 128 + g * 32 + e] = __uint_as_float(v[e]);
                    }
                    // the eight (group of 8 columns, point) maxima of these 32 columns first, branch-free -- eight
                    // independent add / max trees the scheduler can interleave -- and ONE test whether any of them
                    // reaches its threshold; only then the groups are looked at one by one, in the order and with the
                    // running threshold of a column-by-column scan (thr only moves inside the slow path, so the pre-test
                    // with the thresholds at the start of the 32 columns can only err towards "look")
                    float gmx[4][NX];
                    bool any = false;
#pragma unroll
                    for (int s8 = 0; s8 < 4; ++s8) {
#pragma unroll
                        for (int x = 0; x < NX; ++x) {
                            const float4 b0 = *reinterpret_cast<const float4 *>(bs + x * SCR_NB + h * 128 + g * 32 + s8 * 8);
                            const float4 b1 = *reinterpret_cast<const float4 *>(bs + x * SCR_NB + h * 128 + g * 32 + s8 * 8 + 4);
                            float t0, t1, t2, t3, t4, t5, t6, t7;                 // score = accumulator + bias
                            if (PK) {                                             // two per FADD2
                                add2(t0, t1, v[s8 * 8 + 0], v[s8 * 8 + 1], b0.x, b0.y);
                                add2(t2, t3, v[s8 * 8 + 2], v[s8 * 8 + 3], b0.z, b0.w);
                                add2(t4, t5, v[s8 * 8 + 4], v[s8 * 8 + 5], b1.x, b1.y);
                                add2(t6, t7, v[s8 * 8 + 6], v[s8 * 8 + 7], b1.z, b1.w);
                            } else {
                                t0 = __uint_as_float(v[s8 * 8 + 0]) + b0.x; t1 = __uint_as_float(v[s8 * 8 + 1]) + b0.y;
                                t2 = __uint_as_float(v[s8 * 8 + 2]) + b0.z; t3 = __uint_as_float(v[s8 * 8 + 3]) + b0.w;
                                t4 = __uint_as_float(v[s8 * 8 + 4]) + b1.x; t5 = __uint_as_float(v[s8 * 8 + 5]) + b1.y;
                                t6 = __uint_as_float(v[s8 * 8 + 6]) + b1.z; t7 = __uint_as_float(v[s8 * 8 + 7]) + b1.w;
                            }
                            gmx[s8][x] = fmaxf(fmaxf(fmaxf(t0, t1), fmaxf(t2, t3)), fmaxf(fmaxf(t4, t5), fmaxf(t6, t7)));
                            any = any || gmx[s8][x] >= thr[x];
                        }
                    }
                    if (any) {
#pragma unroll
                        for (int s8 = 0; s8 < 4; ++s8) {
#pragma unroll
                            for (int x = 0; x < NX; ++x) {
                                const float gm = gmx[s8][x];
                                if (gm >= thr[x]) {
                                    // rare: some of the eight may still win.  Dead / padded vertices (SCR_DEAD) and
                                    // rows beyond the epigraph's scenarios never enter a list.
                                    const int kb = c * SCR_NB + h * 128 + g * 32 + s8 * 8;
                                    const float *bx = bs + x * SCR_NB + h * 128 + g * 32 + s8 * 8;
#pragma unroll
                                    for (int e = 0; e < 8; ++e) {
                                        const float te = __uint_as_float(v[s8 * 8 + e]) + bx[e];   // the same add as above
                                        if (te >= thr[x] && te > -1.0e38f && valid) {
                                            if (n[x] < SCR_CAP)
                                                a.cand[slot[x] * SCR_CAP + n[x]] =
                                                    make_int2(kb + e, __float_as_int(__fadd_ru(te, E[x])));
                                            ++n[x];
                                            ++emitted;
                                        }
                                    }
                                    thr[x] = fmaxf(thr[x], __fadd_rd(gm, -E2[x]));
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int x = 0; x < NX; ++x) L[x] = __fadd_rd(thr[x], E[x]);   // = max(L, best t of the tile - E)
                tc05::fence_before_sync();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(tempty0 + 8 * b);
                    mbar_arrive(biasempty0 + 8 * bb);
                }
                if (++bb == SCR_BIAS_BUFS) { bb = 0; bph ^= 1u; }
            }
            if (valid) {
#pragma unroll
                for (int x = 0; x < NX; ++x) {
                    a.cnt[slot[x]] = n[x];
                    a.lfin[slot[x]] = L[x];
                    if (n[x] > SCR_CAP) atomicAdd(&a.ctl->overflow, 1u);
                }
            }
        }
        // statistics, one atomic per warp
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) emitted += __shfl_xor_sync(0xffffffffu, emitted, off);
        if (lane == 0 && emitted) atomicAdd(&a.ctl->n_emit, emitted);
    }
    tc05::fence_before_sync();
    __syncthreads();
    if (warp == 9) {
        tc05::fence_after_sync();
        tc05::tmem_dealloc(tmem, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// Operand builders.  Both are idempotent over [lo, n) and run lazily before a screening pass.

__device__ __forceinline__ void split_bf16(double p, __nv_bfloat16 &hi, __nv_bfloat16 &lo)
{
    hi = __double2bfloat16(p);
    lo = __double2bfloat16(p - (double)__bfloat162float(hi));     // the remainder is exact in fp64
}
__device__ __forceinline__ float norm_up(double sumsq)
{
    return __double2float_ru(sqrt(sumsq) * (1.0 + 0x1p-40));
}
__device__ __forceinline__ void atomic_max_pos(float *addr, float v)   // v >= 0: integer order = float order
{
    atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
}

// Centred operands.  The argmax over k of  bias_k + P_k . d_i  does not change when a term that depends on i
// alone is taken away, and
//     bias_k + P_k . d_i  =  [bias_k + P_k . dbar]  +  (P_k - ctr) . (d_i - dbar)  +  ctr . (d_i - dbar)
// for ANY fixed vectors ctr and dbar: the pass multiplies the centred operands, adds P_k . dbar (FP64, once per call,
// k_screen_pdbar) to the bias and never forms the last term.  Every term of the error bound is proportional to
// ||P_k - ctr|| ||d_i - dbar||; real pools are clouds around a common point (storm: ||P|| ~ 8 300 but
// ||P - mean|| ~ 2 000, ||d|| ~ 1 080 but ||d - mean|| ~ 340), so the bound shrinks tenfold and the candidate lists
// with it.  ctr = mean of the view's first columns, dbar = mean of the epigraph's first scenarios; both are frozen
// between (rare, geometric) re-centrings that rebuild the bf16 operands.  Layout: sp values, then the 2-norm.
__global__ void k_screen_centre(const double *__restrict__ pi, int m2, const int *__restrict__ s_rows, int n_rows, int sp,
                                const int *__restrict__ act, const long long *__restrict__ d_K, long long n_max,
                                double *__restrict__ ctr)
{
    griddep_sync();
    __shared__ double red[32];
    const long long n = min(*d_K, n_max);
    double ss = 0.0;
    for (int j = threadIdx.x; j < sp; j += blockDim.x) {
        double m = 0.0;
        if (j < n_rows && n > 0) {
            const int r = s_rows[j];
            for (long long k = 0; k < n; ++k) m += pi[(act ? (long long)act[k] : k) * m2 + r];
            m /= (double)n;
            if (!isfinite(m)) m = 0.0;
        }
        ctr[j] = m;
        ss = fma(m, m, ss);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        ctr[sp] = sqrt(t) * (1.0 + 0x1p-40);
    }
}

// The same for the scenarios of an epigraph: mean of the first n local scenarios (FP64 tiles), slot by slot.
__global__ void k_screen_dbar(const double *__restrict__ D, int s_pad, int sp, long long n, double *__restrict__ dbar)
{
    griddep_sync();
    __shared__ double red[32];
    double ss = 0.0;
    for (int j = threadIdx.x; j < sp; j += blockDim.x) {
        double m = 0.0;
        if (j < s_pad && n > 0) {
            for (long long i = 0; i < n; ++i) m += D[(i >> 7) * (long long)s_pad * SQLP_TILE + tile_off((int)(i & 127), j)];
            m /= (double)n;
            if (!isfinite(m)) m = 0.0;
        }
        dbar[j] = m;
        ss = fma(m, m, ss);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        dbar[sp] = sqrt(t) * (1.0 + 0x1p-40);
    }
}

// pdb[k] = P_k . dbar for every column of the view (row-major copy: one warp reads one contiguous row).
__global__ void k_screen_pdbar(const double *__restrict__ PiR, int s_pad, const double *__restrict__ dbar,
                               const long long *__restrict__ d_K, double *__restrict__ pdb)
{
    griddep_sync();
    const long long K = *d_K;
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long k = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < K; k += nw) {
        const double *row = PiR + k * (long long)s_pad;
        double acc = 0.0;
        for (int j = lane; j < s_pad; j += 32) acc = fma(row[j], dbar[j], acc);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) pdb[k] = acc;
    }
}

// Pool view for screening: vertices [k_lo, K) restricted to the stochastic rows, minus the centre, split in bf16
// hi / lo, plus ||PiS[k] - ctr||_2 (rounded up) and its maximum per chunk of 256.  One warp per vertex.
// k runs over the view's columns [*mark, K); act (twins, kernels_pool.cuh) maps a column to its pool slot.  The
// mark is advanced by k_screen_mark, the next kernel of the stream.
__global__ void k_screen_view_sync(const double *__restrict__ pi, int m2, const int *__restrict__ s_rows, int n_rows,
                                   int sp, __nv_bfloat16 *__restrict__ PiB, float *__restrict__ pn,
                                   float *__restrict__ pnmax, int *__restrict__ bad, long long *__restrict__ mark,
                                   const long long *__restrict__ d_K, const int *__restrict__ act,
                                   const double *__restrict__ ctr)
{
    griddep_sync();
    const long long K = *d_K, k_lo = *mark;
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    const int J = sp / 16;
    for (long long k = k_lo + (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < K; k += nw) {
        const long long c = k / SCR_NB;
        const int v = (int)(k % SCR_NB);
        const long long kp = act ? (long long)act[k] : k;
        double ss = 0.0;
        for (int j = lane; j < sp; j += 32) {
            const double p = j < n_rows ? pi[kp * m2 + s_rows[j]] - (ctr ? ctr[j] : 0.0) : 0.0;
            __nv_bfloat16 hi, lo;
            split_bf16(p, hi, lo);
            const size_t off = ((size_t)(c * J + j / 16) * 2) * (2 * SCR_NB * 8) + (size_t)((j % 16) / 8) * (SCR_NB * 8) +
                               (size_t)v * 8 + (j % 8);
            PiB[off] = hi;
            PiB[off + 2 * SCR_NB * 8] = lo;
            ss = fma(p, p, ss);
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
        if (lane == 0) {
            const float nrm = norm_up(ss);
            pn[k] = nrm;
            if (nrm < 1.0e18f) atomic_max_pos(pnmax + c, nrm);
            else *bad = 1;                                        // Inf / NaN / out of bf16-safe range
        }
    }
}

__global__ void k_screen_mark(long long *__restrict__ mark, const long long *__restrict__ d_K)
{
    griddep_sync();
    if (threadIdx.x == 0 && blockIdx.x == 0) *mark = *d_K;
}

// Scenario store for screening: local scenarios [lo, n_local) from the FP64 tiles, minus dbar, split in bf16
// hi / lo, plus the largest ||d_i - dbar||_2 per unit of 128 and overall.  One warp per scenario.
__global__ void k_screen_scen_sync(const double *__restrict__ D, int s_pad, int sp, __nv_bfloat16 *__restrict__ DB,
                                   float *__restrict__ dnmax_unit, float *__restrict__ dnmax_all,
                                   int *__restrict__ bad, long long lo, long long n_local,
                                   const double *__restrict__ dbar, float *__restrict__ dn_out)
{
    griddep_sync();
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long i = lo + (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n_local; i += nw) {
        const long long u = i / SCR_UNIT;
        const int row = (int)(i % SCR_UNIT);
        const double *Dt = D + (i >> 7) * (long long)s_pad * SQLP_TILE;
        double ss = 0.0;
        for (int j = lane; j < sp; j += 32) {
            const double d = j < s_pad ? Dt[tile_off((int)(i & 127), j)] - (dbar ? dbar[j] : 0.0) : 0.0;
            __nv_bfloat16 hi, lw;
            split_bf16(d, hi, lw);
            const size_t off = (size_t)u * (2 * sp * SCR_UNIT) + (size_t)(j / 8) * (SCR_UNIT * 8) + (size_t)row * 8 + (j % 8);
            DB[off] = hi;
            DB[off + (size_t)sp * SCR_UNIT] = lw;
            ss = fma(d, d, ss);
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
        if (lane == 0) {
            const float nrm = norm_up(ss);
            if (dn_out) dn_out[i] = nrm;
            if (nrm < 1.0e18f) {
                atomic_max_pos(dnmax_unit + u, nrm);
                atomic_max_pos(dnmax_all, nrm);
            } else {
                *bad = 1;
            }
        }
    }
}

// Per call: shifted fp32 biases in the chunk-packed layout, the vertices that cannot win anywhere
// (Cauchy-Schwarz in exact arithmetic: bias_k + ||PiS_k|| max||d|| < max_k' (bias_k' - ||PiS_k'|| max||d||)),
// the error-bound coefficients and the fall-back decision.  One block.
template <int NX>
__global__ void __launch_bounds__(1024) k_screen_prep(const double *__restrict__ bias, long long bias_stride,
                                                      const float *__restrict__ pn, const float *__restrict__ pnmax,
                                                      const float *__restrict__ dnmax_all, const int *__restrict__ view_bad,
                                                      const int *__restrict__ epi_bad, const long long *__restrict__ d_K,
                                                      int sp, unsigned int ovf_limit, float *__restrict__ b32c,
                                                      ScreenCtl *__restrict__ ctl, const double *__restrict__ pdb,
                                                      const double *__restrict__ ctr, const double *__restrict__ dbar)
{
    griddep_sync();
    __shared__ double red[32];
    __shared__ double lbg[NX];
    __shared__ double babs[NX];
    __shared__ float fmx[NX];
    __shared__ int nlive[NX];
    __shared__ int sbad;
    const long long K = *d_K;
    const int nch = (int)((K + SCR_NB - 1) / SCR_NB);
    constexpr int BF = scr_bias_floats<NX>();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double dn = (double)*dnmax_all * (1.0 + 0x1p-20);
    if (tid == 0) sbad = (*view_bad | *epi_bad) ? 1 : 0;
    if (tid < NX) { fmx[tid] = 0.f; nlive[tid] = 0; }
    __syncthreads();
    for (int x = 0; x < NX; ++x) {
        double m = -INFINITY, ma = 0.0;
        bool pinf = false;
        for (long long k = tid; k < K; k += blockDim.x) {
            const double b = bias[x * bias_stride + k] + (pdb ? pdb[k] : 0.0);
            if (b == INFINITY) pinf = true;
            if (isfinite(b)) {
                m = fmax(m, b - (double)pn[k] * dn);
                ma = fmax(ma, fabs(b));
            }
        }
        if (pinf) sbad = 1;                                        // a +Inf score wins everywhere: leave it to the FP64 sweep
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
            ma = fmax(ma, __shfl_xor_sync(0xffffffffu, ma, off));
        }
        if (lane == 0) red[warp] = m;
        __syncthreads();
        if (tid == 0) {
            double mm = -INFINITY;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mm = fmax(mm, red[w]);
            lbg[x] = mm;
        }
        __syncthreads();
        if (lane == 0) red[warp] = ma;
        __syncthreads();
        if (tid == 0) {
            double mm = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mm = fmax(mm, red[w]);
            babs[x] = mm;
        }
        __syncthreads();
    }
    for (int x = 0; x < NX; ++x) {
        const double lb = lbg[x];
        float mx = 0.f;
        int live = 0;
        for (long long k = tid; k < (long long)nch * SCR_NB; k += blockDim.x) {
            float out = SCR_DEAD;
            if (k < K) {
                const double b = bias[x * bias_stride + k] + (pdb ? pdb[k] : 0.0);
                const double q = (double)pn[k] * dn;
                if (isfinite(b) && isfinite(lb) && b + q + 1e-9 * (fabs(lb) + q) >= lb) {
                    out = __double2float_rn(b - lb);
                    if (!(fabsf(out) < 1.0e30f)) sbad = 1;
                    mx = fmaxf(mx, fabsf(out));
                    ++live;
                }
            }
            b32c[(k / SCR_NB) * BF + x * SCR_NB + (k % SCR_NB)] = out;
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            live += __shfl_xor_sync(0xffffffffu, live, off);
        }
        if (lane == 0) {
            atomic_max_pos(&fmx[x], mx);
            atomicAdd(&nlive[x], live);
        }
    }
    float pmx = 0.f;
    for (int c = tid; c < nch; c += blockDim.x) {
        const float p = pnmax[c];
        b32c[(long long)c * BF + NX * SCR_NB] = p;
        b32c[(long long)c * BF + NX * SCR_NB + 1] = 0.f;
        b32c[(long long)c * BF + NX * SCR_NB + 2] = 0.f;
        b32c[(long long)c * BF + NX * SCR_NB + 3] = 0.f;
        pmx = fmaxf(pmx, p);
    }
    if ((double)pmx * dn > 1.0e30) sbad = 1;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) pmx = fmaxf(pmx, __shfl_xor_sync(0xffffffffu, pmx, off));
    __syncthreads();
    if (lane == 0) red[warp] = (double)pmx;
    __syncthreads();
    if (tid == 0) {
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) pmx = fmaxf(pmx, (float)red[w]);
        const double cq = (3.0 * 0x1p-16 * (1.0 + 0x1p-7) + (3.0 * sp + 8.0) * 0x1p-22 * 1.02 + 1.02 * 0x1p-23 + 0x1p-40) *
                          (1.0 + 0x1p-10);
        ctl->coef_q = __double2float_ru(cq);
        ctl->coef_b = __double2float_ru((0x1p-23 + 0x1p-40) * (1.0 + 0x1p-10));
        // What the FP64 arithmetic on the UNcentred operands may lose -- the sweep's chain of s_pad fused
        // multiply-adds and its bias add, the s_pad terms of P_k . dbar, the subtractions of ctr and dbar -- against
        // the magnitudes before centring: (sp + 8) 2^-52 (max |bias'_k| + (max||P'|| + ||ctr||) (max||d'|| + ||dbar||)).
        const double raw = ((double)pmx * (1.0 + 0x1p-20) + (ctr ? ctr[sp] : 0.0)) * (dn + (dbar ? dbar[sp] : 0.0));
        for (int x = 0; x < NX; ++x) {
            ctl->shift[x] = lbg[x];
            ctl->bmax[x] = fmx[x];
            ctl->live[x] = nlive[x];
            const double ea = (sp + 8.0) * 0x1p-52 * (babs[x] + raw) * (1.0 + 0x1p-10) + 1e-30;
            ctl->eabs[x] = ea < 1.0e30 ? __double2float_ru(ea) : 1.0e30f;
            if (!(ea < 1.0e30)) sbad = 1;
        }
        double cdb = 0.0;
        if (ctr && dbar)
            for (int j = 0; j < sp; ++j) cdb = fma(ctr[j], dbar[j], cdb);
        ctl->cdb = cdb;
        ctl->bad = sbad;
        ctl->overflow = 0u;
        ctl->ovf_limit = ovf_limit;
        ctl->n_emit = 0ull;
        ctl->n_eval = 0ull;
    }
}

// ------------------------------------------------------------------------------------------------
// Warm start.  The epilogue of k_screen keeps every vertex that could still win when it is seen, so a list also
// holds the "record breakers" of the scan -- about (candidates) x ln K entries, most of them stale at the end --
// and the warp takes its slow path whenever one of its 32 scenarios meets one.  An SD iteration moves the candidate
// point a little and the incumbent rarely: the vertex that won a scenario at the previous call is almost always a
// near-winner now.  Its score at the new point is a lower bound of the scenario's best score that is valid whatever
// the history (ANY vertex gives one), so the scan starts from it instead of from -Inf.
// It costs no dot product: the exact decision already computed P_a . d_i for the winner a (it does not depend on
// the point, nor on time: vertices and scenarios never change), and in the pass's centred and shifted scale
//     s_a(i) = bias_a - shift + P_a . d_i - ctr . d_i + ctr . dbar,
// with ctr . d_i kept per scenario (k_screen_cd) and ctr . dbar per call (k_screen_prep).  Per scenario and point:
// the larger of the scores of the previous winners at both points, minus what the FP64 operations may have lost
// (the cancellation is on the uncentred magnitudes: that is what eabs bounds), rounded down to fp32.
struct SeedArgs {
    const double *bias;         // [NX][bias_stride]
    long long bias_stride;
    const int *prev;            // [n][2]: 1 + view column, 0 = none
    const double *prevdot;      // [n][2]: P_k . d_i of that column
    const double *cd;           // [n]: ctr . d_i
    const long long *d_K;
    const ScreenCtl *ctl;
    long long n_local, npad;
    float *lseed;               // [NX][npad]
};

template <int NX>
__global__ void __launch_bounds__(256) k_screen_seed(SeedArgs a)
{
    griddep_sync();
    const long long K = *a.d_K;
    const double cdb = a.ctl->cdb;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.npad; i += (long long)gridDim.x * blockDim.x) {
        int ka = -1, kb = -1;
        double da = 0.0, db = 0.0;
        if (i < a.n_local) {
            const int2 pc = reinterpret_cast<const int2 *>(a.prev)[i];
            const double2 pd = reinterpret_cast<const double2 *>(a.prevdot)[i];
            const double off = cdb - a.cd[i];
            ka = pc.x - 1;
            kb = pc.y - 1;
            da = pd.x + off;
            db = pd.y + off;
        }
        if (ka >= K) ka = -1;
        if (kb >= K || kb == ka) kb = -1;
#pragma unroll
        for (int x = 0; x < NX; ++x) {
            double best = -INFINITY;
            if (ka >= 0) {
                const double v = a.bias[x * a.bias_stride + ka] - a.ctl->shift[x] + da;
                if (isfinite(v)) best = v;
            }
            if (kb >= 0) {
                const double v = a.bias[x * a.bias_stride + kb] - a.ctl->shift[x] + db;
                if (isfinite(v)) best = fmax(best, v);
            }
            best -= 2.0 * (double)a.ctl->eabs[x] + 0x1p-40 * fabs(best);
            float out = -INFINITY;
            if (best > -1.0e37 && best < 1.0e37) out = __double2float_rd(best);
            a.lseed[(long long)x * a.npad + i] = out;
        }
    }
}

// cd[i] = ctr . d_i for the local scenarios [lo, n): one warp per scenario over the row-major store.
__global__ void k_screen_cd(const double *__restrict__ DR, int s_pad, const double *__restrict__ ctr, long long lo,
                            long long n, double *__restrict__ cd)
{
    griddep_sync();
    const int lane = threadIdx.x & 31;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long i = lo + (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += nw) {
        double acc = 0.0;
        for (int j = lane; j < s_pad; j += 32) acc = fma(ctr[j], DR[i * (long long)s_pad + j], acc);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) cd[i] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// The exact decision.  One warp per scenario: the candidates that survive the final lower bound are scored
// with the FP64 sweep's own arithmetic -- mma.sync.m8n8k4.f64 over the k-groups in ascending order from a
// zero accumulator, then + bias -- eight vertices per chain (the eight rows of A all hold the scenario), and
// compared on (value desc, index asc).  A scenario whose list overflowed is swept over all K vertices.
struct ResolveArgs {
    const double *D;            // FP64 scenario tiles (fragment-major)
    const double *PiS;          // FP64 pool view (fragment-major): what the sweep multiplies
    const double *PiR;          // the same values row-major [column][s_pad]: one contiguous row per candidate
    const double *DR;           // the scenarios row-major [i][s_pad] (k_screen_decide)
    const double *bias;         // [NX][bias_stride]
    long long bias_stride;
    int s_pad;
    const long long *d_K;
    long long n_local, npad;
    int R;
    const int2 *cand;
    const int *cnt;
    const float *lfin;
    double *best_val;           // [NX][out_stride]
    int *best_idx;
    long long out_stride;
    ScreenCtl *ctl;
    int force_full;             // tests: sweep every vertex for every scenario with this kernel's arithmetic
    int *prev;                  // optional [n][2]: 1 + the view column selected for (scenario, point), for k_screen_seed
    double *prevdot;            // with prev: the selected vertex's dot P_k . d_i
};

// FMA (0 / 1): the chain as mma.sync.m8n8k4.f64 (eight candidates per chain, one of the eight rows used), or as plain
// DFMA in the SAME order -- k-groups ascending, the four products of a group in k order from the running sum, which
// is what the DMMA computes bit for bit (checked on the device at context creation, sqlp_api.cu:ctx_selftest; the
// library uses the lanes variant only if that check passes): 32 candidates per chain, one per lane.
template <int NX, int FMA>
__global__ void __launch_bounds__(256) k_screen_resolve(ResolveArgs a)
{
    griddep_wait();
    if (!a.force_full && screen_falls_back(a.ctl)) return;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long nw = (long long)gridDim.x * (blockDim.x >> 5);
    const long long K = *a.d_K;
    const int nch = (int)((K + SCR_NB - 1) / SCR_NB);
    const int cpr = (nch + a.R - 1) / max(a.R, 1);
    const int ng = a.s_pad / 4;
    const size_t tile_doubles = (size_t)a.s_pad * SQLP_TILE;
    __shared__ int sq[8][32];                 // the warp's queue of candidates: vertex | point << 30  (lanes variant)
    unsigned long long evald = 0;
    for (long long i = (long long)blockIdx.x * (blockDim.x >> 5) + wib; i < a.n_local; i += nw) {
        double best[NX];
        int bidx[NX];
#pragma unroll
        for (int x = 0; x < NX; ++x) { best[x] = -INFINITY; bidx[x] = -1; }
        const int ci = (int)(i & 127);
        const double *Dtile = a.D + (size_t)(i >> 7) * tile_doubles;
        int qk = 0, qx = 0, qn = 0;
        // ---- DMMA chain: up to 8 (vertex, point) pairs held by lanes 0..7 (qk, qx), qn of them ------------------
        auto flush8 = [&]() {
            if (qn == 0) return;
            const double *Drow = Dtile + ((((size_t)(ci >> 4)) * 32 + (ci & 7) * 4 + (lane & 3)) << 1) + ((ci >> 3) & 1);
            const int kfirst = __shfl_sync(0xffffffffu, qk, 0);
            int kk = __shfl_sync(0xffffffffu, qk, lane >> 2);
            if ((lane >> 2) >= qn) kk = kfirst;                      // unused columns repeat a valid vertex
            // B fragment of lane t: vertex kk (column t / 4 of the chain), slot 4 g + t % 4 -- four lanes read one
            // 32-byte sector of the vertex's row
            const double *Prow = a.PiR + (size_t)kk * a.s_pad + (lane & 3);
            double acc0 = 0.0, acc1 = 0.0;
            int g = 0;
            for (; g + 6 <= ng; g += 6) {                            // twelve loads in flight, then the ordered chain
                double av[6], bv[6];
#pragma unroll
                for (int u = 0; u < 6; ++u) { av[u] = Drow[(size_t)(g + u) * 512]; bv[u] = Prow[(size_t)(g + u) * 4]; }
#pragma unroll
                for (int u = 0; u < 6; ++u)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                 : "+d"(acc0), "+d"(acc1)
                                 : "d"(av[u]), "d"(bv[u]));
            }
            for (; g < ng; ++g) {
                const double av = Drow[(size_t)g * 512], bv = Prow[(size_t)g * 4];
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(acc0), "+d"(acc1)
                             : "d"(av), "d"(bv));
            }
            for (int j = 0; j < qn; ++j) {
                const double dot = __shfl_sync(0xffffffffu, (j & 1) ? acc1 : acc0, j >> 1);
                const int kj = __shfl_sync(0xffffffffu, qk, j);
                const int xj = __shfl_sync(0xffffffffu, qx, j);
                if (kj < K) {
#pragma unroll
                    for (int x = 0; x < NX; ++x) {
                        const double v = dot + a.bias[x * a.bias_stride + kj];   // the sweep's epilogue: acc + bias
                        if ((xj < 0 || x == xj) && (v > best[x] || (v == best[x] && bidx[x] >= 0 && kj < bidx[x]))) {
                            best[x] = v;
                            bidx[x] = kj;
                        }
                    }
                }
            }
            evald += qn;
            qn = 0;
        };
        // ---- lanes variant: lane l scores queued pair l with the DFMA chain; then a warp argmax per point -----
        auto flush32 = [&]() {
            if (qn == 0) return;
            if (!FMA || qn <= 8) {
                // few candidates (the usual case on pools with clear winners): DMMA chains of eight are quicker
                const int total = qn;
                for (int b0 = 0; b0 < total; b0 += 8) {
                    const int m = min(8, total - b0);
                    if (lane < m) {
                        const int e8 = sq[wib][b0 + lane];
                        qk = e8 & 0x3FFFFFFF;
                        qx = ((e8 >> 30) & 3) == 2 ? -1 : ((e8 >> 30) & 3);
                    }
                    qn = m;
                    flush8();
                }
                qn = 0;
                __syncwarp();
                return;
            }
            const int e = lane < qn ? sq[wib][lane] : sq[wib][0];
            const int kk = e & 0x3FFFFFFF, xq = (e >> 30) & 3;           // xq = 2: every point (full sweeps)
            // the candidate's row is contiguous (s_pad doubles, 64-byte aligned): 16-byte loads, every sector used
            const double2 *Pc = reinterpret_cast<const double2 *>(a.PiR + (size_t)kk * a.s_pad);
            const double *Dc = Dtile + ((((size_t)(ci >> 4)) * 32 + (ci & 7) * 4) << 1) + ((ci >> 3) & 1);
            double acc = 0.0;
            int g = 0;
            for (; g + 4 <= ng; g += 4) {
                double2 pv[8];
                double dv[16];
#pragma unroll
                for (int u = 0; u < 8; ++u) pv[u] = Pc[(size_t)g * 2 + u];
#pragma unroll
                for (int u = 0; u < 16; ++u) dv[u] = Dc[(size_t)(g + (u >> 2)) * 512 + (u & 3) * 2];
#pragma unroll
                for (int u = 0; u < 8; ++u) {                        // slots in order: the chain of the sweep's DMMA
                    acc = fma(dv[2 * u], pv[u].x, acc);
                    acc = fma(dv[2 * u + 1], pv[u].y, acc);
                }
            }
            for (; g < ng; ++g) {
                const double2 p0 = Pc[(size_t)g * 2], p1 = Pc[(size_t)g * 2 + 1];
                acc = fma(Dc[(size_t)g * 512], p0.x, acc);
                acc = fma(Dc[(size_t)g * 512 + 2], p0.y, acc);
                acc = fma(Dc[(size_t)g * 512 + 4], p1.x, acc);
                acc = fma(Dc[(size_t)g * 512 + 6], p1.y, acc);
            }
#pragma unroll
            for (int x = 0; x < NX; ++x) {
                double v = -INFINITY;
                int kv = -1;
                if (lane < qn && kk < K && (xq == 2 || xq == x)) {
                    const double t = acc + a.bias[x * a.bias_stride + kk];
                    if (t > -INFINITY) { v = t; kv = kk; }               // NaN and -Inf never win (subprob.jl:151-156)
                }
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, v, off);
                    const int oi = __shfl_xor_sync(0xffffffffu, kv, off);
                    if (oi >= 0 && (kv < 0 || ov > v || (ov == v && oi < kv))) { v = ov; kv = oi; }
                }
                if (kv >= 0 && (v > best[x] || (v == best[x] && bidx[x] >= 0 && kv < bidx[x]))) { best[x] = v; bidx[x] = kv; }
            }
            evald += qn;
            qn = 0;
            __syncwarp();
        };
        // the passing entries of a batch of 32 go to the warp's queue in list order, 32 at most per flush
        auto enqueue = [&](unsigned pass, int k, int x) {
            while (pass) {
                const int room = 32 - qn, npass = __popc(pass);
                const int rank = __popc(pass & ((1u << lane) - 1u));
                const bool mine = ((pass >> lane) & 1u) && rank < room;
                if (mine) sq[wib][qn + rank] = k | (x << 30);
                const unsigned took = __ballot_sync(0xffffffffu, mine);
                __syncwarp();
                qn += min(room, npass);
                pass &= ~took;
                if (qn == 32) flush32();
            }
        };
        // final lower bounds and overflow over the thread lists of this scenario
        float LB[NX];
        bool full = a.force_full != 0;
        if (a.R == 1) {
            // the usual case (enough scenarios to fill the GPU): 2 NX lists.  Lane l < 2 NX reads the count and the
            // bound of list l; then every entry of every list is requested before any is looked at, so the scenario
            // costs three dependent memory round trips (counts, entries, operands) whatever the number of lists.
            constexpr int L = 2 * NX;
            int myn = 0;
            float mylb = -INFINITY;
            if (lane < L && nch > 0) {
                const long long slot = (long long)lane * a.npad + i;       // ((x * 1 + 0) * 2 + h) = lane
                myn = a.cnt[slot];
                mylb = a.lfin[slot];
            }
            int nl[L];
#pragma unroll
            for (int l = 0; l < L; ++l) {
                nl[l] = __shfl_sync(0xffffffffu, myn, l);
                full = full || nl[l] > SCR_CAP;
            }
#pragma unroll
            for (int x = 0; x < NX; ++x)
                LB[x] = fmaxf(__shfl_sync(0xffffffffu, mylb, 2 * x), __shfl_sync(0xffffffffu, mylb, 2 * x + 1));
            if (!full) {
                int2 ent[L][SCR_CAP / 32];
#pragma unroll
                for (int l = 0; l < L; ++l)
#pragma unroll
                    for (int b = 0; b < SCR_CAP / 32; ++b) {
                        ent[l][b] = make_int2(0, 0);
                        if (b * 32 + lane < nl[l]) ent[l][b] = a.cand[((long long)l * a.npad + i) * SCR_CAP + b * 32 + lane];
                    }
#pragma unroll
                for (int l = 0; l < L; ++l)
#pragma unroll
                    for (int b = 0; b < SCR_CAP / 32; ++b) {
                        if (b * 32 >= nl[l]) continue;
                        const unsigned pass = __ballot_sync(0xffffffffu, b * 32 + lane < nl[l] &&
                                                                         __int_as_float(ent[l][b].y) >= LB[l >> 1]);
                        enqueue(pass, ent[l][b].x, l >> 1);
                    }
                flush32();
            }
        } else {
#pragma unroll
            for (int x = 0; x < NX; ++x) {
                float lb = -INFINITY;
                int ovf = 0;
                for (int s = lane; s < 2 * a.R; s += 32) {
                    const int r = s >> 1;
                    if (r * cpr >= nch) continue;                         // empty K-range: nothing was written
                    const long long slot = ((long long)(x * a.R) * 2 + s) * a.npad + i;
                    lb = fmaxf(lb, a.lfin[slot]);
                    ovf |= a.cnt[slot] > SCR_CAP;
                }
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    lb = fmaxf(lb, __shfl_xor_sync(0xffffffffu, lb, off));
                    ovf |= __shfl_xor_sync(0xffffffffu, ovf, off);
                }
                LB[x] = lb;
                full = full || ovf;
            }
            if (!full) {
#pragma unroll
                for (int x = 0; x < NX; ++x) {
                    for (int s = 0; s < 2 * a.R; ++s) {
                        if ((s >> 1) * cpr >= nch) continue;
                        const long long slot = ((long long)(x * a.R) * 2 + s) * a.npad + i;
                        const int n = min(a.cnt[slot], SCR_CAP);
                        for (int b0 = 0; b0 < n; b0 += 32) {
                            int2 e = make_int2(0, 0);
                            if (b0 + lane < n) e = a.cand[slot * SCR_CAP + b0 + lane];
                            const unsigned pass = __ballot_sync(0xffffffffu, b0 + lane < n && __int_as_float(e.y) >= LB[x]);
                            enqueue(pass, e.x, x);
                        }
                    }
                }
                flush32();
            }
        }
        if (full) {
            for (long long k0 = 0; k0 < K; k0 += 32) {              // point code 2: the dot serves every point
                if (k0 + lane < K) sq[wib][lane] = (int)(k0 + lane) | (2 << 30);
                __syncwarp();
                qn = (int)min((long long)32, K - k0);
                flush32();
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int x = 0; x < NX; ++x) {
                a.best_val[x * a.out_stride + i] = best[x];
                a.best_idx[x * a.out_stride + i] = bidx[x];
            }
        }
    }
    if (lane == 0 && evald) atomicAdd(&a.ctl->n_eval, evald);
    griddep_launch();
}

// The exact decision, second form (used when the sweep was not split in K-ranges).  k_screen_resolve gives every
// candidate a lane that walks the candidate's row by itself: five active lanes, five cache lines per load
// instruction for 80 useful bytes, ~760 instructions and ~70 dependent loads per scenario (ncu: 16 of 64 warps
// resident, issue slots 28 % busy).  Here the warp moves whole rows: the scenario's row and up to eight candidate
// rows (row-major copies) go to the warp's shared memory as 16-byte asynchronous copies -- one instruction = 512
// contiguous bytes, no registers, every copy of a batch in flight together -- and the chain is the sweep's own
// instruction, mma.sync.m8n8k4.f64 over the k-groups in ascending order from a zero accumulator with the eight
// candidates as the eight columns of B and the scenario in every row of A, then + bias: 30 dependent DMMAs for eight
// exact scores, bit for bit those of the sweep.  A vertex that is a candidate at both points is scored once: the
// dot does not depend on the point, and scoring a vertex at a point where it was no candidate is what the full
// sweep does anyway.
#define SCR_DEC_WARPS 4
#define SCR_DEC_ROWS 8
#define SCR_DEC_QUEUE 256            // 2 points x 2 column halves x SCR_CAP entries
__host__ __device__ inline int scr_decide_stride(int s_pad)    // doubles per staged row: 4 or 12 (mod 16), so that the
{                                                                  // B fragments of four rows fall in four bank groups
    return s_pad + (((s_pad & 15) == 0 || (s_pad & 15) == 8) ? 4 : 0);
}
__host__ __device__ inline size_t scr_decide_smem(int s_pad)
{
    return (size_t)SCR_DEC_WARPS * ((size_t)(SCR_DEC_ROWS + 2) * scr_decide_stride(s_pad) * 8 + SCR_DEC_ROWS * 8 + SCR_DEC_QUEUE * 4);
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}
__device__ __forceinline__ void cp_async_commit()
{
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}
__device__ __forceinline__ void cp_async_wait_but_last()     // every group but the most recent one has landed
{
    asm volatile("cp.async.wait_group 1;\n" ::: "memory");
}

template <int NX>
__global__ void __launch_bounds__(32 * SCR_DEC_WARPS) k_screen_decide(ResolveArgs a)
{
    griddep_wait();
    if (screen_falls_back(a.ctl)) return;
    static_assert(2 * NX * SCR_CAP <= SCR_DEC_QUEUE, "queue too small");
    extern __shared__ __align__(16) unsigned char dec_sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int rs = scr_decide_stride(a.s_pad);
    const int sp2 = a.s_pad / 2, ng = a.s_pad / 4;
    // per warp: two scenario rows (this scenario's and the next one's, on its way), eight candidate rows
    double *drow = reinterpret_cast<double *>(dec_sm) + (size_t)wib * (SCR_DEC_ROWS + 2) * rs;
    double *rows = drow + 2 * rs;
    double *dots = reinterpret_cast<double *>(dec_sm) + (size_t)SCR_DEC_WARPS * (SCR_DEC_ROWS + 2) * rs + wib * SCR_DEC_ROWS;
    int *wq = reinterpret_cast<int *>(reinterpret_cast<double *>(dec_sm) + (size_t)SCR_DEC_WARPS * ((SCR_DEC_ROWS + 2) * rs + SCR_DEC_ROWS)) +
              wib * SCR_DEC_QUEUE;
    const long long nw = (long long)gridDim.x * SCR_DEC_WARPS;
    const long long K = *a.d_K;
    const int nch = (int)((K + SCR_NB - 1) / SCR_NB);
    constexpr int L = 2 * NX;                 // lists of a scenario: (point, column half)
    constexpr int HEAD = 32 / L;              // entries of every list that ONE load instruction fetches (lane = list * HEAD + entry)
    unsigned long long evald = 0;
    // The head of a scenario -- counts and bounds of its lists (lanes < L), the first HEAD entries of every list, its
    // row -- is requested one scenario ahead, before the rows of the current one: of the three dependent round trips
    // (counts, entries, operand rows) only the last is waited for.  (Entries beyond a list's count are read and ignored.)
    auto request = [&](long long i, int par, int &myn, float &mylb, int2 &e8) {
        myn = 0;
        mylb = -INFINITY;
        e8 = make_int2(0, 0);
        if (nch > 0) {
            if (lane < L) {
                const long long slot = (long long)lane * a.npad + i;       // list (x, h) = lane 2 x + h
                myn = a.cnt[slot];
                mylb = a.lfin[slot];
            }
            e8 = a.cand[((long long)(lane / HEAD) * a.npad + i) * SCR_CAP + (lane % HEAD)];
        }
        const double2 *Dr = reinterpret_cast<const double2 *>(a.DR + (size_t)i * a.s_pad) + lane;
        double2 *r2 = reinterpret_cast<double2 *>(drow + (size_t)par * rs) + lane;
        if (lane < sp2) cp_async16(r2, Dr);
        if (lane + 32 < sp2) cp_async16(r2 + 32, Dr + 32);
        for (int q = 64; q + lane < sp2; q += 32) cp_async16(r2 + q, Dr + q);
    };
    long long i = (long long)blockIdx.x * SCR_DEC_WARPS + wib;
    int par = 0, myn = 0;
    float mylb = -INFINITY;
    int2 e8 = make_int2(0, 0);
    if (i < a.n_local) request(i, par, myn, mylb, e8);
    cp_async_commit();
    for (; i < a.n_local; i += nw, par ^= 1) {
        double best = -INFINITY, bdot = 0.0;      // of point x = lane (lanes < NX)
        int bidx = -1;
        int nl[L];
        bool full = false;
#pragma unroll
        for (int l = 0; l < L; ++l) {
            nl[l] = __shfl_sync(0xffffffffu, myn, l);
            full = full || nl[l] > SCR_CAP;
        }
        float LB[NX];
#pragma unroll
        for (int x = 0; x < NX; ++x)
            LB[x] = fmaxf(__shfl_sync(0xffffffffu, mylb, 2 * x), __shfl_sync(0xffffffffu, mylb, 2 * x + 1));
        int n = 0;
        if (!full) {
            // the heads of all lists in one ballot
            int mycnt = 0;
            float mylbx = INFINITY;
#pragma unroll
            for (int l = 0; l < L; ++l)
                if (lane / HEAD == l) { mycnt = nl[l]; mylbx = LB[l >> 1]; }
            {
                const bool ok = (lane % HEAD) < mycnt && __int_as_float(e8.y) >= mylbx;
                const unsigned pass = __ballot_sync(0xffffffffu, ok);
                if (ok) wq[__popc(pass & ((1u << lane) - 1u))] = e8.x;
                n = __popc(pass);
            }
            // the rest of a long list (rare once the scan is warm-started)
#pragma unroll
            for (int l = 0; l < L; ++l) {
                for (int b0 = HEAD; b0 < nl[l]; b0 += 32) {
                    int2 e = make_int2(0, 0);
                    if (b0 + lane < nl[l]) e = a.cand[((long long)l * a.npad + i) * SCR_CAP + b0 + lane];
                    const bool ok = b0 + lane < nl[l] && __int_as_float(e.y) >= LB[l >> 1];
                    const unsigned pass = __ballot_sync(0xffffffffu, ok);
                    if (ok) wq[n + __popc(pass & ((1u << lane) - 1u))] = e.x;
                    n += __popc(pass);
                }
            }
            __syncwarp();
            // one chain per vertex: the second copy of a column (the other point's list) is struck out.  Queues of
            // more than 32 are left as they are -- a vertex scored twice is only work
            if (NX > 1 && n > 1 && n <= 32) {
                const int e = lane < n ? wq[lane] : -1 - lane;
                const unsigned same = __match_any_sync(0xffffffffu, e);
                const bool keep = lane < n && (__ffs(same) - 1) == lane;
                const unsigned km = __ballot_sync(0xffffffffu, keep);
                __syncwarp();
                if (keep) wq[__popc(km & ((1u << lane) - 1u))] = e;
                n = __popc(km);
                __syncwarp();
            }
        }
        const double *dcur = drow + (size_t)par * rs;
        auto stage = [&](int base, int nb) {       // candidate rows of a batch -> the warp's shared memory, asynchronously
            for (int r = 0; r < nb; ++r) {
                const double2 *P = reinterpret_cast<const double2 *>(a.PiR + (size_t)wq[base + r] * a.s_pad) + lane;
                double2 *r2 = reinterpret_cast<double2 *>(rows + (size_t)r * rs) + lane;
                if (lane < sp2) cp_async16(r2, P);
                if (lane + 32 < sp2) cp_async16(r2 + 32, P + 32);
                for (int q = 64; q + lane < sp2; q += 32) cp_async16(r2 + q, P + q);
            }
        };
        auto chain = [&](int base, int nb) {
            // A: the scenario in every row (lane t: slot 4 g + t % 4); B: column t / 4 = staged row t / 4.  Columns
            // beyond nb read whatever the rows hold: a column's result depends on that column alone
            const double *ap = dcur + (lane & 3);
            const double *bp = rows + (size_t)(lane >> 2) * rs + (lane & 3);
            double acc0 = 0.0, acc1 = 0.0;
            int g = 0;
            for (; g + 6 <= ng; g += 6) {
                double av[6], bv[6];
#pragma unroll
                for (int u = 0; u < 6; ++u) { av[u] = ap[(g + u) * 4]; bv[u] = bp[(g + u) * 4]; }
#pragma unroll
                for (int u = 0; u < 6; ++u)
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                 : "+d"(acc0), "+d"(acc1)
                                 : "d"(av[u]), "d"(bv[u]));
            }
            for (; g < ng; ++g) {
                const double av = ap[g * 4], bv = bp[g * 4];
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(acc0), "+d"(acc1)
                             : "d"(av), "d"(bv));
            }
            // row 0 of C: lane l < 4 holds the dots of columns 2 l and 2 l + 1.  They go to the warp's scratch and
            // lane x < NX walks the batch for point x, keeping that point's best in its own registers
            if (lane < 4) {
                dots[2 * lane] = acc0;
                dots[2 * lane + 1] = acc1;
            }
            __syncwarp();
            if (lane < NX) {
                const double *bx = a.bias + (size_t)lane * a.bias_stride;
                for (int cidx = 0; cidx < nb; ++cidx) {
                    const int kj = wq[base + cidx];
                    if (kj >= K) continue;
                    const double dj = dots[cidx];
                    const double t = dj + bx[kj];                               // the sweep's epilogue: acc + bias
                    // NaN and -Inf never win (subprob.jl:151-156); equal scores: the smaller index
                    if (t > best || (t == best && bidx >= 0 && kj < bidx)) { best = t; bdot = dj; bidx = kj; }
                }
            }
            evald += nb;
            __syncwarp();
        };
        // First batch of rows (group A), THEN the next scenario's head and row (group B): waiting for "all but the
        // last group" gives this scenario's row (group B of the previous iteration) and its first rows without
        // waiting for the next scenario's row, which has a trip to DRAM ahead of it.
        const int nb0 = full ? 0 : min(SCR_DEC_ROWS, n);
        stage(0, nb0);
        cp_async_commit();
        const long long inext = i + nw;
        if (inext < a.n_local) request(inext, par ^ 1, myn, mylb, e8);
        cp_async_commit();
        cp_async_wait_but_last();
        __syncwarp();
        if (!full) {
            if (nb0) chain(0, nb0);
            for (int base = SCR_DEC_ROWS; base < n; base += SCR_DEC_ROWS) {
                const int nb = min(SCR_DEC_ROWS, n - base);
                stage(base, nb);
                cp_async_wait_all();
                __syncwarp();
                chain(base, nb);
            }
        } else {
            for (long long k0 = 0; k0 < K; k0 += SCR_DEC_QUEUE) {    // a list overflowed: every vertex, for every point
                const int cnt = (int)min((long long)SCR_DEC_QUEUE, K - k0);
                for (int t = lane; t < cnt; t += 32) wq[t] = (int)(k0 + t);
                __syncwarp();
                for (int base = 0; base < cnt; base += SCR_DEC_ROWS) {
                    const int nb = min(SCR_DEC_ROWS, cnt - base);
                    stage(base, nb);
                    cp_async_wait_all();
                    __syncwarp();
                    chain(base, nb);
                }
            }
        }
        if (lane < NX) {
            a.best_val[lane * a.out_stride + i] = best;
            a.best_idx[lane * a.out_stride + i] = bidx;
            if (a.prev) {
                a.prev[2 * i + lane] = bidx + 1;
                a.prevdot[2 * i + lane] = bdot;                      // the winner's dot P_k . d_i, as the chain gave it
            }
        }
        __syncwarp();
    }
    if (lane == 0 && evald) atomicAdd(&a.ctl->n_eval, evald);
    griddep_launch();
}

// Device check behind the lanes variant: on `n` random chains of `ng` k-groups, is mma.sync.m8n8k4.f64 the DFMA
// chain in k order, bit for bit?  *bad counts the chains where it is not.
__global__ void k_dmma_is_fma_chain(int ng, int n, unsigned int *__restrict__ bad)
{
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n) return;
    double acc0 = 0.0, acc1 = 0.0, f = 0.0;
    for (int g = 0; g < ng; ++g) {
        double av[4], bv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {      // wide dynamic range and signs: cancellation, so that the order matters
            const unsigned long long q = ((unsigned long long)w * ng + g) * 4 + k;
            av[k] = (2.0 * u01(21, q) - 1.0) * exp2(floor(40.0 * u01(23, q)) - 20.0);
            bv[k] = (2.0 * u01(22, q) - 1.0) * exp2(floor(40.0 * u01(24, q)) - 20.0);
        }
        const double a1 = av[lane & 3], b1 = bv[lane & 3];
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(acc0), "+d"(acc1) : "d"(a1), "d"(b1));
#pragma unroll
        for (int k = 0; k < 4; ++k) f = fma(av[k], bv[k], f);
    }
    const bool differ = __double_as_longlong(acc0) != __double_as_longlong(f) || __double_as_longlong(acc1) != __double_as_longlong(f);
    if (__any_sync(0xffffffffu, differ) && lane == 0) atomicAdd(bad, 1u);
}

}  // namespace sqlp
