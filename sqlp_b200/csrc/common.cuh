// common.cuh -- shared device helpers for libsqlp_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define SQLP_TILE 128  // scenarios per tile == vertices per chunk
// kernel classes of the built-in profiler (sqlp_ctx_profile_classes)
#define SQLP_PROF_CONTRACT 0
#define SQLP_PROF_DELTA 1
#define SQLP_PROF_REDUCE 2
#define SQLP_PROF_POOL 3
#define SQLP_PROF_BIAS 4
#define SQLP_PROF_SCREEN 5     // tcgen05 screening kernel (work = executed bf16 flops)
#define SQLP_PROF_RESOLVE 6    // exact decision among the candidates (work = exact evaluations)
#define SQLP_PROF_FALLBACK 7   // the FP64 sweep launched behind a screening pass (runs only if that fell back)
#define SQLP_PROF_CLASSES 8
#define SQLP_BK 8      // stochastic rows per pipeline slab (s is padded to a multiple)

namespace sqlp {

// Programmatic dependent launch (sm_90+): every kernel of the library starts with griddep_sync().
// `launch_dependents` lets the NEXT kernel of the stream be scheduled while this one still runs (its
// blocks become resident as resources free up and stop at their own `wait`); `wait` returns once every
// prerequisite grid has COMPLETED and its writes are visible, so each kernel sees exactly what plain
// stream order would show it -- only launch latency and block scheduling overlap the predecessor's tail.
// Both instructions are no-ops for a kernel launched without the programmatic-serialization attribute.
__device__ __forceinline__ void griddep_sync()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
// The two halves apart, for kernels whose grid runs in SEVERAL WAVES: a dependent kernel whose blocks need a whole
// SM (the FP64 sweep, the screening kernel) would otherwise take the SMs the first wave frees and sit there at
// its own `wait`, leaving the remaining waves of this kernel fewer SMs (measured: 8 x on k_screen_resolve).  Such
// kernels wait at their start and release their dependents when a block has done its work.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Fragment-major tile layout shared by the scenario store D and the pool view PiS.
// A tile holds 128 columns (scenarios or vertices) x s_pad row slots.  Slots are grouped by
// four (one k-step of mma.sync.m8n8k4.f64) and columns by sixteen; inside a (group, column
// pair-block) cell the 64 doubles are ordered [lane][h] so that lane t of a warp finds the
// two m8n8k4 operand fragments of column blocks 2P and 2P+1,
//     element (column (2P + h) * 8 + t / 4, slot 4g + t % 4),
// in ONE 16-byte word at ((g * 8 + P) * 32 + t) * 2.  A warp-wide LDS.128 therefore reads
// 512 contiguous bytes (conflict free) and yields two fragments per lane, and every 8-slot
// pipeline slab of a tile is one contiguous 8 KB block in HBM.
__host__ __device__ __forceinline__ long long tile_off(int c, int j)
{
    return ((((long long)(j >> 2) * 8 + (c >> 4)) * 32 + (c & 7) * 4 + (j & 3)) << 1) + ((c >> 3) & 1);
}

// Base.round(x; base=2, sigdigits=16) -- call sites dual_set.jl:32-33,51 of the reference.
// hidigit = 1 + exponent(x); scale by 2^(16 - hidigit), round half to even, scale back.
// All scalings are exact powers of two, so this is bit-identical to the CPU oracle.
__device__ __forceinline__ double round_sig16(double x)
{
    if (!isfinite(x) || x == 0.0) return x;
    int digits = 15 - ilogb(x);
    double r;
    if (digits >= 0) {
        if (digits > 1023) return x;  // 2.0^digits = Inf -> NaN -> reference returns x
        r = scalbn(rint(scalbn(x, digits)), -digits);
    } else {
        r = scalbn(rint(scalbn(x, digits)), -digits);
    }
    if (!isfinite(r)) return (digits > 0) ? x : copysign(0.0, x);
    return r;
}

// splitmix64 counter generator of SURVEY.md 8(d); twin of orc_u01 in the oracle.
__host__ __device__ __forceinline__ double u01(uint64_t seed, uint64_t idx)
{
    uint64_t z = seed ^ (idx * 0x9E3779B97F4A7C15ULL);
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}

// The same stream mapped into the OPEN interval: (top 53 bits + 1/2) / 2^53 (exact in fp64 up to
// the final rounding), for inverse-CDF sampling of continuous distributions.
__host__ __device__ __forceinline__ double u01_open(uint64_t seed, uint64_t idx)
{
    uint64_t z = seed ^ (idx * 0x9E3779B97F4A7C15ULL);
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return ((double)(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

// Scenario g of an epigraph lives on rank (g / 128) % world, at this local slot.
__host__ __device__ __forceinline__ int owner_of(int64_t g, int world)
{
    return (int)((g / SQLP_TILE) % world);
}
__host__ __device__ __forceinline__ int64_t local_of(int64_t g, int world)
{
    return (g / ((int64_t)SQLP_TILE * world)) * SQLP_TILE + g % SQLP_TILE;
}
// Number of scenarios with ordinal < n_global held by `rank`.
__host__ __device__ __forceinline__ int64_t local_count(int64_t n_global, int rank, int world)
{
    int64_t full = n_global / SQLP_TILE;  // complete 128-blocks
    int64_t rem = n_global % SQLP_TILE;
    int64_t mine = full / world + ((full % world) > rank ? 1 : 0);
    int64_t n = mine * SQLP_TILE;
    if (rem && (int)(full % world) == rank) n += rem;
    return n;
}

// ---- control block of a screening pass (kernels_screen.cuh); the FP64 sweep reads it as its gate ----
struct ScreenCtl {                 // written by k_screen_prep, read by every kernel of the chain
    double shift[2];               // c_x: b32 = fp32(bias_x - c_x)
    float bmax[2];                 // largest |b32| over the live vertices of point x
    float coef_q;                  // E = coef_q * pnmax * dnmax + coef_b * bmax + tiny
    float coef_b;
    int bad;                       // 1: non-finite or out-of-range operands -- screening is skipped
    unsigned int overflow;         // candidate lists that overflowed (k_screen), reset by k_screen_prep
    unsigned int ovf_limit;        // more overflowed lists than this: the FP64 sweep runs instead
    unsigned long long n_emit;     // statistics: candidates emitted / evaluated exactly
    unsigned long long n_eval;
    int live[2];                   // vertices that may win at point x
    float eabs[2];                 // absolute part of E: FP64 roundings on the uncentred magnitudes (k_screen_prep)
    double cdb;                    // ctr . dbar (k_screen_seed)
};

__device__ __forceinline__ bool screen_falls_back(const ScreenCtl *ctl)
{
    return ctl->bad != 0 || ctl->overflow > ctl->ovf_limit;
}

// ---- mbarrier + bulk async copy (TMA 1-D, SASS UBLKCP) -------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}\n" ::"r"(bar),
                 "r"(bytes)
                 : "memory");
}
// Wait for the phase with the given parity to complete.  A pipeline bug would otherwise hang
// the GPU, so the spin is bounded (about two seconds) and traps instead.
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned done;
    long long t0 = 0;
    for (;;) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity), "r"(2000u)    // suspend-time hint (ns): the warp sleeps in the instruction, not in this loop --
            : "memory");                            // polls of the single-thread producer / MMA warps were 17 % of k_screen's instructions
        if (done) return;
        if (t0 == 0) t0 = clock64();
        else if (clock64() - t0 > 4000000000ll) __trap();
    }
}
// Non-blocking probe of the same condition.
__device__ __forceinline__ bool mbar_test(unsigned bar, unsigned parity)
{
    unsigned done;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
// global -> shared bulk copy; completion is signalled on `bar` as `bytes` of transaction.
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile(
        "cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}

}  // namespace sqlp
