// tc_probe.cu -- stand-alone check of the screening pass (sqlp_b200/csrc/kernels_screen.cuh) on a B200:
//   1. the raw tcgen05 accumulators of one tile against a host emulation of the three bf16 products
//      (operand layout, shared-memory descriptors, instruction descriptor, TMEM addressing);
//   2. the screened argmax against the same resolve kernel sweeping EVERY vertex (bit for bit) and against a
//      plain host fp64 evaluation (to rounding);
//   3. timing of k_screen alone and of the chain at a storm-like shape.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/tc_probe tools/tc_probe.cu
// Run:   tools/tc_probe [K N n_rows mode]     mode 0 spread pool, 1 clustered pool (near ties)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>
#include "../sqlp_b200/csrc/kernels_screen.cuh"

using namespace sqlp;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

static double bf16_rn(double x)
{
    float f = (float)x;   // probe values are exactly representable after this for hi; used only in the emulation
    uint32_t b;
    memcpy(&b, &f, 4);
    b = (b + 0x7FFFu + ((b >> 16) & 1u)) & 0xFFFF0000u;
    memcpy(&f, &b, 4);
    return (double)f;
}
static double bf16_from_double(double x)
{
    // round-to-nearest-even of the fp64 value to 8 significant bits (no double rounding)
    if (x == 0.0 || !std::isfinite(x)) return x;
    int e;
    double m = frexp(x, &e);          // x = m 2^e, 0.5 <= |m| < 1
    double s = ldexp(m, 8);           // 128 <= |s| < 256
    double r = nearbyint(s);
    return ldexp(r, e - 8);
}

// Is mma.sync.m8n8k4.f64 the chain d = fma(a3, b3, fma(a2, b2, fma(a1, b1, fma(a0, b0, c))))?  One warp per sample:
// a chain of `ng` k-groups on random operands, element (0, 0) against the FMA chain of lane 0's view of the data.
__global__ void k_dmma_vs_fma(const double *__restrict__ A, const double *__restrict__ B, int ng, int nsamp,
                              unsigned long long *__restrict__ mismatches, double *__restrict__ worst)
{
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= nsamp) return;
    const double *a = A + (size_t)w * ng * 4, *b = B + (size_t)w * ng * 4;
    double acc0 = 0.0, acc1 = 0.0, f = 0.0;
    for (int g = 0; g < ng; ++g) {
        // row r of A = a[4g + k] for every r; column c of B = b[4g + k] for every c  ->  every D element = same dot
        const double av = a[4 * g + (lane & 3)], bv = b[4 * g + (lane & 3)];
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(acc0), "+d"(acc1) : "d"(av), "d"(bv));
        for (int k = 0; k < 4; ++k) f = fma(a[4 * g + k], b[4 * g + k], f);
    }
    if (lane == 0 && __double_as_longlong(acc0) != __double_as_longlong(f)) {
        atomicAdd(mismatches, 1ull);
        *worst = fabs(acc0 - f);
    }
    if (lane == 5 && __double_as_longlong(acc1) != __double_as_longlong(f)) atomicAdd(mismatches + 1, 1ull);
}

int main(int argc, char **argv)
{
    long long K = argc > 1 ? atoll(argv[1]) : 4096;
    long long N = argc > 2 ? atoll(argv[2]) : 20000;
    int n_rows = argc > 3 ? atoi(argv[3]) : 117;
    int mode = argc > 4 ? atoi(argv[4]) : 0;
    int desc_mode = argc > 5 ? atoi(argv[5]) : 0;
    const int NX = 2;
    const int s_pad = (n_rows + 7) / 8 * 8, sp = (n_rows + 15) / 16 * 16, J = sp / 16;
    const long long ntiles = (N + 127) / 128, nunits = ntiles, npad = nunits * 128;
    const long long nch = (K + 255) / 256, kpad = (K + 127) / 128 * 128;
    printf("probe: K=%lld N=%lld n_rows=%d s_pad=%d sp=%d mode=%d desc_mode=%d\n", K, N, n_rows, s_pad, sp, mode, desc_mode);
    (void)bf16_rn;

    // ---- host data ------------------------------------------------------------------------
    std::vector<double> pi((size_t)K * n_rows), d((size_t)N * n_rows), bias((size_t)NX * kpad, -INFINITY);
    std::vector<double> rbar(n_rows);
    for (int j = 0; j < n_rows; ++j) rbar[j] = 100.0 + 400.0 * u01(6, j);
    for (long long k = 0; k < K; ++k)
        for (int j = 0; j < n_rows; ++j) {
            if (mode == 0) pi[k * n_rows + j] = 1000.0 * (2.0 * u01(2, k * n_rows + j) - 1.0);
            else {          // 64 base vertices, perturbed copies 1e-6 apart: near ties everywhere
                long long b = k % 64;
                pi[k * n_rows + j] = 1000.0 * (2.0 * u01(2, b * n_rows + j) - 1.0) * (1.0 + 1e-6 * (2.0 * u01(9, k * n_rows + j) - 1.0));
            }
        }
    for (long long i = 0; i < N; ++i)
        for (int j = 0; j < n_rows; ++j) d[i * n_rows + j] = rbar[j] * (0.1 * floor(5.0 * u01(1, i * n_rows + j)) - 0.2);
    for (int x = 0; x < NX; ++x)
        for (long long k = 0; k < K; ++k) {
            double b = 0.0;
            for (int j = 0; j < n_rows; ++j) b += pi[k * n_rows + j] * rbar[j] * (1.0 + 0.01 * x);
            bias[x * kpad + k] = b + 3.0e4 * (2.0 * u01(11 + x, k) - 1.0);
        }
    std::vector<double> Dt((size_t)ntiles * s_pad * 128, 0.0), PiS((size_t)(kpad / 128) * s_pad * 128, 0.0);
    for (long long i = 0; i < N; ++i)
        for (int j = 0; j < n_rows; ++j) Dt[(i >> 7) * (size_t)s_pad * 128 + tile_off((int)(i & 127), j)] = d[i * n_rows + j];
    for (long long k = 0; k < K; ++k)
        for (int j = 0; j < n_rows; ++j) PiS[(k >> 7) * (size_t)s_pad * 128 + tile_off((int)(k & 127), j)] = pi[k * n_rows + j];
    std::vector<double> PiR((size_t)kpad * s_pad, 0.0);
    for (long long k = 0; k < K; ++k)
        for (int j = 0; j < n_rows; ++j) PiR[(size_t)k * s_pad + j] = pi[k * n_rows + j];
    std::vector<int> srows(n_rows);
    for (int j = 0; j < n_rows; ++j) srows[j] = j;

    // ---- device buffers ---------------------------------------------------------------------
    double *d_pi, *d_D, *d_PiS, *d_bias, *d_bv, *d_bv2;
    int *d_srows, *d_bi, *d_bi2, *d_bad, *d_cnt;
    long long *d_K;
    __nv_bfloat16 *d_PiB, *d_DB;
    float *d_pn, *d_pnmax, *d_dnu, *d_dnall, *d_b32c, *d_lfin, *d_dbg;
    int2 *d_cand;
    ScreenCtl *d_ctl;
    const int R = 1;
    CK(cudaMalloc(&d_pi, pi.size() * 8)); CK(cudaMemcpy(d_pi, pi.data(), pi.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_D, Dt.size() * 8)); CK(cudaMemcpy(d_D, Dt.data(), Dt.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_PiS, PiS.size() * 8)); CK(cudaMemcpy(d_PiS, PiS.data(), PiS.size() * 8, cudaMemcpyHostToDevice));
    double *d_PiR;
    CK(cudaMalloc(&d_PiR, PiR.size() * 8)); CK(cudaMemcpy(d_PiR, PiR.data(), PiR.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_bias, bias.size() * 8)); CK(cudaMemcpy(d_bias, bias.data(), bias.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_srows, n_rows * 4)); CK(cudaMemcpy(d_srows, srows.data(), n_rows * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_K, 8)); CK(cudaMemcpy(d_K, &K, 8, cudaMemcpyHostToDevice));
    const size_t pib_bytes = (size_t)nch * J * SCR_STAGE_BYTES, db_bytes = (size_t)nunits * 512 * sp;
    CK(cudaMalloc(&d_PiB, pib_bytes)); CK(cudaMemset(d_PiB, 0, pib_bytes));
    CK(cudaMalloc(&d_DB, db_bytes)); CK(cudaMemset(d_DB, 0, db_bytes));
    CK(cudaMalloc(&d_pn, nch * 256 * 4)); CK(cudaMemset(d_pn, 0, nch * 256 * 4));
    CK(cudaMalloc(&d_pnmax, nch * 4)); CK(cudaMemset(d_pnmax, 0, nch * 4));
    CK(cudaMalloc(&d_dnu, nunits * 4)); CK(cudaMemset(d_dnu, 0, nunits * 4));
    CK(cudaMalloc(&d_dnall, 4)); CK(cudaMemset(d_dnall, 0, 4));
    CK(cudaMalloc(&d_bad, 8)); CK(cudaMemset(d_bad, 0, 8));
    CK(cudaMalloc(&d_b32c, (size_t)nch * scr_bias_floats<NX>() * 4));
    CK(cudaMalloc(&d_ctl, sizeof(ScreenCtl))); CK(cudaMemset(d_ctl, 0, sizeof(ScreenCtl)));
    const size_t nslots = (size_t)NX * R * 2 * npad;
    CK(cudaMalloc(&d_cand, nslots * SCR_CAP * 8));
    CK(cudaMalloc(&d_cnt, nslots * 4)); CK(cudaMemset(d_cnt, 0, nslots * 4));
    CK(cudaMalloc(&d_lfin, nslots * 4)); CK(cudaMemset(d_lfin, 0, nslots * 4));
    CK(cudaMalloc(&d_dbg, 128 * 256 * 4)); CK(cudaMemset(d_dbg, 0, 128 * 256 * 4));
    CK(cudaMalloc(&d_bv, (size_t)NX * npad * 8)); CK(cudaMalloc(&d_bi, (size_t)NX * npad * 4));
    CK(cudaMalloc(&d_bv2, (size_t)NX * npad * 8)); CK(cudaMalloc(&d_bi2, (size_t)NX * npad * 4));

    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, smem optin %zu\n", prop.name, sms, (size_t)prop.sharedMemPerBlockOptin);

    // ---- operand builders + prep ---------------------------------------------------------------
    long long *d_mark;
    CK(cudaMalloc(&d_mark, 8)); CK(cudaMemset(d_mark, 0, 8));
    k_screen_view_sync<<<2 * sms, 256>>>(d_pi, n_rows, d_srows, n_rows, sp, d_PiB, d_pn, d_pnmax, d_bad, d_mark, d_K, nullptr, nullptr);
    k_screen_scen_sync<<<8 * sms, 256>>>(d_D, s_pad, sp, d_DB, d_dnu, d_dnall, d_bad + 1, 0, N, nullptr, nullptr);
    k_screen_prep<NX><<<1, 1024>>>(d_bias, kpad, d_pn, d_pnmax, d_dnall, d_bad, d_bad + 1, d_K, sp, (unsigned)(N / 64 + 16),
                                   d_b32c, d_ctl, nullptr, nullptr, nullptr);
    CK(cudaDeviceSynchronize());
    ScreenCtl ctl;
    CK(cudaMemcpy(&ctl, d_ctl, sizeof ctl, cudaMemcpyDeviceToHost));
    printf("prep: bad=%d coef_q=%.4g coef_b=%.4g shift=(%.6g, %.6g) bmax=(%.4g, %.4g) live=(%d, %d) of %lld\n", ctl.bad, ctl.coef_q,
           ctl.coef_b, ctl.shift[0], ctl.shift[1], ctl.bmax[0], ctl.bmax[1], ctl.live[0], ctl.live[1], K);

    // ---- the screening kernel --------------------------------------------------------------------
    int nstages = SCR_MAX_STAGES;
    size_t smem = scr_smem_bytes(sp, nstages, NX);
    while (smem > prop.sharedMemPerBlockOptin && nstages > 2) smem = scr_smem_bytes(sp, --nstages, NX);
    CK(cudaFuncSetAttribute(k_screen<NX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ScreenArgs sa;
    sa.DB = d_DB; sa.PiB = d_PiB; sa.b32c = d_b32c; sa.dnmax_unit = d_dnu; sa.ctl = d_ctl; sa.d_K = d_K;
    sa.sp = sp; sa.nunits = (int)nunits; sa.R = R; sa.nstages = nstages; sa.n_local = N; sa.npad = npad;
    sa.cand = d_cand; sa.cnt = d_cnt; sa.lfin = d_lfin; sa.dbg = d_dbg; sa.desc_mode = desc_mode; sa.lseed = nullptr; sa.dn = nullptr;
    const int grid = (int)std::min<long long>(sms, nunits * R);
    printf("k_screen: grid %d, %d threads, %zu B smem, %d stages\n", grid, SCR_THREADS, smem, nstages);
    k_screen<NX><<<grid, SCR_THREADS, smem>>>(sa);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&ctl, d_ctl, sizeof ctl, cudaMemcpyDeviceToHost));
    printf("k_screen done: emitted %llu (%.2f per scenario-point), overflowed lists %u (limit %u)\n", ctl.n_emit,
           (double)ctl.n_emit / (double)(N * NX), ctl.overflow, ctl.ovf_limit);

    // ---- check 1: raw accumulators of tile (unit 0, chunk 0) ------------------------------------
    {
        std::vector<float> dbg(128 * 256);
        CK(cudaMemcpy(dbg.data(), d_dbg, dbg.size() * 4, cudaMemcpyDeviceToHost));
        double worst = 0.0, worst_rel = 0.0;
        int bad = 0;
        for (int r = 0; r < 128 && r < N; ++r)
            for (int c = 0; c < 256 && c < K; ++c) {
                double acc = 0.0, mag = 0.0;
                for (int j = 0; j < n_rows; ++j) {
                    double p = pi[(size_t)c * n_rows + j], q = d[(size_t)r * n_rows + j];
                    double ph = bf16_from_double(p), pl = bf16_from_double(p - ph);
                    double dh = bf16_from_double(q), dl = bf16_from_double(q - dh);
                    acc += dh * ph + dl * ph + dh * pl;
                    mag += fabs(p * q);
                }
                double err = fabs((double)dbg[r * 256 + c] - acc);
                worst = std::max(worst, err);
                worst_rel = std::max(worst_rel, err / (mag + 1e-300));
                if (err > 2e-5 * mag + 1e-20) ++bad;
            }
        printf("check 1 (tcgen05 tile vs host emulation): max abs err %.4g, max err / sum|p d| %.4g, %d of %d beyond 2e-5 -> %s\n",
               worst, worst_rel, bad, 128 * 256, bad ? "FAIL" : "ok");
        printf("   sample dbg[0][0..3] = %.6g %.6g %.6g %.6g ; dbg[1][0] = %.6g ; dbg[0][255] = %.6g\n", dbg[0], dbg[1], dbg[2], dbg[3],
               dbg[256], dbg[255]);
    }

    // ---- resolve: screened and full ------------------------------------------------------------
    ResolveArgs ra;
    ra.D = d_D; ra.PiS = d_PiS; ra.PiR = d_PiR; ra.bias = d_bias; ra.bias_stride = kpad; ra.s_pad = s_pad; ra.d_K = d_K;
    ra.n_local = N; ra.npad = npad; ra.R = R; ra.cand = d_cand; ra.cnt = d_cnt; ra.lfin = d_lfin;
    ra.best_val = d_bv; ra.best_idx = d_bi; ra.out_stride = npad; ra.ctl = d_ctl; ra.force_full = 0; ra.prev = nullptr; ra.prevdot = nullptr; ra.DR = nullptr;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaMemset(d_bi, 0xff, (size_t)NX * npad * 4));
    CK(cudaEventRecord(e0));
    k_screen_resolve<NX, 0><<<8 * sms, 256>>>(ra);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms_res = 0;
    CK(cudaEventElapsedTime(&ms_res, e0, e1));
    {   // the lanes variant (DFMA chains, 32 candidates per pass) must give the same bits
        std::vector<double> v0((size_t)NX * npad), v1((size_t)NX * npad);
        std::vector<int> i0((size_t)NX * npad), i1((size_t)NX * npad);
        CK(cudaMemcpy(v0.data(), d_bv, v0.size() * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(i0.data(), d_bi, i0.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemset(d_bi, 0xff, (size_t)NX * npad * 4));
        CK(cudaEventRecord(e0));
        k_screen_resolve<NX, 1><<<8 * sms, 256>>>(ra);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms1 = 0;
        CK(cudaEventElapsedTime(&ms1, e0, e1));
        CK(cudaMemcpy(v1.data(), d_bv, v1.size() * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(i1.data(), d_bi, i1.size() * 4, cudaMemcpyDeviceToHost));
        long long mm = 0;
        for (int x = 0; x < NX; ++x)
            for (long long i = 0; i < N; ++i)
                if (i0[x * npad + i] != i1[x * npad + i] || memcmp(&v0[x * npad + i], &v1[x * npad + i], 8)) ++mm;
        printf("check 5 (DFMA-lanes resolve == DMMA resolve, %lld scenarios x %d points): %lld mismatches; %.3f ms vs %.3f ms\n", N, NX, mm, ms1, ms_res);
    }
    CK(cudaMemcpy(&ctl, d_ctl, sizeof ctl, cudaMemcpyDeviceToHost));
    printf("resolve: %.3f ms, %llu exact evaluations (%.2f per scenario-point)\n", ms_res, ctl.n_eval, (double)ctl.n_eval / (double)(N * NX));
    const long long Nfull = std::min<long long>(N, 4096);      // the full sweep with this kernel is slow: a prefix
    ra.n_local = Nfull; ra.best_val = d_bv2; ra.best_idx = d_bi2; ra.force_full = 1;
    k_screen_resolve<NX, 1><<<8 * sms, 256>>>(ra);
    CK(cudaDeviceSynchronize());
    std::vector<double> bv((size_t)NX * npad), bv2((size_t)NX * npad);
    std::vector<int> bi((size_t)NX * npad), bi2((size_t)NX * npad);
    CK(cudaMemcpy(bv.data(), d_bv, bv.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(bi.data(), d_bi, bi.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(bv2.data(), d_bv2, bv2.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(bi2.data(), d_bi2, bi2.size() * 4, cudaMemcpyDeviceToHost));
    {
        long long mism = 0;
        for (int x = 0; x < NX; ++x)
            for (long long i = 0; i < Nfull; ++i)
                if (bi[x * npad + i] != bi2[x * npad + i] || memcmp(&bv[x * npad + i], &bv2[x * npad + i], 8)) {
                    if (mism < 5) printf("   mismatch x=%d i=%lld: screened (%d, %.17g) full (%d, %.17g)\n", x, i, bi[x * npad + i], bv[x * npad + i], bi2[x * npad + i], bv2[x * npad + i]);
                    ++mism;
                }
        printf("check 2 (screened == full sweep, bit for bit, %lld scenarios x %d points): %lld mismatches -> %s\n", Nfull, NX, mism, mism ? "FAIL" : "ok");
        long long hm = 0;
        const long long Nh = std::min<long long>(Nfull, 64);
        for (int x = 0; x < NX; ++x)
            for (long long i = 0; i < Nh; ++i) {
                double best = -INFINITY;
                for (long long k = 0; k < K; ++k) {
                    double s = 0.0;
                    for (int j = 0; j < n_rows; ++j) s += pi[k * n_rows + j] * d[i * n_rows + j];
                    s += bias[x * kpad + k];
                    if (s > best) best = s;
                }
                if (fabs(best - bv2[x * npad + i]) > 1e-9 * (fabs(best) + 1.0)) ++hm;
            }
        printf("check 3 (full sweep value vs host fp64, %lld scenarios): %lld beyond 1e-9 -> %s\n", Nh, hm, hm ? "FAIL" : "ok");
    }

    // ---- is DMMA a chain of FMAs in k order? ------------------------------------------------------
    {
        const int ng = s_pad / 4, nsamp = 1 << 18;
        std::vector<double> ha((size_t)nsamp * ng * 4), hb((size_t)nsamp * ng * 4);
        for (size_t q = 0; q < ha.size(); ++q) {
            const double ua = u01(21, q), ub = u01(22, q);
            ha[q] = (2.0 * ua - 1.0) * exp2(floor(40.0 * u01(23, q)) - 20.0);      // wide dynamic range, cancellation
            hb[q] = (2.0 * ub - 1.0) * exp2(floor(40.0 * u01(24, q)) - 20.0);
        }
        double *dA, *dB, *dw;
        unsigned long long *dm;
        CK(cudaMalloc(&dA, ha.size() * 8)); CK(cudaMalloc(&dB, hb.size() * 8)); CK(cudaMalloc(&dm, 16)); CK(cudaMalloc(&dw, 8));
        CK(cudaMemcpy(dA, ha.data(), ha.size() * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, hb.data(), hb.size() * 8, cudaMemcpyHostToDevice));
        CK(cudaMemset(dm, 0, 16)); CK(cudaMemset(dw, 0, 8));
        k_dmma_vs_fma<<<nsamp / 8, 256>>>(dA, dB, ng, nsamp, dm, dw);
        CK(cudaDeviceSynchronize());
        unsigned long long mm[2]; double ww;
        CK(cudaMemcpy(mm, dm, 16, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&ww, dw, 8, cudaMemcpyDeviceToHost));
        printf("check 4 (DMMA m8n8k4 chain == FMA chain in k order, %d samples x %d k-groups): %llu / %llu mismatches (last |diff| %.3g)\n",
               nsamp, ng, mm[0], mm[1], ww);
    }

    // ---- timing ----------------------------------------------------------------------------------
    sa.dbg = nullptr;
    for (int rep = 0; rep < 3; ++rep) {
        k_screen_prep<NX><<<1, 1024>>>(d_bias, kpad, d_pn, d_pnmax, d_dnall, d_bad, d_bad + 1, d_K, sp, (unsigned)(N / 64 + 16), d_b32c, d_ctl, nullptr, nullptr, nullptr);
        CK(cudaEventRecord(e0));
        k_screen<NX><<<grid, SCR_THREADS, smem>>>(sa);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 3.0 * sp * (double)(nch * 256) * (double)(nunits * 128);
        printf("k_screen rep %d: %.3f ms, %.1f TFLOP/s executed bf16 (3 products, padded), %.3g evals/s for %d points\n", rep, ms,
               flops / (ms * 1e-3) * 1e-12, (double)NX * K * N / (ms * 1e-3), NX);
    }
    printf("probe done\n");
    return 0;
}
