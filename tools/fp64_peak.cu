// fp64_peak.cu -- measures the FP64 roofline denominators MEASURED_PEAKS.json lacks:
//   (a) register-resident DFMA chains, (b) mma.sync.m8n8k4.f64 (DMMA) chains,
//   (c) cuBLAS DGEMM 8192^3; each as a burst (best of 10) and sustained (~4 s back to back).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu -lcublas
// Prints one JSON object.
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int CHAINS = 16;     // independent accumulators per thread
constexpr int ITERS = 4096;

__global__ void __launch_bounds__(256) k_dfma(double *out, double a, double b)
{
    double acc[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 8x8 register tile, three distinct 64-bit sources per DFMA (the shape of a register-tiled
// GEMM inner loop): tests whether register-file bandwidth, not the FP64 pipe, is the limit.
template <bool SNAKE>
__global__ void __launch_bounds__(256) k_dfma_tile(double *out, const double *in)
{
    double a[8], b[8], acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = in[threadIdx.x % 32 + i]; b[i] = in[64 + threadIdx.x % 16 + i]; }
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = r + c;
    for (int it = 0; it < ITERS / 4; ++it) {
        if (SNAKE) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    const int c = (r & 1) ? 7 - cc : cc;
                    asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc[r][c]) : "d"(a[r]), "d"(b[c]));
                }
        } else {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
        }
    }
    double s = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) s += acc[r][c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

constexpr int MMA_CHAINS = 8;
__global__ void __launch_bounds__(256) k_dmma(double *out, double a, double b)
{
    double c0[MMA_CHAINS], c1[MMA_CHAINS];
#pragma unroll
    for (int i = 0; i < MMA_CHAINS; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < MMA_CHAINS; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < MMA_CHAINS; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 32 independent m8n8k4 accumulator blocks per warp (the contraction's warp tile), issued
// back to back: what the FP64 tensor pipe sustains at 1, 2 or 4 warps per sub-partition.
template <int THR>
__global__ void __launch_bounds__(THR) k_dmma_tile(double *out, const double *in)
{
    double c[32][2], a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x % 32 + i];
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = in[64 + threadIdx.x % 16 + i];
#pragma unroll
    for (int i = 0; i < 32; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; }
    for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[i >> 2]), "d"(b[i & 3]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// The same warp tile with its operand fragments re-read from shared memory for every k-group
// (4 + 2 LDS.128 per 32 DMMA, the contraction's inner loop without barriers or epilogue): the
// ceiling of a DMMA stream that is fed through the load/store unit.
template <int THR>
__global__ void __launch_bounds__(THR) k_dmma_tile_lds(double *out, const double *in)
{
    __shared__ __align__(16) double sA[4 * 256], sB[4 * 512];   // 4 k-groups of A (2 KB each) and B (4 KB each)
    for (int i = threadIdx.x; i < 4 * 256; i += THR) sA[i] = in[i % 96];
    for (int i = threadIdx.x; i < 4 * 512; i += THR) sB[i] = in[i % 96];
    __syncthreads();
    const int lane = threadIdx.x & 31, wx = (threadIdx.x >> 5) & 3;
    double c[32][2];
#pragma unroll
    for (int i = 0; i < 32; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; }
    for (int it = 0; it < ITERS / 4; it += 2) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int g = (it + u) & 3;
            const double *Ag = sA + g * 256 + lane * 2, *Bg = sB + g * 512 + wx * 128 + lane * 2;
            double2 av[4], bv[2];
#pragma unroll
            for (int q = 0; q < 4; ++q) av[q] = *reinterpret_cast<const double2 *>(Ag + q * 64);
#pragma unroll
            for (int q = 0; q < 2; ++q) bv[q] = *reinterpret_cast<const double2 *>(Bg + q * 64);
#pragma unroll
            for (int mi = 0; mi < 8; ++mi) {
                const double af = (mi & 1) ? av[mi >> 1].y : av[mi >> 1].x;
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    const double bf = (ni & 1) ? bv[ni >> 1].y : bv[ni >> 1].x;
                    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                                 : "+d"(c[mi * 4 + ni][0]), "+d"(c[mi * 4 + ni][1]) : "d"(af), "d"(bf));
                }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// How much DMMA throughput a sub-partition loses while its OTHER warp runs the contraction's
// epilogue instruction mix (DADD + DSETP + selects on 64 independent values per point):
// warps 0-3 issue DMMA chains, warps 4-7 (one per sub-partition) run MODE: 0 nothing, 1 DADD only,
// 2 DADD + DSETP + FSEL/SEL (running max with index).  Both sides loop for a fixed count; the
// kernel's time is the DMMA warps' time as long as the scalar side is shorter.
template <int MODE>
__global__ void __launch_bounds__(256) k_dmma_vs_scalar(double *out, const double *in, int scalar_iters)
{
    const int warp = threadIdx.x >> 5;
    if (warp < 4) {
        double c[32][2], a[8], b[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x % 32 + i];
#pragma unroll
        for (int i = 0; i < 4; ++i) b[i] = in[64 + threadIdx.x % 16 + i];
#pragma unroll
        for (int i = 0; i < 32; ++i) { c[i][0] = i; c[i][1] = threadIdx.x; }
        for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[i >> 2]), "d"(b[i & 3]));
        }
        double s = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) s += c[i][0] + c[i][1];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (MODE > 0) {
        double v[16], best[4];
        int bidx[4] = {0, 0, 0, 0};
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = in[threadIdx.x % 32 + i] + i;
#pragma unroll
        for (int i = 0; i < 4; ++i) best[i] = -1e300;
        for (int it = 0; it < scalar_iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                v[i] = v[i] + 1.0000001;                       // DADD
                if (MODE == 2 && v[i] > best[i & 3]) {         // DSETP + FSEL x2 + SEL
                    best[i & 3] = v[i];
                    bidx[i & 3] = it * 16 + i;
                }
                if (MODE == 3) {   // the same running max on order-preserving integer keys (no DSETP)
                    long long b = __double_as_longlong(v[i]);
                    long long key = b ^ ((b >> 63) & 0x7fffffffffffffffll);
                    long long bk = __double_as_longlong(best[i & 3]);   // best[] holds keys in this mode
                    if (key > bk) {
                        best[i & 3] = __longlong_as_double(key);
                        bidx[i & 3] = it * 16 + i;
                    }
                }
            }
        }
        double s = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += v[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) s += best[i] + bidx[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
}

// m16n8k16 f64 (sm_90+ shape): A 16x16 (8 regs/thread), B 16x8 (4 regs), C 16x8 (4 regs)
__global__ void __launch_bounds__(256) k_dmma16(double *out, double a, double b)
{
    double c[MMA_CHAINS][4];
#pragma unroll
    for (int i = 0; i < MMA_CHAINS; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) c[i][q] = threadIdx.x * 1e-9 + q;
    for (int it = 0; it < ITERS / 4; ++it) {
#pragma unroll
        for (int i = 0; i < MMA_CHAINS; ++i)
            asm volatile(
                "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
                "{%4,%4,%4,%4,%4,%4,%4,%4}, {%5,%5,%5,%5}, {%0,%1,%2,%3};\n"
                : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < MMA_CHAINS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
void measure(const char *name, double flops_per_launch, F launch, bool last = false)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < 10; ++i) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    int reps = (int)(4000.0f / best) + 1;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float tot; CK(cudaEventElapsedTime(&tot, e0, e1));
    printf(" \"%s\": {\"burst_tflops\": %.2f, \"sustained_tflops\": %.2f, \"burst_ms\": %.4f, \"sustained_s\": %.2f}%s\n",
           name, flops_per_launch / best * 1e-9, flops_per_launch * reps / tot * 1e-9, best, tot * 1e-3,
           last ? "" : ",");
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    int grid = sms * 8;   // 8 CTAs x 256 threads = 64 warps per SM
    double *out; CK(cudaMalloc(&out, (size_t)grid * 256 * 8));
    printf("{\n \"gpu\": \"%s\", \"sms\": %d, \"clock_mhz\": %d,\n", p.name, sms, p.clockRate / 1000);
    measure("dfma", 2.0 * CHAINS * ITERS * 256.0 * grid, [&] { k_dfma<<<grid, 256>>>(out, 1.0000001, 1e-9); });
    {
        double *in; CK(cudaMalloc(&in, 1024)); CK(cudaMemset(in, 0, 1024));
        int g1 = sms;   // 1 CTA x 8 warps per SM, like the contraction kernel
        measure("dfma_tile8x8_compiler_order_8warps", 2.0 * 64 * (ITERS / 4) * 256.0 * g1, [&] { k_dfma_tile<false><<<g1, 256>>>(out, in); });
        measure("dfma_tile8x8_snake_order_8warps", 2.0 * 64 * (ITERS / 4) * 256.0 * g1, [&] { k_dfma_tile<true><<<g1, 256>>>(out, in); });
    }
    {
        double *in; CK(cudaMalloc(&in, 1024)); CK(cudaMemset(in, 0, 1024));
        const double f = 2.0 * 256 * 32 * (ITERS / 4) * sms;
        measure("dmma_tile32_1warp_per_subpartition", f * 4, [&] { k_dmma_tile<128><<<sms, 128>>>(out, in); });
        measure("dmma_tile32_2warp_per_subpartition", f * 8, [&] { k_dmma_tile<256><<<sms, 256>>>(out, in); });
        CK(cudaFuncSetAttribute(k_dmma_tile_lds<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 0));
        measure("dmma_tile32_lds_fed_1warp_per_subpartition", f * 4, [&] { k_dmma_tile_lds<128><<<sms, 128>>>(out, in); });
        measure("dmma_tile32_lds_fed_2warp_per_subpartition", f * 8, [&] { k_dmma_tile_lds<256><<<sms, 256>>>(out, in); });
        measure("dmma_tile32_lds_fed_4warp_per_subpartition", f * 16, [&] { k_dmma_tile_lds<512><<<sms, 512>>>(out, in); });
    }
    {
        double *in; CK(cudaMalloc(&in, 1024)); CK(cudaMemset(in, 0, 1024));
        const double f = 2.0 * 256 * 32 * (ITERS / 4) * sms * 4;   // four DMMA warps per SM
        // scalar side: 16 values per iteration; the DMMA side runs 32 * ITERS/4 DMMAs = 16 * 8192 cycles
        for (int frac = 0; frac < 3; ++frac) {
            const int iters = frac == 0 ? 0 : (frac == 1 ? 512 : 1024);
            char name[96];
            snprintf(name, sizeof name, "dmma_1warp_plus_dadd_x%d_per_lane", iters * 16);
            if (iters) measure(name, f, [&] { k_dmma_vs_scalar<1><<<sms, 256>>>(out, in, iters); });
            snprintf(name, sizeof name, "dmma_1warp_plus_argmax_x%d_per_lane", iters * 16);
            if (iters) measure(name, f, [&] { k_dmma_vs_scalar<2><<<sms, 256>>>(out, in, iters); });
            snprintf(name, sizeof name, "dmma_1warp_plus_intkey_argmax_x%d_per_lane", iters * 16);
            if (iters) measure(name, f, [&] { k_dmma_vs_scalar<3><<<sms, 256>>>(out, in, iters); });
            if (!iters) measure("dmma_1warp_alone", f, [&] { k_dmma_vs_scalar<0><<<sms, 256>>>(out, in, 0); });
        }
    }
    measure("dmma_m8n8k4", 2.0 * 256 * MMA_CHAINS * ITERS * 8.0 * grid, [&] { k_dmma<<<grid, 256>>>(out, 1.0000001, 1e-9); });
    measure("dmma_m16n8k16", 2.0 * 16 * 8 * 16 * MMA_CHAINS * (ITERS / 4) * 8.0 * grid, [&] { k_dmma16<<<grid, 256>>>(out, 1.0000001, 1e-9); });
    CK(cudaGetLastError());
    {
        const int n = 8192;
        double *A, *B, *C;
        CK(cudaMalloc(&A, (size_t)n * n * 8)); CK(cudaMalloc(&B, (size_t)n * n * 8)); CK(cudaMalloc(&C, (size_t)n * n * 8));
        CK(cudaMemset(A, 0, (size_t)n * n * 8)); CK(cudaMemset(B, 0, (size_t)n * n * 8));
        cublasHandle_t h; cublasCreate(&h);
        double one = 1.0, zero = 0.0;
        measure("cublas_dgemm_8192", 2.0 * n * (double)n * n, [&] {
            cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, C, n); }, true);
    }
    printf("}\n");
    return 0;
}
