// hmma_peak.cu -- what the LEGACY tensor path (mma.sync.m16n8k16 bf16, fp32 accumulate) sustains on this GPU:
// a register-resident loop of independent accumulation chains, no memory traffic.  One number for DESIGN.md
// section 10 (is a screening pass worth writing with mma.sync, or only with tcgen05?).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/hmma_peak tools/hmma_peak.cu && tools/hmma_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void __launch_bounds__(256) k_hmma(float *out, int iters, unsigned seed)
{
    unsigned a[4] = {seed + threadIdx.x, seed * 3u + 1u, seed * 5u + 2u, seed * 7u + 3u};
    unsigned b[2] = {seed * 11u + threadIdx.x, seed * 13u + 5u};
    float c[CHAINS][4];
#pragma unroll
    for (int q = 0; q < CHAINS; ++q) c[q][0] = c[q][1] = c[q][2] = c[q][3] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < CHAINS; ++q)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[q][0]), "+f"(c[q][1]), "+f"(c[q][2]), "+f"(c[q][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < CHAINS; ++q) s += c[q][0] + c[q][1] + c[q][2] + c[q][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
void run(int blocks_per_sm, int sms, float *out)
{
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_hmma<CHAINS><<<sms * blocks_per_sm, 256>>>(out, 100, 1u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_hmma<CHAINS><<<sms * blocks_per_sm, 256>>>(out, iters, 1u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 16 * 8 * 16 * (double)CHAINS * iters * 8.0 * sms * blocks_per_sm;   // 8 warps per block
    printf("{\"kernel\": \"mma.sync.m16n8k16.bf16\", \"chains_per_warp\": %d, \"blocks_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.1f}\n",
           CHAINS, blocks_per_sm, ms, flops / (ms * 1e-3) * 1e-12);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    float *out;
    cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    run<4>(1, p.multiProcessorCount, out);
    run<8>(1, p.multiProcessorCount, out);
    run<8>(2, p.multiProcessorCount, out);
    run<16>(2, p.multiProcessorCount, out);
    run<8>(4, p.multiProcessorCount, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
