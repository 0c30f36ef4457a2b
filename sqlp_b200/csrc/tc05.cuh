// tc05.cuh -- thin PTX wrappers for the 5th-generation tensor cores of sm_100a: tensor memory (TMEM)
// allocation, tcgen05.mma with shared-memory operand descriptors, tcgen05.commit onto an mbarrier,
// tcgen05.ld for the epilogue, and the fences that order them against ordinary thread synchronisation.
//
// Operand layout used throughout (kernels_screen.cuh): K-major, no swizzle ("interleaved" canonical
// layout).  A [rows x K] bf16 operand is stored as K / 8 SLABS; a slab holds, for every row, the 8
// consecutive K-elements (16 bytes) of that row, rows contiguous:
//     byte offset of element (r, k) = (k / 8) * rows * 16  +  r * 16  +  (k % 8) * 2 .
// A core matrix of the hardware (8 rows x 16 bytes) is therefore 128 contiguous bytes; the descriptor's
// stride-dimension byte offset (SBO, between 8-row groups) is 128 and its leading-dimension byte offset
// (LBO, between the two 16-byte K-halves of one K = 16 instruction) is rows * 16.  Such a layout needs no
// tensor map: a pipeline stage is one contiguous block of global memory, moved by a 1-D bulk copy.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace sqlp {
namespace tc05 {

// ---- shared-memory matrix descriptor (64 bit) ---------------------------------------------
//   [ 0,14) start address >> 4      [16,30) LBO >> 4      [32,46) SBO >> 4
//   [46,48) version = 1 (sm_100)    [49,52) base offset 0 [52] LBO mode 0     [61,64) swizzle 0 = none
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// ---- instruction descriptor (32 bit) for kind::f16, bf16 x bf16 -> fp32, both operands K-major -----
//   [4,6) D format 1 = f32   [7,10) A format 1 = bf16   [10,13) B format 1 = bf16
//   [15] A major 0 = K   [16] B major 0 = K   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t idesc_bf16_f32(int M, int N)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- TMEM allocation (one full warp executes these) -----------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_result_addr),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void fence_before_sync()
{
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
__device__ __forceinline__ void fence_after_sync()
{
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.  accumulate = 0 overwrites D.
__device__ __forceinline__ void mma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate)
{
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// The mbarrier receives one arrival when every tcgen05 operation this thread issued so far has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void commit(uint32_t mbar_saddr)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(mbar_saddr)
                 : "memory");
}

// 32 lanes x 32 columns of 32-bit: thread t of the warp receives columns [col, col + 32) of TMEM lane
// (lane base of the warp's quarter) + t.  The warp may only address lanes 32 * (warp % 4) ... + 31.
__device__ __forceinline__ void ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void wait_ld()
{
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

}  // namespace tc05
}  // namespace sqlp
