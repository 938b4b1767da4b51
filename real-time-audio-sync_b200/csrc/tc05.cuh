// tcgen05 / TMEM primitives for sm_100a as inline PTX (SASS: UTCHMMA / UTCBAR / LDTM / STTM).
// Bit layouts follow the PTX ISA "tcgen05" chapter; the field positions were cross-checked against
// the CuTe headers vendored in this image (cute/arch/mma_sm100_desc.hpp), no CuTe code is used.
#pragma once
#include <cstdint>

#include "afs_common.cuh"

namespace tc {

// ---- shared-memory matrix descriptor, K-major operand, 128-byte swizzle ------------------------
// Canonical layout (bf16): a tile is [rows][64 elements] = rows x 128 B, 1024-byte aligned; inside every
// group of 8 rows the 16-byte chunk c of row r sits at chunk (c ^ (r & 7)).  Groups of 8 rows follow
// each other at SBO = 1024 B.  One MMA consumes K = 16 elements = 32 B of every row: the K-step j inside
// the 64-element block is selected by adding j * 32 B to the start address.
__device__ __forceinline__ uint64_t smem_desc_k_sw128(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);          // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset = 1024 B, bits [32,46)
    d |= (uint64_t)1 << 46;                            // descriptor version 1 (sm_100), bits [46,48)
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B, bits [61,64)
    return d;
}

// byte offset of element (row r, k) inside a [rows][64] bf16 tile stored in the layout above
__device__ __host__ __forceinline__ uint32_t sw128_offset(int r, int k)
{
    return (uint32_t)(r * 128 + ((((k >> 3) ^ (r & 7)) & 7) << 4) + ((k & 7) << 1));
}

// ---- instruction descriptor: kind::f16, A and B bf16 K-major, D fp32 --------------------------
__host__ __device__ constexpr uint32_t idesc_bf16_f32(int m, int n)
{
    return (1u << 4)                    // D format f32
           | (1u << 7)                  // A format bf16
           | (1u << 10)                 // B format bf16
           | ((uint32_t)(n >> 3) << 17) // N / 8
           | ((uint32_t)(m >> 4) << 24);// M / 16
}

// ---- TMEM allocation (one warp, all 32 lanes) -----------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_result, uint32_t ncols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(afs::smem_addr(smem_result)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- MMA issue (one thread) ------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]; A: lane = row, 32-bit column c holds K elements 2c (low half) and 2c+1
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Warp-uniform variants: the whole (converged) warp executes the statement and one elected lane issues.  With uniform
// operands the compiler keeps descriptors in uniform registers and the issue loop is a handful of instructions per MMA;
// issuing from a divergent single lane instead costs an R2UR / ELECT / branch sequence per MMA (~100 cycles each).
__device__ __forceinline__ void mma_ss_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p, e;\n\tsetp.ne.b32 p, %4, 0;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit_elect(uint64_t *bar)
{
    asm volatile(
        "{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(afs::smem_addr(bar))
        : "memory");
}
// all MMAs issued so far by this thread arrive on the mbarrier when they have completed
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(afs::smem_addr(bar)) : "memory");
}

// ---- TMEM <-> registers: 32 lanes x 32 bit, N consecutive columns per thread -----------------
// taddr = (lane base of this warp's quarter << 16) | column; thread i of the warp accesses lane base + i
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&v)[4])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3])
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- bf16 two-term split by truncation: x = hi + lo + O(2^-16 |x|) ---------------------------
// The bf16 bit pattern of a truncated float is its upper half-word, so two values pack with one PRMT.
__device__ __forceinline__ uint32_t pack_hi16(float even, float odd)      // {bf16(odd) : bf16(even)}
{
    return __byte_perm(__float_as_uint(even), __float_as_uint(odd), 0x7632);
}
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t &hi, uint32_t &lo)
{
    const float ra = a - __uint_as_float(__float_as_uint(a) & 0xFFFF0000u);
    const float rb = b - __uint_as_float(__float_as_uint(b) & 0xFFFF0000u);
    hi = pack_hi16(a, b);
    lo = pack_hi16(ra, rb);
}

}  // namespace tc
