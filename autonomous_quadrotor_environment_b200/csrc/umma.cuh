// umma.cuh — minimal hand-written tcgen05 (5th-gen tensor core) building blocks for sm_100a:
// TMEM allocation, shared-memory matrix descriptors (K-major, no swizzle), instruction descriptor for
// kind::f16 (BF16 x BF16 -> FP32), single-thread MMA issue, commit to an mbarrier and TMEM -> register loads.
//
// Canonical no-swizzle K-major operand layout (PTX ISA "shared memory matrix layout"; one "core matrix" is
// 8 rows x 16 bytes = 8x8 bf16, stored contiguously in 128 bytes):
//     byte_offset(row r, col k) = (r / 8) * SBO + (k / 8) * LBO + (r % 8) * 16 + (k % 8) * 2
// We use LBO = 128 (core matrices adjacent along K are contiguous) and SBO = (K / 8) * 128.
// One tcgen05.mma of kind::f16 consumes K = 16 (two core matrices along K): advance the start address by 2*LBO.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include "quad_device.cuh"

namespace qs {

__device__ __forceinline__ uint32_t umma_canon_offset(int r, int k, int K) {   // bytes, bf16, K-major, no swizzle
    return (uint32_t)((r >> 3) * ((K >> 3) * 128) + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

// 64-bit shared-memory matrix descriptor (sm_100 format): start addr [0,14) >>4, LBO [16,30) >>4, SBO [32,46) >>4,
// version [46,48) = 1, layout type [61,64) = 0 (SWIZZLE_NONE).
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// 32-bit instruction descriptor for kind::f16: D = F32 (bits [4,6) = 1), A = B = BF16 ([7,10) = [10,13) = 1),
// A and B K-major (bits 15,16 = 0), N >> 3 at [17,23), M >> 4 at [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_bf16_f32(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// The same with the "major" bits: a_mn / b_mn = 1 declares the operand MN-major, i.e. stored with the M (or N) index contiguous and
// K as the row index.  The no-swizzle MN-major canonical layout has the same 128-byte core block (8 K-rows x 16 bytes = 8 MN elements);
// in its descriptor SBO is the stride between 8-element MN groups and LBO the stride between 8-row K groups.  A tile written as
// [row r][col c] with umma_canon_offset(r, c, C) is therefore BOTH a K-major operand with K = c (LBO 128, SBO C/8*128) and an MN-major
// operand with K = r (LBO C/8*128, SBO 128): the weight-gradient products X^T dZ take their operands from the forward pass's own
// tiles without any transpose.
__host__ __device__ constexpr uint32_t umma_idesc_bf16_f32_major(int M, int N, bool a_mn, bool b_mn) {
    return umma_idesc_bf16_f32(M, N) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16);
}

// ---- TMEM ------------------------------------------------------------------------------------------------------
// executed by ONE full warp; writes the TMEM base address to *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- MMA issue (one thread) ---------------------------------------------------------------------------------------
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// A operand from TENSOR MEMORY (the "TS" form): A[128 x 16] BF16 sits in 8 consecutive 32-bit columns starting at tmem_a,
// lane = row, column c holds the K elements (2c, 2c+1) packed lo/hi — what a thread that owns row r produces with
// tcgen05.st of its packed activations.  K-major only.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// registers -> TMEM: 16 consecutive 32-bit columns of this thread's lane (thread = lane = row)
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t r[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: 32 lanes x 32 consecutive 32-bit columns (thread = lane = row of D) ----------------------
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, float v[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, float v[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float tanh_fast(float x) {              // MUFU.TANH, max rel. error 2^-11
    float r;
    asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Issue D[128 x N] (+)= A[128 x K] * B[N x K]^T as K/16 MMAs; A, B in the canonical layout described above with their
// own K extents (KA, KB give the SBO of each operand); a_k0 / b_k0 = first K column of each operand to use.
__device__ __forceinline__ void umma_gemm_k(uint32_t tmem_d, uint32_t a_smem, int KA, int a_k0, uint32_t b_smem, int KB,
                                            int b_k0, int k_len, int N, bool accumulate_first) {
    const uint32_t idesc = umma_idesc_bf16_f32(128, N);
    for (int k = 0; k < k_len; k += 16) {
        const uint64_t da = umma_smem_desc(a_smem + (uint32_t)((a_k0 + k) >> 3) * 128u, 128u, (uint32_t)(KA >> 3) * 128u);
        const uint64_t db = umma_smem_desc(b_smem + (uint32_t)((b_k0 + k) >> 3) * 128u, 128u, (uint32_t)(KB >> 3) * 128u);
        umma_bf16(tmem_d, da, db, idesc, accumulate_first || k > 0);
    }
}

// D[128 x N] (+)= A^T B with BOTH operands MN-major over the same K = 128 rows: A stored [128 k][a_ext] (columns a_c0 .. a_c0+127 are
// the M index), B stored [128 k][b_ext] (columns b_c0 .. b_c0+N-1 are the N index), canonical [row][col] layout.  8 MMAs of K = 16.
__device__ __forceinline__ void umma_gemm_mn(uint32_t tmem_d, uint32_t a_smem, int a_ext, int a_c0, uint32_t b_smem, int b_ext, int b_c0,
                                             int N, bool accumulate_first) {
    const uint32_t idesc = umma_idesc_bf16_f32_major(128, N, true, true);
    const uint32_t lbo_a = (uint32_t)(a_ext >> 3) * 128u, lbo_b = (uint32_t)(b_ext >> 3) * 128u;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint64_t da = umma_smem_desc(a_smem + (uint32_t)(a_c0 >> 3) * 128u + (uint32_t)(2 * k) * lbo_a, lbo_a, 128u);
        const uint64_t db = umma_smem_desc(b_smem + (uint32_t)(b_c0 >> 3) * 128u + (uint32_t)(2 * k) * lbo_b, lbo_b, 128u);
        umma_bf16(tmem_d, da, db, idesc, accumulate_first || k > 0);
    }
}
// D[128 x N] (+)= A B with A K-major [128 m][K = k_len] (extent a_ext) and B MN-major, stored [k_len k][b_ext] (columns b_c0.. = N index)
__device__ __forceinline__ void umma_gemm_k_mn(uint32_t tmem_d, uint32_t a_smem, int a_ext, uint32_t b_smem, int b_ext, int b_c0, int k_len,
                                               int N, bool accumulate_first) {
    const uint32_t idesc = umma_idesc_bf16_f32_major(128, N, false, true);
    const uint32_t lbo_b = (uint32_t)(b_ext >> 3) * 128u;
    for (int k = 0; k < k_len; k += 16) {
        const uint64_t da = umma_smem_desc(a_smem + (uint32_t)(k >> 3) * 128u, 128u, (uint32_t)(a_ext >> 3) * 128u);
        const uint64_t db = umma_smem_desc(b_smem + (uint32_t)(b_c0 >> 3) * 128u + (uint32_t)(k >> 3) * lbo_b, lbo_b, 128u);
        umma_bf16(tmem_d, da, db, idesc, accumulate_first || k > 0);
    }
}

}  // namespace qs
