// Blackwell (sm_100a) tensor-core plumbing used by the tensor-core form of the MLP layer products
// (hpv_varfwd_tc.cuh): tcgen05.mma kind::tf32 with the accumulators -- and, for the activations, the A operand --
// in tensor memory (TMEM), mbarrier completion, shared-memory matrix descriptors, the bulk-copy engine (TMA,
// cp.async.bulk) for the staging of the test-function tables.  Thin wrappers around the PTX; device only.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t hpv_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void hpv_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(hpv_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void hpv_mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void hpv_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hpv_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool hpv_mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(hpv_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Wait for the phase with the given parity.  try_wait suspends the thread in hardware for a bounded time per call; a
// wait that is still unsatisfied after 2^24 calls (seconds) means the producer (tensor core / copy engine) never
// completed: trap instead of hanging the GPU.  (No clock reads in the loop: it is the hottest spin of the kernels.)
__device__ __forceinline__ void hpv_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!hpv_mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) { asm volatile("trap;"); }
    }
}

// ---- bulk-copy engine (TMA): global -> shared, completion on an mbarrier ------------------------------------------
// bytes must be a multiple of 16, both addresses 16-byte aligned.
__device__ __forceinline__ void hpv_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(hpv_smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(hpv_smem_u32(bar))
                 : "memory");
}
// Orders earlier generic-proxy accesses of shared memory (ordinary loads/stores) before later async-proxy accesses
// (bulk copies, tensor-core operand reads).
__device__ __forceinline__ void hpv_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tensor memory ---------------------------------------------------------------------------------------------
// One warp (all 32 lanes) allocates ncols (power of two >= 32) columns; the base address lands in *smem_dst.
__device__ __forceinline__ void hpv_tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(hpv_smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void hpv_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void hpv_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void hpv_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void hpv_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void hpv_tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// TMEM address: lane in bits [31:16], column in bits [15:0].  A warp reaches the 32 lanes of its sub-partition
// (warp index mod 4); the .32x32b shapes give thread i of the warp lane 32*(warp%4)+i and N consecutive columns.
__device__ __forceinline__ void hpv_tmem_ld2(uint32_t taddr, uint32_t (&r)[2]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void hpv_tmem_ld4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void hpv_tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void hpv_tmem_st2(uint32_t taddr, uint32_t a, uint32_t b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void hpv_tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void hpv_tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// ---- descriptors -------------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, no swizzle ("interleave"): core matrices of 8 rows x 16 bytes, 128 contiguous
// bytes each.
//   K-major operand (K contiguous within a row: the activations [point][unit], the weights [out][in]):
//     byte(row r, element k) = (r%8)*16 + (r/8)*SBO + (k/4)*LBO + (k%4)*4          (tf32: 4 elements per 16 bytes)
//   MN-major operand (the row index contiguous):
//     byte(row r, element k) = (r%4)*4 + (k%8)*16 + (r/4)*SBO + (k/8)*LBO
// (canonical layouts of the UMMA descriptor as CUTLASS states them: cute/atom/mma_traits_sm100.hpp, make_umma_desc).
// layout_type: 0 none, 2 128-byte swizzle, 4 64-byte, 6 32-byte (cute::UMMA::LayoutType).
__device__ __forceinline__ uint64_t hpv_umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type = 0) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;                                   // descriptor version of sm_100
    d |= (uint64_t)(layout_type & 7u) << 61;
    return d;                                          // base offset 0
}
// Instruction descriptor of kind::tf32 with fp32 accumulation (cute/arch/mma_sm100_desc.hpp, InstrDescriptor).
__host__ __device__ constexpr uint32_t hpv_umma_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4)                     // accumulator format: F32
           | (2u << 7) | (2u << 10)      // A, B format: TF32
           | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] . B[smem]^T ; issued by ONE thread on behalf of the CTA.
__device__ __forceinline__ void hpv_umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T  (A: lane = row, one 32-bit column per K element).
__device__ __forceinline__ void hpv_umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// All tensor-core operations issued so far by this thread arrive on the mbarrier when they have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void hpv_umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(hpv_smem_u32(bar)) : "memory");
}

// ---- 3xTF32 split ---------------------------------------------------------------------------------------------------
// x = hi + lo with hi the TF32 rounding of x (10 explicit mantissa bits, low 13 bits zero) and lo = x - hi exact in
// fp32 (the tensor core reads only the TF32 part of lo, an error of 2^-22 |x|).  A . B ~ Ahi.Bhi + Alo.Bhi + Ahi.Blo
// drops only lo.lo (2^-22 relative): fp32-class accuracy from three TF32 products.
__device__ __forceinline__ void hpv_split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = h;
    lo = __float_as_uint(x - __uint_as_float(h));
}

#endif  // __CUDACC__
