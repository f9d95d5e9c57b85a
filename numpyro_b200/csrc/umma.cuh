// tcgen05 / TMEM / bulk-copy PTX wrappers and the shared-memory tile layouts of the many-chain GEMM regime (R3).
// sm_100a only.  Layout facts follow the canonical UMMA layouts spelled out in CUTLASS (cute/atom/mma_traits_sm100.hpp
// make_umma_desc, cute/arch/mma_sm100_desc.hpp); every wrapper in this file is exercised by scripts/ubench/umma_probe2.cu,
// whose output on a B200 is committed under profiles/.
//
// Tile layout used everywhere in R3 ("K-major, 128-byte swizzle", LayoutType::SWIZZLE_128B):
//   a tile holds R rows (the M or N index of the MMA, R % 8 == 0) of 32 fp32 (= 128 bytes, the K index);
//   row r starts at byte r * 128; inside the row the 16-byte chunk j (4 consecutive k) sits at chunk j ^ (r & 7).
//   Eight rows (1024 bytes) form one swizzle atom; atoms are stacked with stride SBO = 1024 bytes.  The tile base must be
//   1024-byte aligned because the hardware applies the XOR to absolute shared-memory address bits [4,7) ^= [7,10).
//   One kind::tf32 MMA consumes K = 8 (32 bytes); the k-th MMA of a tile uses start address + 32 k bytes.
// The images in HBM are stored in exactly this byte order, so one 1-D bulk copy (cp.async.bulk, the TMA engine) moves a
// tile; no tensor map is needed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2 {

#ifndef B2_D
#define B2_D __device__ __forceinline__
#endif

// ---- tile layout (host + device) ---------------------------------------------------------------------------
// float index of element (row r, k) inside a K-major SW128 tile of fp32
__host__ __device__ inline int sw128_index(int r, int k) { return r * 32 + ((((k >> 2) ^ (r & 7)) << 2) | (k & 3)); }

// ---- descriptors -------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (SmemDescriptor of mma_sm100_desc.hpp): bits [0,14) start >> 4, [16,30) leading byte
// offset >> 4, [32,46) stride byte offset >> 4, [46,48) version = 1, [61,64) layout type.
enum UmmaLayout : uint32_t { UMMA_SW_NONE = 0, UMMA_SW128_BASE32B = 1, UMMA_SW128 = 2, UMMA_SW64 = 4, UMMA_SW32 = 6 };
B2_D uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
// K-major SW128 tile: LBO is not used by the hardware for swizzled K-major operands (CUTLASS sets 1), SBO = 1024
B2_D uint64_t umma_desc_k_sw128(uint32_t saddr) { return umma_smem_desc(saddr, 16u, 1024u, UMMA_SW128); }

// Instruction descriptor (InstrDescriptor): [4,6) D format (1 = f32), [7,10) A format, [10,13) B format (2 = tf32,
// 1 = bf16), [15] A major, [16] B major (0 = K, 1 = MN), [17,23) N >> 3, [24,29) M >> 4.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, int a_mn = 0, int b_mn = 0) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- shared memory / mbarrier / bulk copy ------------------------------------------------------------------
B2_D uint32_t u_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
B2_D void u_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(u_smem(bar)), "r"(count) : "memory");
}
B2_D void u_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
B2_D void u_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(u_smem(bar)), "r"(bytes) : "memory");
}
B2_D void u_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(u_smem(bar)) : "memory");
}
B2_D bool u_mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(u_smem(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: gives up after `limit` clocks, raises *abort_flag (device memory) and returns false, so that a protocol
// bug can never wedge the GPU.
B2_D bool u_mbar_wait(uint64_t* bar, uint32_t parity, long long limit, unsigned int* abort_flag, unsigned int code) {
    if (u_mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    unsigned int it = 0;
    while (!u_mbar_try_wait(bar, parity)) {
        if ((++it & 63u) == 0u) {
            if (*(volatile unsigned int*)abort_flag) return false;
            if (clock64() - t0 > limit) { atomicCAS(abort_flag, 0u, code); return false; }
        }
    }
    return true;
}
B2_D void u_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(u_smem(dst)), "l"(src), "r"(bytes), "r"(u_smem(bar)) : "memory");
}
B2_D void u_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM ----------------------------------------------------------------------------------------------------
// A TMEM address is (lane << 16) | column; a CTA may allocate up to 512 columns (power of two >= 32) of 128 lanes x 32 bit.
B2_D void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {          // one whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(u_smem(slot_in_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
B2_D void tmem_dealloc(uint32_t taddr, uint32_t ncols) {                // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
B2_D void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
B2_D void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
B2_D void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
B2_D void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 bit, 16 consecutive columns: thread t of warp w owns lane 32 (w % 4) + t; v[i] <-> column i
B2_D void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
B2_D void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}

B2_D void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
}
B2_D void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

// ---- MMA (one elected thread issues for the CTA) ---------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]
B2_D void umma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]   (A: lane = m, one 32-bit column per k)
B2_D void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// all MMAs issued so far by this thread -> one arrival on `bar` when they have completed (implies fence::before_thread_sync)
B2_D void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(u_smem(bar)) : "memory");
}

// lo part of the tf32 split: x - trunc_tf32(x) (exact in fp32); the tensor core reads the top 19 bits of an operand
B2_D float u_tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

}  // namespace b2
