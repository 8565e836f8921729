// tcgen05 / TMEM / mbarrier primitives for sm_100a, hand-written (inline PTX).
//
// Operand layout used throughout (K-major, no swizzle -- the "interleaved" canonical UMMA layout):
// a tile of R rows x KC fp32 columns is stored as KC/4 slabs, one per 16-byte K chunk; inside a slab
// the rows are packed in groups of eight "core matrix" rows of 16 bytes:
//     byte(row, k) = (k / 4) * LBO + (row / 8) * 128 + (row % 8) * 16 + (k % 4) * 4,   LBO = R * 16
// so a warp that writes 32 consecutive rows of one K chunk writes 512 contiguous bytes (no bank
// conflicts), and the tensor core reads whole 128-byte core matrices.
// One tcgen05.mma.kind::tf32 consumes K = 8 (two K chunks): descriptor start = tile + kstep * 2 * LBO.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tmb {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout), SWIZZLE_NONE, K-major
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, dense (cute::UMMA::InstrDescriptor)
__device__ __forceinline__ uint32_t idesc_tf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// exactly one lane of a converged warp gets true: the form ptxas recognises for issuing UTC* instructions without
// a per-instruction lane-election loop (issue from `if (threadIdx.x == 0)` costs ~75 cycles per MMA instead)
__device__ __forceinline__ uint32_t elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
        "elect.sync rx|px, 0xFFFFFFFF;\n\t"
        "@px mov.s32 %0, 1;\n\t}\n"
        : "+r"(pred));
    return pred;
}

// The whole (converged) warp executes these; `leader` (from elect_one) predicates the instruction itself, so the
// operands stay warp-uniform (uniform registers) and ptxas emits back-to-back UTC* instructions.
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate, uint32_t leader = 1) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader) : "memory");
}
// TS mode: A operand read from TMEM (lane = row, one 32-bit column per K element), B from shared memory
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate, uint32_t leader = 1) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader) : "memory");
}
// One K = 8 step of a 3xTF32 product with the A operand in TMEM: D (+)= Ah Bh + Al Bh + Ah Bl.  The B descriptors are passed as their low
// words (start address | LBO, smem_desc_lo) with the shared high word (smem_desc_hi): stepping along K is a 32-bit add, and ptxas keeps
// the operands in uniform registers.
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16); }
__device__ __forceinline__ uint32_t smem_desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }
__device__ __forceinline__ void mma3_tf32_ts(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t bh_lo, uint32_t bl_lo, uint32_t desc_hi, uint32_t idesc,
                                             uint32_t accumulate, uint32_t leader) {
    asm volatile(
        "{\n\t.reg .pred p, q, one;\n\t.reg .b64 dh, dl;\n\t"
        "setp.ne.b32 p, %7, 0;\n\tsetp.ne.b32 q, %8, 0;\n\tsetp.eq.b32 one, 0, 0;\n\t"
        "mov.b64 dh, {%3, %5};\n\tmov.b64 dl, {%4, %5};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], dh, %6, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%2], dh, %6, one;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], dl, %6, one;\n\t}\n"
        :: "r"(tmem_d), "r"(a_hi), "r"(a_lo), "r"(bh_lo), "r"(bl_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate), "r"(leader) : "memory");
}
// all MMAs issued so far by the leader arrive on the mbarrier when they complete
__device__ __forceinline__ void mma_commit(uint64_t *mbar, uint32_t leader = 1) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" :: "r"(smem_u32(mbar)), "r"(leader) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (the tensor core's operand reads)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t *mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(mbar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
// try_wait is potentially blocking: with a suspend-time hint the hardware parks the thread until the phase completes (or the
// hint expires) instead of returning after its short default limit -- far fewer spin iterations competing for issue slots
__device__ __forceinline__ void mbar_wait(uint64_t *mbar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}\n"
        :: "r"(smem_u32(mbar)), "r"(parity), "r"(20000u) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" :: "r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared (16-byte aligned, size multiple of 16); completes on the mbarrier
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
}

__device__ __forceinline__ void tma_load_1d_s(uint32_t dst_smem_addr, const void *src_gmem, uint32_t bytes, uint64_t *mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 :: "r"(dst_smem_addr), "l"(src_gmem), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
}

// TMEM allocation: one full warp; ncols power of two in [32, 512]; address lands in *slot (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(taddr), "r"(ncols) : "memory");
}

// 32 lanes x 32 bit, 16 consecutive columns: thread t of warp w reads TMEM lane 32*(w%4)+t, columns [c, c+16)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float v[8]) {
    uint32_t r[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// registers -> TMEM: thread t of warp w writes 16 consecutive columns of lane 32*(w%4)+t
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float v[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n"
        :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
           "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
           "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
           "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float v[8]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n"
        :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
           "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// 3xTF32 split: hi keeps the 10 explicit mantissa bits the tensor core reads, lo = exact remainder
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    lo = x - hi;
}

// byte offset of element (row, k) inside a tile of R rows (see header comment); k multiple of 4 for 16-byte stores
__device__ __forceinline__ uint32_t tile_off(int R, int row, int k) {
    return (uint32_t)((k >> 2) * (R * 16) + (row >> 3) * 128 + (row & 7) * 16 + (k & 3) * 4);
}

}  // namespace tc
}  // namespace tmb
