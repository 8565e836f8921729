// tcgen05 (5th-gen tensor core) path of the motif scorer: 3xTF32 GEMM chains with TMEM accumulators.
// This file starts with a self-test GEMM that pins the descriptor / TMEM conventions of tc.cuh on hardware.
#include "common.cuh"
#include "tc.cuh"

namespace tmb {

// C[128 x N] = A[128 x K] * B[N x K]^T, one CTA of 128 threads.  mode 0: single TF32 pass, 1: 3xTF32.
__global__ void __launch_bounds__(128)
selftest_gemm_kernel(const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ C, int K, int N, int mode) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5;
    uint8_t *a_hi = smem, *a_lo = a_hi + 128 * K * 4, *b_hi = a_lo + 128 * K * 4, *b_lo = b_hi + N * K * 4;
    for (int k = 0; k < K; k += 4) {
        const float4 v = *reinterpret_cast<const float4 *>(A + (size_t)t * K + k);
        float4 h, l;
        tc::split_tf32(v.x, h.x, l.x); tc::split_tf32(v.y, h.y, l.y); tc::split_tf32(v.z, h.z, l.z); tc::split_tf32(v.w, h.w, l.w);
        if (mode == 0) h = v;
        *reinterpret_cast<float4 *>(a_hi + tc::tile_off(128, t, k)) = h;
        *reinterpret_cast<float4 *>(a_lo + tc::tile_off(128, t, k)) = l;
    }
    for (int n = t; n < N; n += 128)
        for (int k = 0; k < K; k += 4) {
            const float4 v = *reinterpret_cast<const float4 *>(B + (size_t)n * K + k);
            float4 h, l;
            tc::split_tf32(v.x, h.x, l.x); tc::split_tf32(v.y, h.y, l.y); tc::split_tf32(v.z, h.z, l.z); tc::split_tf32(v.w, h.w, l.w);
            if (mode == 0) h = v;
            *reinterpret_cast<float4 *>(b_hi + tc::tile_off(N, n, k)) = h;
            *reinterpret_cast<float4 *>(b_lo + tc::tile_off(N, n, k)) = l;
        }
    uint32_t ncols = 32;
    while ((int)ncols < N) ncols <<= 1;
    if (t == 0) tc::mbar_init(&mbar, 1);
    if (warp == 0) tc::tmem_alloc(&tmem_slot, ncols);
    tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (t == 0) {
        const uint32_t idesc = tc::idesc_tf32(128, N);
        const uint32_t lbo_a = 128 * 16, lbo_b = (uint32_t)N * 16;
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint64_t ah = tc::smem_desc(tc::smem_u32(a_hi) + ks * 2 * lbo_a, lbo_a, 128);
            const uint64_t al = tc::smem_desc(tc::smem_u32(a_lo) + ks * 2 * lbo_a, lbo_a, 128);
            const uint64_t bh = tc::smem_desc(tc::smem_u32(b_hi) + ks * 2 * lbo_b, lbo_b, 128);
            const uint64_t bl = tc::smem_desc(tc::smem_u32(b_lo) + ks * 2 * lbo_b, lbo_b, 128);
            tc::mma_tf32(tmem, ah, bh, idesc, ks > 0);
            if (mode == 1) { tc::mma_tf32(tmem, al, bh, idesc, 1); tc::mma_tf32(tmem, ah, bl, idesc, 1); }
        }
        tc::mma_commit(&mbar);
    }
    tc::mbar_wait(&mbar, 0);
    tc::fence_after_sync();
    for (int c = 0; c < N; c += 16) {
        float v[16];
        tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) C[(size_t)t * N + c + i] = v[i];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, ncols);
}

}  // namespace tmb

using namespace tmb;

extern "C" int tm_selftest_gemm(const float *d_A, const float *d_B, float *d_C, int K, int N, int mode, tm_stream stream) {
    if (!d_A || !d_B || !d_C || K <= 0 || K % 8 || N < 16 || N > 256 || N % 16) { set_error("tm_selftest_gemm: need K %% 8 == 0, 16 <= N <= 256, N %% 16 == 0"); return TM_ERR_ARG; }
    const size_t smem = (size_t)(2 * 128 + 2 * N) * K * 4;
    if (smem > 200 * 1024) { set_error("tm_selftest_gemm: tile too large"); return TM_ERR_UNSUPPORTED; }
    TM_CUDA(cudaFuncSetAttribute(selftest_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    selftest_gemm_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(d_A, d_B, d_C, K, N, mode);
    TM_LAUNCH_CHECK();
    return TM_OK;
}
