// Hardware self-test GEMM of the tcgen05 / TMEM conventions in tc.cuh (descriptor layout, 3xTF32, TS mode).
// tests/test_gpu_tc.py runs it on the device; the scorer kernels in encoder_tc.cu rely on exactly these conventions.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "tc.cuh"
#include "timeenc.cuh"

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
    while ((int)ncols < (mode == 3 ? N + 2 * K : N)) ncols <<= 1;
    if (t == 0) tc::mbar_init(&mbar, 1);
    if (warp == 0) tc::tmem_alloc(&tmem_slot, ncols);
    tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (mode == 3) {       // TS mode: this thread's A row goes to TMEM columns [N, N+K) (hi) and [N+K, N+2K) (lo)
        for (int k = 0; k < K; k += 16) {
            float h[16], l[16];
            for (int i = 0; i < 16; ++i) { const float x = k + i < K ? A[(size_t)t * K + k + i] : 0.f; tc::split_tf32(x, h[i], l[i]); }
            tc::tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + N + k, h);
            tc::tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + N + K + k, l);
        }
        tc::tmem_st_wait();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    }
    long long t_issue0 = 0, t_issue1 = 0;
    if (warp == 0) {     // warp-uniform: every lane computes the (uniform) descriptors, one elected lane issues
        const uint32_t leader = tc::elect_one();
        const uint32_t idesc = tc::idesc_tf32(128, N);
        const uint32_t lbo_a = 128 * 16, lbo_b = (uint32_t)N * 16;
        t_issue0 = clock64();
        uint64_t ah = tc::smem_desc(tc::smem_u32(a_hi), lbo_a, 128), al = tc::smem_desc(tc::smem_u32(a_lo), lbo_a, 128);
        uint64_t bh = tc::smem_desc(tc::smem_u32(b_hi), lbo_b, 128), bl = tc::smem_desc(tc::smem_u32(b_lo), lbo_b, 128);
        const uint64_t da = (2 * lbo_a) >> 4, db = (2 * lbo_b) >> 4;       // descriptor start-address step per K = 8
        for (int ks = 0; ks < K / 8; ++ks) {
            if (mode == 3) {
                tc::mma_tf32_ts(tmem, tmem + N + 8 * ks, bh, idesc, ks > 0, leader);
                tc::mma_tf32_ts(tmem, tmem + N + K + 8 * ks, bh, idesc, 1, leader);
                tc::mma_tf32_ts(tmem, tmem + N + 8 * ks, bl, idesc, 1, leader);
            } else {
                tc::mma_tf32(tmem, ah, bh, idesc, ks > 0, leader);
                if (mode >= 1) { tc::mma_tf32(tmem, al, bh, idesc, 1, leader); tc::mma_tf32(tmem, ah, bl, idesc, 1, leader); }
            }
            ah += da; al += da; bh += db; bl += db;
        }
        tc::mma_commit(&mbar, leader);
        t_issue1 = clock64();
        if (mode >= 2 && leader) printf("[selftest] K=%d N=%d: %d MMAs, issue %lld cycles\n", K, N, (K / 8) * 3, t_issue1 - t_issue0);
        __syncwarp();
    }
    tc::mbar_wait(&mbar, 0);
    if (mode >= 2 && t == 0) printf("[selftest]   mode %d: MMAs complete %lld cycles after issue start\n", mode, clock64() - t_issue0);
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

// Issue-rate probe: `groups` x 8 back-to-back tcgen05.mma (M = 128, K = 8, tf32, A operand in TMEM) with identical, loop-invariant
// operands, so the loop holds nothing but the MMAs.  out[0] = cycles to issue them, out[1] = cycles until they have completed.
__global__ void __launch_bounds__(128) mma_rate_kernel(int N, int groups, long long *__restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < N * 32; i += 128) reinterpret_cast<float *>(smem)[i] = 0.f;
    if (t == 0) tc::mbar_init(&mbar, 1);
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
    tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    {
        float z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0.f;
        for (int c = 0; c < 512; c += 16) tc::tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + c, z);
        tc::tmem_st_wait();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        const uint32_t leader = tc::elect_one();
        const uint32_t idesc = tc::idesc_tf32(128, N);
        const uint64_t bdesc = tc::smem_desc(tc::smem_u32(smem), (uint32_t)N * 16, 128);
        const uint32_t ta = tmem + 256;
        t0 = clock64();
        for (int g = 0; g < groups; ++g)
            asm volatile(
                "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, 1, 0;\n\tsetp.ne.b32 q, %4, 0;\n\t"
                "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
                "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
                "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
                "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
                "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
                "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
                "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
                "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                :: "r"(tmem), "r"(ta), "l"(bdesc), "r"(idesc), "r"(leader) : "memory");
        tc::mma_commit(&mbar, leader);
        t1 = clock64();
        __syncwarp();
    }
    tc::mbar_wait(&mbar, 0);
    const long long t2 = clock64();
    if (t == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace tmb

using namespace tmb;

namespace tmb {
__global__ void selftest_cos_kernel(const float *__restrict__ x, float *__restrict__ out, int64_t n) {
    __shared__ uint2 ctab[kInv2PiN];
    cos_table_to_smem(ctab);
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = cos_accurate(x[i], ctab);
}
}  // namespace tmb

namespace tmb {
// raw dump of what eight gather4 instructions per warp leave in shared memory: 128 staging rows x 32 floats
__global__ void selftest_gather4_kernel(const __grid_constant__ CUtensorMap map, const int32_t *__restrict__ idx, int col, float *__restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    float *stg = reinterpret_cast<float *>(smem);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int i = t; i < 128 * 32; i += blockDim.x) stg[i] = -12345.f;
    if (t == 0) tc::mbar_init(&bar, 4);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    __syncthreads();
    const int id = idx[t], l4 = (lane & 7) * 4;
    const int r0 = __shfl_sync(0xffffffffu, id, l4), r1 = __shfl_sync(0xffffffffu, id, l4 + 1), r2 = __shfl_sync(0xffffffffu, id, l4 + 2), r3 = __shfl_sync(0xffffffffu, id, l4 + 3);
    if (lane == 0) tc::mbar_expect_tx(&bar, 8u * 512u);
    __syncwarp();
    if (lane < 8)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n"
                     :: "r"(tc::smem_u32(stg + (warp * 32 + l4) * 32)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3),
                        "r"(tc::smem_u32(&bar)) : "memory");
    tc::mbar_wait(&bar, 0);
    for (int i = t; i < 128 * 32; i += blockDim.x) out[i] = stg[i];
}
}  // namespace tmb

extern "C" int tm_selftest_gather4(const float *d_table, int64_t rows, int dim, const int32_t *d_idx128, int col, int swizzle128, float *d_out, tm_stream stream) {
    TM_DEVICE(device_of(d_out));
    CUtensorMap map;
    if (!make_gather_map(&map, d_table, rows, dim, swizzle128)) { set_error("tm_selftest_gather4: cuTensorMapEncodeTiled failed"); return TM_ERR_CUDA; }
    selftest_gather4_kernel<<<1, 128, 128 * 32 * 4, (cudaStream_t)stream>>>(map, d_idx128, col, d_out);
    TM_LAUNCH_CHECK();
    return TM_OK;
}

extern "C" int tm_selftest_cos(const float *d_x, float *d_out, int64_t n, tm_stream stream) {
    if (!d_x || !d_out || n < 0) { set_error("tm_selftest_cos: bad argument"); return TM_ERR_ARG; }
    if (n == 0) return TM_OK;
    TM_DEVICE(device_of(d_out));
    selftest_cos_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 1184), 256, 0, (cudaStream_t)stream>>>(d_x, d_out, n);
    TM_LAUNCH_CHECK();
    return TM_OK;
}

extern "C" int tm_selftest_gemm(const float *d_A, const float *d_B, float *d_C, int K, int N, int mode, tm_stream stream) {
    if (mode == 3 && (K % 16 || N + 2 * K > 512)) { set_error("tm_selftest_gemm: TS mode needs K %% 16 == 0 and N + 2K <= 512"); return TM_ERR_ARG; }
    if (!d_A || !d_B || !d_C || K <= 0 || K % 8 || N < 16 || N > 256 || N % 16) { set_error("tm_selftest_gemm: need K %% 8 == 0, 16 <= N <= 256, N %% 16 == 0"); return TM_ERR_ARG; }
    const size_t smem = (size_t)(2 * 128 + 2 * N) * K * 4;
    if (smem > 200 * 1024) { set_error("tm_selftest_gemm: tile too large"); return TM_ERR_UNSUPPORTED; }
    TM_DEVICE(device_of(d_C));
    TM_CUDA(cudaFuncSetAttribute(selftest_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    selftest_gemm_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(d_A, d_B, d_C, K, N, mode);
    TM_LAUNCH_CHECK();
    return TM_OK;
}

extern "C" int tm_selftest_mma_rate(int N, int groups, long long *d_out, tm_stream stream) {
    if (!d_out || N < 16 || N > 256 || N % 16 || groups < 1) { set_error("tm_selftest_mma_rate: need 16 <= N <= 256, N %% 16 == 0, groups >= 1"); return TM_ERR_ARG; }
    TM_DEVICE(device_of(d_out));
    mma_rate_kernel<<<1, 128, (size_t)N * 32 * 4, (cudaStream_t)stream>>>(N, groups, d_out);
    TM_LAUNCH_CHECK();
    return TM_OK;
}
