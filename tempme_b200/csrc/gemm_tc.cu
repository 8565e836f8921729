// General fp32-accurate GEMM on the 5th-generation tensor cores (tcgen05, 3xTF32 split, TMEM accumulators) for the training side of the
// explainer: the forward, dgrad and wgrad products of every nn.Linear of TempME.forward when gradients are requested
// (tempme_b200/training.py: TcLinear).  C[M, N] (+)= A[M, K] . B[N, K]^T (+ bias[N]), all row-major fp32 with leading dimensions.
//   dgrad:  dX[M, K] = dY[M, N] . W[N, K]        -> A = dY, B = W^T   (the caller materialises W^T: a few KB)
//   wgrad:  dW[N, K] = dY[M, N]^T . X[M, K]      -> A = dY^T, B = X^T (reduction over the rows)
// One CTA of 128 threads per (128-row, <=128-column) tile of C; K in chunks of 32: the thread that owns a row writes the row's chunk of A into
// TMEM (hi | lo halves, TS-mode MMA), the CTA writes the chunk of B into shared memory in the K-major interleaved UMMA layout (tc.cuh), warp 0
// issues the 3 MMAs per K = 8 step from one elected lane.  The sizes here are small (a training batch is ~27,000 event rows): the kernel
// is written for exactness and reuse of the scorer's conventions, not for peak rate -- single-buffered, no TMA.
#include <stdint.h>

#include <algorithm>

#include "common.cuh"
#include "tc.cuh"

namespace tmb {

constexpr int kGK = 32;      // K columns per chunk

__global__ void __launch_bounds__(128)
gemm_tc_kernel(int64_t M, int N, int K, const float *__restrict__ A, int64_t lda, const float *__restrict__ B, int64_t ldb, float *__restrict__ C,
               int64_t ldc, const float *__restrict__ bias, int accumulate) {
    extern __shared__ __align__(128) uint8_t smem[];       // B chunk: hi tile | lo tile, each nt16 x 32 fp32 in the interleaved layout
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5;
    const int64_t row = (int64_t)blockIdx.x * 128 + t;
    const int n0 = blockIdx.y * 128, nt = min(128, N - n0), nt16 = (nt + 15) & ~15;
    uint8_t *b_hi = smem, *b_lo = smem + (size_t)nt16 * kGK * 4;
    if (t == 0) tc::mbar_init(&mbar, 1);
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 256);        // accumulator [0, nt16) | A hi [128, 160) | A lo [160, 192)
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot, lane_base = (uint32_t)(warp * 32) << 16;
    const bool live = row < M;
    const float *arow = A + (live ? row : 0) * lda;
    uint32_t phase = 0;
    for (int k0 = 0; k0 < K; k0 += kGK) {
        // ---- A chunk of this thread's row -> TMEM (zero beyond K and beyond M)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int k = k0 + 16 * h + i;
                const float x = (live && k < K) ? __ldg(arow + k) : 0.f;
                tc::split_tf32(x, hi[i], lo[i]);
            }
            tc::tmem_st16(tmem + lane_base + 128 + 16 * h, hi);
            tc::tmem_st16(tmem + lane_base + 160 + 16 * h, lo);
        }
        tc::tmem_st_wait();
        // ---- B chunk -> shared memory: thread t writes (n, k) pairs with k fastest so that global reads run along the rows
        for (int i = t; i < nt16 * kGK; i += 128) {
            const int n = i / kGK, k = i - n * kGK;
            const float x = (n < nt && k0 + k < K) ? __ldg(B + (int64_t)(n0 + n) * ldb + k0 + k) : 0.f;
            float h, l;
            tc::split_tf32(x, h, l);
            const uint32_t off = tc::tile_off(nt16, n, k);
            *reinterpret_cast<float *>(b_hi + off) = h;
            *reinterpret_cast<float *>(b_lo + off) = l;
        }
        tc::fence_smem_to_async();
        tc::fence_before_sync();
        __syncthreads();
        if (warp == 0) {
            tc::fence_after_sync();
            const uint32_t leader = tc::elect_one();
            const uint32_t idesc = tc::idesc_tf32(128, nt16), lbo_b = (uint32_t)nt16 * 16;
            uint64_t bh = tc::smem_desc(tc::smem_u32(b_hi), lbo_b, 128), bl = tc::smem_desc(tc::smem_u32(b_lo), lbo_b, 128);
            const uint64_t db = (2 * lbo_b) >> 4;
            for (int ks = 0; ks < kGK / 8; ++ks) {
                tc::mma_tf32_ts(tmem, tmem + 128 + 8 * ks, bh, idesc, (k0 > 0 || ks > 0) ? 1u : 0u, leader);
                tc::mma_tf32_ts(tmem, tmem + 160 + 8 * ks, bh, idesc, 1, leader);
                tc::mma_tf32_ts(tmem, tmem + 128 + 8 * ks, bl, idesc, 1, leader);
                bh += db; bl += db;
            }
            tc::mma_commit(&mbar, leader);
            __syncwarp();
        }
        tc::mbar_wait(&mbar, phase);       // the MMAs have read the A buffer and the B chunk: both may be overwritten
        phase ^= 1;
        tc::fence_after_sync();
    }
    // ---- epilogue: accumulator row -> (+ bias) (+ C) -> C
    for (int c = 0; c < nt16; c += 16) {
        float v[16];
        tc::tmem_ld16(tmem + lane_base + c, v);
        if (live) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int n = c + i;
                if (n < nt) {
                    float o = v[i] + (bias ? __ldg(bias + n0 + n) : 0.f);
                    float *p = C + row * ldc + n0 + n;
                    if (accumulate) o += *p;
                    *p = o;
                }
            }
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

}  // namespace tmb

using namespace tmb;

extern "C" int tm_gemm_tf32x3(int64_t M, int64_t N, int64_t K, const float *d_A, int64_t lda, const float *d_B, int64_t ldb, float *d_C, int64_t ldc,
                              const float *d_bias_or_null, int accumulate, tm_stream stream) {
    if (M < 0 || N < 0 || K < 0 || N > (1 << 20) || K > (1 << 20) || (M > 0 && N > 0 && (!d_C || ldc < N || (K > 0 && (!d_A || !d_B || lda < K || ldb < K))))) {
        set_error("tm_gemm_tf32x3: bad argument");
        return TM_ERR_ARG;
    }
    if (M == 0 || N == 0) return TM_OK;
    TM_DEVICE(device_of(d_C));
    const dim3 grid((unsigned)((M + 127) / 128), (unsigned)((N + 127) / 128));
    if (grid.y > 65535) { set_error("tm_gemm_tf32x3: N too large"); return TM_ERR_UNSUPPORTED; }
    const size_t smem = (size_t)2 * 128 * kGK * 4;          // hi + lo tiles of a 128 x 32 chunk of B
    gemm_tc_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(M, (int)N, (int)K, d_A, lda, d_B, ldb, d_C, ldc, d_bias_or_null, accumulate);
    TM_LAUNCH_CHECK();
    return TM_OK;
}
