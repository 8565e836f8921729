// tcgen05 (5th-gen tensor core) path of the motif scorer: 3xTF32 GEMM chains with TMEM accumulators.
// This file starts with a self-test GEMM that pins the descriptor / TMEM conventions of tc.cuh on hardware.
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

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

}  // namespace tmb

using namespace tmb;

extern "C" int tm_selftest_gemm(const float *d_A, const float *d_B, float *d_C, int K, int N, int mode, tm_stream stream) {
    if (mode == 3 && (K % 16 || N + 2 * K > 512)) { set_error("tm_selftest_gemm: TS mode needs K %% 16 == 0 and N + 2K <= 512"); return TM_ERR_ARG; }
    if (!d_A || !d_B || !d_C || K <= 0 || K % 8 || N < 16 || N > 256 || N % 16) { set_error("tm_selftest_gemm: need K %% 8 == 0, 16 <= N <= 256, N %% 16 == 0"); return TM_ERR_ARG; }
    const size_t smem = (size_t)(2 * 128 + 2 * N) * K * 4;
    if (smem > 200 * 1024) { set_error("tm_selftest_gemm: tile too large"); return TM_ERR_UNSUPPORTED; }
    TM_CUDA(cudaFuncSetAttribute(selftest_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    selftest_gemm_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(d_A, d_B, d_C, K, N, mode);
    TM_LAUNCH_CHECK();
    return TM_OK;
}

// =================================================================================================
// Tensor-core scorer.  Two kernels per slab of motifs:
//   event_tc_kernel : one CTA = 256 threads = 128 event rows (3 per motif) x 2 column halves.  lin_event -> event
//                     MLP for both orientations; writes updated_feature rows (explainer.py:179-185) as 16 KB
//                     [128 motifs x 32 columns] slabs, one per (motif tile, position, column chunk).
//   motif_tc_kernel : one CTA = 128 motifs x 2 column halves.  W1/W2 projections, temporal attention, attention
//                     MLP, category one-hot, final MLP, sigmoid (explainer.py:190-200, 789-846); its operand rows
//                     arrive as TMA bulk copies of those slabs.
// Every Linear is a 3xTF32 tcgen05.mma chain accumulating in TMEM.  Thread (row, half) owns half of the columns
// of row `row` of the tile = TMEM lane `row`: the "A-fill" of a layer reads the previous layer's accumulator row
// from TMEM (or the gathered features), applies bias / activation in fp32 registers, splits into tf32 hi + lo and
// stores into the K-major operand tile in shared memory; weights arrive pre-split and pre-tiled by TMA.
// =================================================================================================
namespace tmb {

constexpr int kKC = 32;           // K columns per operand chunk
constexpr int kTcThreads = 256;   // 128 rows x 2 column halves
constexpr int kSlabFloats = 128 * kKC;
constexpr int kReplicas = 1;      // copies of the packed weights; CTA b streams from copy b % kReplicas (spreads the L2 slices that serve the broadcast)

struct TcLin { int64_t w; int b, K8, N16; };   // chunk c at w + c * 2 * N16 * kKC floats: [hi tile | lo tile]; bias at cst[b]
struct TcLayout {
    int D, Ed, H, M, ev, use_temporal, if_cat;
    TcLin evt, g0, g2, w1, w2, a0, a3, m0, m3;
    int w5, b5, freq, phase, n_cst;     // offsets inside the constant block (biases, MLP.5, TimeEncode parameters)
    int64_t cst, total;
};

__host__ __device__ static inline int r8(int x) { return (x + 7) & ~7; }
__host__ __device__ static inline int r16(int x) { return (x + 15) & ~15; }

TcLayout make_tc_layout(const tm_encoder_desc &d) {
    TcLayout L;
    memset(&L, 0, sizeof L);
    L.D = d.node_dim; L.Ed = d.edge_dim; L.H = d.hid_dim; L.use_temporal = d.use_temporal; L.if_cat = d.if_cat;
    L.M = d.if_cat ? d.hid_dim + 12 : d.hid_dim; L.ev = L.Ed + 3 + L.D;
    int64_t o = 0;
    int co = 0;
    auto lin = [&](int K, int N) {
        TcLin l; l.K8 = r8(K); l.N16 = r16(N);
        const int nch = (l.K8 + kKC - 1) / kKC;
        l.w = o; o += (int64_t)nch * 2 * l.N16 * kKC; l.b = co; co += l.N16;
        return l;
    };
    L.evt = lin(L.ev, L.D); L.g0 = lin(L.D, L.H); L.g2 = lin(L.H, L.H);
    L.w1 = lin(2 * L.H, 2 * L.H); L.w2 = lin(2 * L.H, 2 * L.H); L.a0 = lin(2 * L.H, L.H); L.a3 = lin(L.H, L.H);
    L.m0 = lin(L.M, L.M); L.m3 = lin(L.M, L.H);
    L.w5 = co; co += r16(L.H); L.b5 = co; co += 16;
    L.freq = co; co += r16(L.D); L.phase = co; co += r16(L.D);
    L.n_cst = co; L.cst = o; L.total = o + co;
    return L;
}

// host: nn.Linear weight [N][K] -> per K chunk the [hi | lo] operand tiles in the tc.cuh layout (R = N16 rows)
void pack_tc_lin(const TcLayout &L, const TcLin &l, int K, int N, const float *w, const float *b, float *blob) {
    const int nch = (l.K8 + kKC - 1) / kKC;
    for (int c = 0; c < nch; ++c) {
        float *hi = blob + l.w + (int64_t)c * 2 * l.N16 * kKC, *lo = hi + (int64_t)l.N16 * kKC;
        for (int n = 0; n < N; ++n)
            for (int kk = 0; kk < kKC && c * kKC + kk < K; ++kk) {
                const float x = w[(int64_t)n * K + c * kKC + kk];
                uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u;
                float h; memcpy(&h, &u, 4);
                const int64_t off = ((kk >> 2) * (l.N16 * 16) + (n >> 3) * 128 + (n & 7) * 16 + (kk & 3) * 4) / 4;
                hi[off] = h; lo[off] = x - h;
            }
    }
    for (int n = 0; n < N; ++n) blob[L.cst + l.b + n] = b[n];
}

// static per-tile schedules: weight chunks (TMA, global -> smem) and, for the motif kernel, the updated_feature
// slabs a round's A-fill reads (TMA into the staging buffer one round ahead)
struct ChunkTab { int n; int64_t off[40]; int bytes[40]; };
struct StageTab { int8_t ns[40]; int8_t pos[40][2]; int8_t ch[40][2]; };

struct TcCtx {
    uint8_t *a;            // A operand region: tile (m-block mb, hi/lo h) at a + (2*mb + h) * 128*kKC*4
    uint32_t a_s, b_s;     // shared-space addresses of the A region and of weight buffer 0
    uint32_t b_bytes;      // bytes of one weight chunk buffer [hi | lo]
    float *stage;          // 2 slabs of updated_feature columns (motif kernel)
    const float *cst;      // constant block in shared memory
    uint64_t *bars;        // [0] MMA done, [1], [2] weight buffers, [3] staging
    uint32_t mma_phase, s_phase;
    int nbuf, ri;          // weight buffers (1 or 2); round index inside the tile
    long long *dbg;        // optional per-chunk timestamps of CTA 0 (TEMPME_TC_TIMING), 6 slots per chunk
    int64_t seq, total;    // running chunk counter of this CTA / chunks it will consume in total
    const float *blob, *F;
};
constexpr uint32_t kATile = 128 * kKC * 4;

__device__ __forceinline__ void tc_prefetch_b(const TcCtx &x, const ChunkTab &tab, int64_t seq) {   // thread 0 only
    const int i = (int)(seq % tab.n), buf = x.nbuf == 2 ? (int)(seq & 1) : 0;
    tc::mbar_expect_tx(x.bars + 1 + buf, (uint32_t)tab.bytes[i]);
    tc::tma_load_1d_s(x.b_s + buf * x.b_bytes, x.blob + tab.off[i], (uint32_t)tab.bytes[i], x.bars + 1 + buf);
}
__device__ __forceinline__ void tc_prefetch_stage(const TcCtx &x, const ChunkTab &tab, const StageTab &st, int64_t seq) {   // thread 0 only
    const int i = (int)(seq % tab.n), ns = st.ns[i];
    if (!ns) return;
    const int64_t tile = blockIdx.x + (seq / tab.n) * gridDim.x;
    tc::mbar_expect_tx(x.bars + 3, (uint32_t)(ns * kSlabFloats * 4));
    for (int k = 0; k < ns; ++k)
        tc::tma_load_1d(x.stage + k * kSlabFloats, x.F + ((tile * 3 + st.pos[i][k]) * 4 + st.ch[i][k]) * kSlabFloats, kSlabFloats * 4, x.bars + 3);
}

__device__ __forceinline__ void store_a4(const TcCtx &x, int mb, int row, int k, float4 v) {
    float4 h, l;
    tc::split_tf32(v.x, h.x, l.x); tc::split_tf32(v.y, h.y, l.y); tc::split_tf32(v.z, h.z, l.z); tc::split_tf32(v.w, h.w, l.w);
    uint8_t *p = x.a + (uint32_t)(2 * mb) * kATile + tc::tile_off(128, row, k);
    *reinterpret_cast<float4 *>(p) = h;
    *reinterpret_cast<float4 *>(p + kATile) = l;
}
// updated_feature slab element (row, k): 128-byte rows with the 16-byte pieces XOR-swizzled by the row (bank spread)
__device__ __forceinline__ int slab_off(int row, int k) { return row * kKC + ((((k >> 2) ^ (row & 7)) << 2) | (k & 3)); }

// One Linear over MB row blocks that share the weight: acc[mb] (TMEM column) = A[mb] * W^T, K streamed in chunks
// of kKC columns.  fill(c, kcols) writes this thread's share of columns [c*kKC, c*kKC + kcols) of every A block.
template <int MB, typename Fill>
__device__ __forceinline__ void tc_linear(const TcLin l, TcCtx &x, const ChunkTab &tab, const StageTab *st, uint32_t tmem,
                                          const int (&acc_col)[MB], Fill fill) {
    const int t = threadIdx.x;
    const int nch = (l.K8 + kKC - 1) / kKC;
    const uint32_t idesc = tc::idesc_tf32(128, l.N16);
    for (int c = 0; c < nch; ++c) {
        const int kcols = min(kKC, l.K8 - c * kKC);
        const bool tim = x.dbg && blockIdx.x == 0 && t == 0 && x.seq < 64;
        const bool tim7 = x.dbg && blockIdx.x == 0 && t == 224 && x.seq < 64;
        if (tim7) x.dbg[768 + x.seq * 4 + 0] = clock64();
        if (tim) x.dbg[x.seq * 6 + 0] = clock64();
        if (x.nbuf == 2 && t == 0 && x.seq + 1 < x.total) tc_prefetch_b(x, tab, x.seq + 1);   // buffer released by the MMA wait of chunk seq-1
        if (st && st->ns[x.ri]) { tc::mbar_wait(x.bars + 3, x.s_phase); x.s_phase ^= 1; }        // this round's slabs have landed
        fill(c, kcols);
        if (tim) x.dbg[x.seq * 6 + 1] = clock64();
        if (tim7) x.dbg[768 + x.seq * 4 + 1] = clock64();
        tc::fence_smem_to_async();
        if (tim7) x.dbg[768 + x.seq * 4 + 2] = clock64();
        tc::fence_before_sync();
        __syncthreads();
        if (tim7) x.dbg[768 + x.seq * 4 + 3] = clock64();
        if (t < 32) {        // warp 0 (warp-uniform branch): lane 0 feeds the TMA queues, one elected lane issues the MMAs
            if (t == 0 && st && x.seq + 1 < x.total) tc_prefetch_stage(x, tab, *st, x.seq + 1);   // staging buffer is free again
            __syncwarp();
            const int buf = x.nbuf == 2 ? (int)(x.seq & 1) : 0;
            if (tim) x.dbg[x.seq * 6 + 2] = clock64();
            tc::mbar_wait(x.bars + 1 + buf, (uint32_t)((x.nbuf == 2 ? (x.seq >> 1) : x.seq) & 1));   // weight chunk has landed (TMA)
            if (tim) x.dbg[x.seq * 6 + 3] = clock64();
            tc::fence_after_sync();
            const uint32_t leader = tc::elect_one();
            const uint32_t lbo_a = 128 * 16, lbo_b = (uint32_t)l.N16 * 16;
            const uint32_t b_base = x.b_s + (uint32_t)buf * x.b_bytes;
            uint64_t bh = tc::smem_desc(b_base, lbo_b, 128), bl = tc::smem_desc(b_base + (uint32_t)l.N16 * kKC * 4, lbo_b, 128);
            uint64_t ah[MB], al[MB];
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) { ah[mb] = tc::smem_desc(x.a_s + (uint32_t)(2 * mb) * kATile, lbo_a, 128); al[mb] = tc::smem_desc(x.a_s + (uint32_t)(2 * mb + 1) * kATile, lbo_a, 128); }
            const uint64_t da = (2 * lbo_a) >> 4, db = (2 * lbo_b) >> 4;      // descriptor start-address step per K = 8
            for (int ks = 0; ks < kcols / 8; ++ks) {
#pragma unroll
                for (int mb = 0; mb < MB; ++mb) {
                    tc::mma_tf32(tmem + acc_col[mb], ah[mb], bh, idesc, (c | ks) != 0, leader);
                    tc::mma_tf32(tmem + acc_col[mb], al[mb], bh, idesc, 1, leader);
                    tc::mma_tf32(tmem + acc_col[mb], ah[mb], bl, idesc, 1, leader);
                    ah[mb] += da; al[mb] += da;
                }
                bh += db; bl += db;
            }
            tc::mma_commit(x.bars, leader);
            if (tim) x.dbg[x.seq * 6 + 4] = clock64();
            __syncwarp();
        }
        tc::mbar_wait(x.bars, x.mma_phase);
        if (tim) x.dbg[x.seq * 6 + 5] = clock64();
        x.mma_phase ^= 1;
        if (x.nbuf == 1 && t == 0 && x.seq + 1 < x.total) tc_prefetch_b(x, tab, x.seq + 1);   // single buffer: refill right after the MMA released it
        x.seq++;
        x.ri = x.ri + 1 == tab.n ? 0 : x.ri + 1;
        tc::fence_after_sync();
    }
}

struct TcArgs {
    int64_t n_motifs, W, group, m_begin;     // this launch scores motifs [m_begin, m_begin + slab)
    int64_t slab;
    const int32_t *nodes, *eidx;
    const float *t;
    const uint8_t *cat;
    const float *cut, *eid, *node_feat, *edge_feat, *std_;
    int64_t n_node_rows, n_edge_rows;
    float *F;                                // updated_feature slabs of the slab of motifs: [tile][pos][chunk][128][kKC]
    float *scores;
    uint32_t tmem_cols;
    int b_bytes, nbuf;                       // bytes of one weight-chunk buffer, number of buffers
    int replicas;                            // weight copies in the blob (CTA b uses copy b % replicas)
    long long *dbg;
};

__device__ __forceinline__ void tc_setup(uint8_t *smem, TcCtx &x, const TcLayout &L, const TcArgs &a, int mb, bool stage, uint64_t *bars,
                                         uint32_t *tmem_slot, const float *blob) {
    uint8_t *p = smem;
    x.a = p; x.a_s = tc::smem_u32(p); p += (uint32_t)(2 * mb) * kATile;
    x.b_s = tc::smem_u32(p); x.b_bytes = (uint32_t)a.b_bytes; p += (size_t)a.nbuf * a.b_bytes;
    x.stage = reinterpret_cast<float *>(p); if (stage) p += 2 * kSlabFloats * 4;
    float *cst = reinterpret_cast<float *>(p);
    x.cst = cst;
    x.bars = bars;
    x.mma_phase = 0; x.s_phase = 0; x.seq = 0; x.ri = 0; x.nbuf = a.nbuf; x.blob = blob; x.F = a.F; x.dbg = a.dbg;
    for (int i = threadIdx.x; i < L.n_cst; i += blockDim.x) cst[i] = __ldg(blob + L.cst + i);
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) tc::mbar_init(bars + i, 1); }
    if ((threadIdx.x >> 5) == 0) tc::tmem_alloc(tmem_slot, a.tmem_cols);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
}

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }

// cos(x) for the TimeEncode arguments (they reach 1e8 and beyond, where the library cosf takes its slow path).
// Branch-free exact argument reduction in 64-bit integer arithmetic: |x| = m * 2^e with a 24-bit integer m, so
// frac(|x| / 2pi) = frac(m * frac(2^e / 2pi)); kInv2Pi[e + 44] holds frac(2^e / 2pi) in 0.64 fixed point and the
// product wraps mod 2^64 for free (error < 2^-40 turns).  The turn fraction is split into a quadrant and an angle in
// [-pi/4, pi/4) for the fdlibm single-precision sin/cos kernels.  Max error 1.4 ulp of 1.0 against the exact cosine
// of the fp32 argument over |x| <= 1e11 (cosf: 1-2 ulp); |x| >= 2^43, inf and nan go to cosf.
__constant__ unsigned long long kInv2Pi[64] = {
    0x0000000000028be6ull, 0x00000000000517ccull, 0x00000000000a2f98ull, 0x0000000000145f30ull, 0x000000000028be60ull, 0x0000000000517cc1ull,
    0x0000000000a2f983ull, 0x000000000145f306ull, 0x00000000028be60dull, 0x000000000517cc1bull, 0x000000000a2f9836ull, 0x00000000145f306dull,
    0x0000000028be60dbull, 0x00000000517cc1b7ull, 0x00000000a2f9836eull, 0x0000000145f306dcull, 0x000000028be60db9ull, 0x0000000517cc1b72ull,
    0x0000000a2f9836e4ull, 0x000000145f306dc9ull, 0x00000028be60db93ull, 0x000000517cc1b727ull, 0x000000a2f9836e4eull, 0x00000145f306dc9cull,
    0x0000028be60db939ull, 0x00000517cc1b7272ull, 0x00000a2f9836e4e4ull, 0x0000145f306dc9c8ull, 0x000028be60db9391ull, 0x0000517cc1b72722ull,
    0x0000a2f9836e4e44ull, 0x000145f306dc9c88ull, 0x00028be60db93910ull, 0x000517cc1b727220ull, 0x000a2f9836e4e441ull, 0x00145f306dc9c882ull,
    0x0028be60db939105ull, 0x00517cc1b727220aull, 0x00a2f9836e4e4415ull, 0x0145f306dc9c882aull, 0x028be60db9391054ull, 0x0517cc1b727220a9ull,
    0x0a2f9836e4e44152ull, 0x145f306dc9c882a5ull, 0x28be60db9391054aull, 0x517cc1b727220a94ull, 0xa2f9836e4e441529ull, 0x45f306dc9c882a53ull,
    0x8be60db9391054a7ull, 0x17cc1b727220a94full, 0x2f9836e4e441529full, 0x5f306dc9c882a53full, 0xbe60db9391054a7full, 0x7cc1b727220a94feull,
    0xf9836e4e441529fcull, 0xf306dc9c882a53f8ull, 0xe60db9391054a7f0ull, 0xcc1b727220a94fe1ull, 0x9836e4e441529fc2ull, 0x306dc9c882a53f84ull,
    0x60db9391054a7f09ull, 0xc1b727220a94fe13ull, 0x836e4e441529fc27ull, 0x06dc9c882a53f84eull};

__device__ __forceinline__ float cos_accurate(float x) {
    const uint32_t bits = __float_as_uint(x) & 0x7fffffffu;
    const int e = (int)(bits >> 23) - 150;                 // |x| = m * 2^e
    if (e > 19) return cosf(x);
    const unsigned long long m = e < -44 ? 0ull : (unsigned long long)((bits & 0x7fffffu) | 0x800000u);   // tiny |x|: angle 0
    const unsigned long long p = m * kInv2Pi[max(e, -44) + 44] + (1ull << 61);     // turn fraction + 1/8 turn, 0.64 fixed point
    const int q = (int)(p >> 62);
    const long long r = (long long)(p & ((1ull << 62) - 1)) - (1ll << 61);           // angle inside the quadrant, [-1/8, 1/8) turn
    const float th = (float)r * 3.40612158008655459e-19f /* 2 pi / 2^64 */, z = th * th;
    const float cs = fmaf(z, fmaf(z, fmaf(z, fmaf(z, 2.43904487962774090654e-5f, -1.38867637746099294692e-3f), 4.16666233237390631894e-2f), -4.99999997251031003120e-1f), 1.f);
    const float sn = fmaf(th * z, fmaf(z, fmaf(z, fmaf(z, 2.7183114939898219064e-6f, -1.98393348360966317347e-4f), 8.3333293858894631756e-3f), -1.66666666416265235595e-1f), th);
    const float v = (q & 1) ? sn : cs;                     // cos(q pi/2 + th) = {cs, -sn, -cs, sn}[q]
    return ((q + 1) & 2) ? -v : v;
}

// ---------------------------------------------------------------------------------------------
// event kernel: rows r = 3 * motif + position of the slab
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads)
event_tc_kernel(const TcLayout L, const ChunkTab tab, const float *__restrict__ blob0, const TcArgs a) {
    const float *__restrict__ blob = blob0 + (int64_t)(blockIdx.x % a.replicas) * ((L.total + 31) & ~(int64_t)31);
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[4];
    __shared__ uint32_t tmem_slot;
    TcCtx x;
    tc_setup(smem, x, L, a, 2, false, bars, &tmem_slot, blob);
    const int t = threadIdx.x, warp = t >> 5, row = t & 127, half = t >> 7, kb = 16 * half;
    const uint32_t tmem = tmem_slot, lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int H = L.H, D = L.D, Ed = L.Ed;
    const int colZ = 0, colE = 2 * H, colF = 2 * H;      // Z [0,2H) ; E [2H, 2H + r16(D)) ; F aliases E (dead by then)
    const int64_t n_rows = 3 * min(a.slab, a.n_motifs - a.m_begin);
    const int64_t n_tiles = (n_rows + 127) / 128;
    x.total = ((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) * tab.n;
    if (t == 0 && x.total > 0) tc_prefetch_b(x, tab, 0);
    const float *cst = x.cst;
    const bool ed_vec = (Ed & 3) == 0, d_vec = (D & 3) == 0;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t r = tile * 128 + row;
        const bool live = r < n_rows;
        const int64_t ml = live ? r / 3 : 0, gm = a.m_begin + ml;
        const int pos = live ? (int)(r - 3 * ml) : 0;
        int64_t e = 0, ns = 0, nt = 0; float dt = 0.f;
        if (live) {
            e = a.eidx[gm * 3 + pos]; ns = a.nodes[gm * 6 + 2 * pos]; nt = a.nodes[gm * 6 + 2 * pos + 1];
            dt = __fsub_rn(a.t[gm * 3 + 2], a.t[gm * 3 + pos]);                       // explainer.py:326
        }
        const bool e_ok = live && e >= 0 && e < a.n_edge_rows, s_ok = live && ns >= 0 && ns < a.n_node_rows, t_ok = live && nt >= 0 && nt < a.n_node_rows;
        const float *ef = a.edge_feat + e * Ed, *sf = a.node_feat + ns * D, *tf = a.node_feat + nt * D;
        const float *ei = a.eid ? a.eid + gm * 9 + pos * 3 : nullptr;
        auto xval = [&](int j) -> float {                                              // event_features column j (:179)
            if (j < Ed) return e_ok ? __ldg(ef + j) : 0.f;
            if (j < Ed + 3) return (live && ei) ? __ldg(ei + (j - Ed)) : 0.f;
            if (j < L.ev) { const int k = j - Ed - 3; return live ? cos_accurate(__fadd_rn(__fmul_rn(dt, cst[L.freq + k]), cst[L.phase + k])) : 0.f; }   // :55-58
            return 0.f;
        };
        // ---- lin_event (:93)
        { const int acc[1] = {colE};
          tc_linear<1>(L.evt, x, tab, nullptr, tmem, acc, [&](int c, int kcols) {
              float4 v[4];
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                  const int k = kb + 4 * g, j = c * kKC + k;
                  if (k >= kcols) { v[g] = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
                  if (ed_vec && j + 3 < Ed) v[g] = e_ok ? ldg4(ef + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                  else v[g] = make_float4(xval(j), xval(j + 1), xval(j + 2), xval(j + 3));
              }
#pragma unroll
              for (int g = 0; g < 4; ++g) if (kb + 4 * g < kcols) store_a4(x, 0, row, kb + 4 * g, v[g]);
          }); }
        // ---- event_conv.MLP.0 on src + relu(tgt + event) and tgt + relu(src + event) (:94-95,182-184)
        { const int acc[2] = {colZ, colZ + H};
          tc_linear<2>(L.g0, x, tab, nullptr, tmem, acc, [&](int c, int kcols) {
              if (kb < kcols) {
                  float sv[16], gv[16];
#pragma unroll
                  for (int k = 0; k < 16; k += 4) {       // issue the gathers first (explainer.py:348-351)
                      const int j = c * kKC + kb + k;
                      if (d_vec && j + 3 < D) {
                          const float4 s4 = s_ok ? ldg4(sf + j) : make_float4(0.f, 0.f, 0.f, 0.f), g4 = t_ok ? ldg4(tf + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                          sv[k] = s4.x; sv[k + 1] = s4.y; sv[k + 2] = s4.z; sv[k + 3] = s4.w; gv[k] = g4.x; gv[k + 1] = g4.y; gv[k + 2] = g4.z; gv[k + 3] = g4.w;
                      } else {
#pragma unroll
                          for (int i = 0; i < 4; ++i) { sv[k + i] = (j + i < D && s_ok) ? __ldg(sf + j + i) : 0.f; gv[k + i] = (j + i < D && t_ok) ? __ldg(tf + j + i) : 0.f; }
                      }
                  }
                  float ev[16];
                  tc::tmem_ld16(tmem + lane_base + colE + c * kKC + kb, ev);
#pragma unroll
                  for (int k = 0; k < 16; k += 4) {
                      if (kb + k >= kcols) break;
                      float hs[4], hg[4];
#pragma unroll
                      for (int i = 0; i < 4; ++i) {
                          const int j = c * kKC + kb + k + i;
                          const float e_ = j < D ? ev[k + i] + cst[L.evt.b + j] : 0.f;
                          hs[i] = j < D ? sv[k + i] + fmaxf(gv[k + i] + e_, 0.f) : 0.f;
                          hg[i] = j < D ? gv[k + i] + fmaxf(sv[k + i] + e_, 0.f) : 0.f;
                      }
                      store_a4(x, 0, row, kb + k, make_float4(hs[0], hs[1], hs[2], hs[3]));
                      store_a4(x, 1, row, kb + k, make_float4(hg[0], hg[1], hg[2], hg[3]));
                  }
              }
          }); }
        // ---- event_conv.MLP.2 (:84)
        { const int acc[2] = {colF, colF + H};
          tc_linear<2>(L.g2, x, tab, nullptr, tmem, acc, [&](int c, int kcols) {
              (void)kcols;
#pragma unroll
              for (int mb = 0; mb < 2; ++mb) {
                  float z[16];
                  tc::tmem_ld16(tmem + lane_base + colZ + mb * H + c * kKC + kb, z);
#pragma unroll
                  for (int k = 0; k < 16; k += 4) {
                      const float4 bb = lds4(cst + L.g0.b + c * kKC + kb + k);
                      store_a4(x, mb, row, kb + k, make_float4(fmaxf(z[k] + bb.x, 0.f), fmaxf(z[k + 1] + bb.y, 0.f), fmaxf(z[k + 2] + bb.z, 0.f), fmaxf(z[k + 3] + bb.w, 0.f)));
                  }
              }
          }); }
        // ---- updated_feature row = [MLP(src side) | MLP(tgt side)] (:185): half h owns columns [h*H, h*H + H), i.e.
        //      column chunks 2h and 2h+1 of the (motif tile, position) slabs
        {
            const int mrow = (int)(ml & 127);
            float *fo = a.F + (((ml >> 7) * 3 + pos) * 4 + 2 * half) * kSlabFloats;
            for (int c0 = 0; c0 < H; c0 += 16) {
                float v[16];
                tc::tmem_ld16(tmem + lane_base + colF + half * H + c0, v);
                if (live) {
                    float *fc = fo + (c0 >> 5) * kSlabFloats;
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        const float4 bb = lds4(cst + L.g2.b + c0 + i);
                        *reinterpret_cast<float4 *>(fc + slab_off(mrow, (c0 & 31) + i)) = make_float4(v[i] + bb.x, v[i + 1] + bb.y, v[i + 2] + bb.z, v[i + 3] + bb.w);
                    }
                }
            }
        }
        tc::fence_before_sync();
        __syncthreads();            // all TMEM reads of this tile done before the next tile's MMAs overwrite it
        tc::fence_after_sync();
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, a.tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// motif kernel: 256 threads = 128 motifs x 2 column halves.  TMEM: X = [0, 2H), Y = [2H, 4H).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads)
motif_tc_kernel(const TcLayout L, const ChunkTab tab, const StageTab stab, const float *__restrict__ blob0, const TcArgs a) {
    const float *__restrict__ blob = blob0 + (int64_t)(blockIdx.x % a.replicas) * ((L.total + 31) & ~(int64_t)31);
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[4];
    __shared__ uint32_t tmem_slot;
    __shared__ float part[2][128];
    TcCtx x;
    tc_setup(smem, x, L, a, 1, true, bars, &tmem_slot, blob);
    const int t = threadIdx.x, warp = t >> 5, row = t & 127, half = t >> 7, kb = 16 * half;
    const uint32_t tmem = tmem_slot, lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int H = L.H, H2 = 2 * L.H;
    const int colX = 0, colY = H2, colA1 = 0, colA2 = H, colM0 = H2, colM1 = 0;
    const int64_t n_m = min(a.slab, a.n_motifs - a.m_begin);
    const int64_t n_tiles = (n_m + 127) / 128;
    x.total = ((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) * tab.n;
    if (t == 0 && x.total > 0) { tc_prefetch_b(x, tab, 0); tc_prefetch_stage(x, tab, stab, 0); }
    const float *cst = x.cst;
    const float *sg0 = x.stage, *sg1 = x.stage + kSlabFloats;
    auto both_halves = [&](float v) {      // sum of the two column-half partials of every row
        part[half][row] = v;
        __syncthreads();
        const float s_ = part[0][row] + part[1][row];
        __syncthreads();
        return s_;
    };

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t ml = tile * 128 + row;         // motif index inside the slab
        const bool live = ml < n_m;
        const int64_t gm = a.m_begin + (live ? ml : 0);
        // dot over this thread's column half of (X + b1) . (Y + b2)
        auto score_half = [&]() {
            float sc = 0.f;
            for (int c0 = half * H; c0 < half * H + H; c0 += 16) {
                float p[16], q[16];
                tc::tmem_ld16(tmem + lane_base + colX + c0, p); tc::tmem_ld16(tmem + lane_base + colY + c0, q);
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 b1 = lds4(cst + L.w1.b + c0 + i), b2 = lds4(cst + L.w2.b + c0 + i);
                    sc = fmaf(p[i] + b1.x, q[i] + b2.x, sc); sc = fmaf(p[i + 1] + b1.y, q[i + 1] + b2.y, sc);
                    sc = fmaf(p[i + 2] + b1.z, q[i + 2] + b2.z, sc); sc = fmaf(p[i + 3] + b1.w, q[i + 3] + b2.w, sc);
                }
            }
            return sc;
        };
        auto copy_slab = [&](int c, int kcols) { (void)c; (void)kcols;      // A row = the staged updated_feature columns
#pragma unroll
            for (int k = kb; k < kb + 16; k += 4) store_a4(x, 0, row, k, lds4(sg0 + slab_off(row, k))); };
        // ---- Wp = W1 f2 -> X ; Wq_0 = W2 f_0 -> Y ; score_0 ; Wq_1 = W2 f_1 -> Y ; score_1 (:806-808)
        { const int acc[1] = {colX}; tc_linear<1>(L.w1, x, tab, &stab, tmem, acc, copy_slab); }
        { const int acc[1] = {colY}; tc_linear<1>(L.w2, x, tab, &stab, tmem, acc, copy_slab); }
        float s0 = both_halves(score_half());
        tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();        // Y is about to be overwritten
        { const int acc[1] = {colY}; tc_linear<1>(L.w2, x, tab, &stab, tmem, acc, copy_slab); }
        float s1 = both_halves(score_half());
        tc::fence_before_sync(); __syncthreads(); tc::fence_after_sync();
        // ---- temporal weighting + softmax (:811-839)
        if (L.use_temporal && live) {
            const int64_t b = gm / a.W;
            const float cut = a.cut[b], sd = __fadd_rn(a.std_[b / a.group], 1e-6f);
            const float d0 = fabsf(__fsub_rn(cut, a.t[gm * 3 + 0])), d1 = fabsf(__fsub_rn(cut, a.t[gm * 3 + 1]));
            s0 = __fmul_rn(s0, __fadd_rn(0.7f, __fmul_rn(0.3f, expf(__fdiv_rn(-d0, sd)))));      // :828,836
            s1 = __fmul_rn(s1, __fadd_rn(0.7f, __fmul_rn(0.3f, expf(__fdiv_rn(-d1, sd)))));
        }
        const float mx = fmaxf(s0, s1), e0 = expf(s0 - mx), e1 = expf(s1 - mx);
        const float al0 = e0 / (e0 + e1), al1 = e1 / (e0 + e1);
        // ---- sum_k alpha_k (W2 f_k + b2) = W2 (alpha_0 f_0 + alpha_1 f_1) + b2 since alpha sums to one -> Y (:841)
        { const int acc[1] = {colY};
          tc_linear<1>(L.w2, x, tab, &stab, tmem, acc, [&](int c, int kcols) { (void)c; (void)kcols;
#pragma unroll
              for (int k = kb; k < kb + 16; k += 4) {
                  const float4 u = lds4(sg0 + slab_off(row, k)), v = lds4(sg1 + slab_off(row, k));
                  store_a4(x, 0, row, k, make_float4(fmaf(al0, u.x, al1 * v.x), fmaf(al0, u.y, al1 * v.y), fmaf(al0, u.z, al1 * v.z), fmaf(al0, u.w, al1 * v.w)));
              } }); }
        // ---- attention.MLP.0 on f2 + (Y + b2) -> A1 (:842-843)
        { const int acc[1] = {colA1};
          tc_linear<1>(L.a0, x, tab, &stab, tmem, acc, [&](int c, int kcols) { (void)kcols;
              float q[16];
              tc::tmem_ld16(tmem + lane_base + colY + c * kKC + kb, q);
#pragma unroll
              for (int k = 0; k < 16; k += 4) {
                  const float4 f = lds4(sg0 + slab_off(row, kb + k)), b2 = lds4(cst + L.w2.b + c * kKC + kb + k);
                  store_a4(x, 0, row, kb + k, make_float4(f.x + (q[k] + b2.x), f.y + (q[k + 1] + b2.y), f.z + (q[k + 2] + b2.z), f.w + (q[k + 3] + b2.w)));
              } }); }
        // ---- attention.MLP.3 -> A2
        { const int acc[1] = {colA2};
          tc_linear<1>(L.a3, x, tab, &stab, tmem, acc, [&](int c, int kcols) { (void)kcols;
              float z[16];
              tc::tmem_ld16(tmem + lane_base + colA1 + c * kKC + kb, z);
#pragma unroll
              for (int k = 0; k < 16; k += 4) {
                  const float4 bb = lds4(cst + L.a0.b + c * kKC + kb + k);
                  store_a4(x, 0, row, kb + k, make_float4(fmaxf(z[k] + bb.x, 0.f), fmaxf(z[k + 1] + bb.y, 0.f), fmaxf(z[k + 2] + bb.z, 0.f), fmaxf(z[k + 3] + bb.w, 0.f)));
              } }); }
        // ---- MLP.0 on [attention out | one-hot(category)] -> M0 (:195-200)
        const int cat = (L.if_cat && live && a.cat) ? (int)a.cat[gm] : -1;
        { const int acc[1] = {colM0};
          tc_linear<1>(L.m0, x, tab, &stab, tmem, acc, [&](int c, int kcols) {
              if (kb < kcols) {
                  float z[16];
                  tc::tmem_ld16(tmem + lane_base + colA2 + min(c * kKC + kb, H - 16), z);   // columns >= H come from the one-hot
#pragma unroll
                  for (int k = 0; k < 16; k += 4) {
                      float v[4];
#pragma unroll
                      for (int i = 0; i < 4; ++i) {
                          const int j = c * kKC + kb + k + i;
                          v[i] = j < H ? z[k + i] + cst[L.a3.b + j] : (j - H == cat ? 1.f : 0.f);
                      }
                      store_a4(x, 0, row, kb + k, make_float4(v[0], v[1], v[2], v[3]));
                  }
              } }); }
        // ---- MLP.3 -> M1
        { const int acc[1] = {colM1};
          tc_linear<1>(L.m3, x, tab, &stab, tmem, acc, [&](int c, int kcols) {
              if (kb < kcols) {
                  float z[16];
                  tc::tmem_ld16(tmem + lane_base + colM0 + c * kKC + kb, z);
#pragma unroll
                  for (int k = 0; k < 16; k += 4) {
                      float v[4];
#pragma unroll
                      for (int i = 0; i < 4; ++i) { const int j = c * kKC + kb + k + i; v[i] = j < L.M ? fmaxf(z[k + i] + cst[L.m0.b + j], 0.f) : 0.f; }
                      store_a4(x, 0, row, kb + k, make_float4(v[0], v[1], v[2], v[3]));
                  }
              } }); }
        // ---- MLP.5 + sigmoid
        float z5 = 0.f;
        for (int c0 = half * (H / 2); c0 < half * (H / 2) + H / 2; c0 += 16) {
            float z[16];
            tc::tmem_ld16(tmem + lane_base + colM1 + c0, z);
#pragma unroll
            for (int i = 0; i < 16; ++i) z5 = fmaf(fmaxf(z[i] + cst[L.m3.b + c0 + i], 0.f), cst[L.w5 + c0 + i], z5);
        }
        z5 = both_halves(z5);
        if (live && half == 0) a.scores[gm] = 1.f / (1.f + expf(-(z5 + cst[L.b5])));
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, a.tmem_cols);
}


// ---------------------------------------------------------------------------------------------
// motif kernel, TS mode, warp specialised.  The A operand lives in TMEM (written with tcgen05.st by the thread
// that owns the row), so shared memory only carries the weight chunks and the staged updated_feature slabs
// (112 KB instead of 192 KB per K-chunk round).  One CTA per SM: warps 0-7 (256 threads = 128 motifs x 2 column
// halves) run the A-fills and the register-level epilogues; warp 8 is the issuer: it feeds the TMA queues (weight
// ring of 3, staging ring of 2) and issues the MMAs of chunk i as soon as the fill threads have arrived on
// a_full[i & 1].  The A region is double buffered, so fill(i+1) overlaps MMA(i); fill threads only wait when they
// need a buffer back (chunk i-2) or read an accumulator (layer boundary).
// TMEM: X [0,2H) Y [2H,4H) A0 [4H,4H+64) A1 [4H+64,4H+128).
// ---------------------------------------------------------------------------------------------
constexpr int kTsThreads = 288;
struct TsTab { int16_t n16[40], kcols[40], acc[40]; int8_t first[40]; };
// barriers: [0,1] MMA done per A buffer, [2,3,4] weight ring, [5,6] staging ring, [7,8] A full per buffer
__device__ __forceinline__ void named_sync_fill() { asm volatile("bar.sync 1, 256;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *mbar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(tc::smem_u32(mbar)) : "memory"); }

struct TsPipe {
    float *stage;                   // 2 x 2 slabs
    uint64_t *bars;
    int64_t seq, waited;            // chunk counter; all MMAs of chunks < waited are known complete (fill side)
    int ri;                         // round inside the tile
    int64_t su_use;                 // staging uses consumed
};

__device__ __forceinline__ void ts_wait_mma(TsPipe &p, int64_t upto) {        // fill threads: MMAs of chunks < upto complete
    for (; p.waited < upto; ++p.waited) tc::mbar_wait(p.bars + (p.waited & 1), (uint32_t)((p.waited >> 1) & 1));
    tc::fence_after_sync();
}

// fill(c, kcols, sg, v): this thread's 16 columns [c*kKC + kb, +16) of the A row into v[16]; sg = staged slabs of the round
template <typename Fill>
__device__ __forceinline__ void ts_linear(const TcLin l, TsPipe &p, const StageTab &st, int n_tab, uint32_t tmem, int colA, uint32_t lane_base, int kb, Fill fill) {
    const int nch = (l.K8 + kKC - 1) / kKC;
    for (int c = 0; c < nch; ++c) {
        const int kcols = min(kKC, l.K8 - c * kKC);
        const int64_t i = p.seq;
        if (i >= 2) ts_wait_mma(p, i - 1);                     // chunk i-2 done: A buffer (i & 1) is free
        const float *sg = p.stage;
        if (st.ns[p.ri]) {
            const int sb = (int)(p.su_use & 1);
            tc::mbar_wait(p.bars + 5 + sb, (uint32_t)((p.su_use >> 1) & 1));      // this round's slabs have landed
            sg = p.stage + sb * 2 * kSlabFloats;
            p.su_use++;
        }
        if (kb < kcols) {
            float v[16], h[16], lo[16];
            fill(c, kcols, sg, v);
#pragma unroll
            for (int k = 0; k < 16; ++k) tc::split_tf32(v[k], h[k], lo[k]);
            const uint32_t a0 = tmem + lane_base + colA + (uint32_t)(i & 1) * 64 + kb;
            tc::tmem_st16(a0, h);
            tc::tmem_st16(a0 + 32, lo);
            tc::tmem_st_wait();
        }
        tc::fence_before_sync();
        mbar_arrive(p.bars + 7 + (i & 1));                     // A buffer full (and the staged slabs consumed)
        p.seq++;
        p.ri = p.ri + 1 == n_tab ? 0 : p.ri + 1;
    }
}

__global__ void __launch_bounds__(kTsThreads, 1)
motif_ts_kernel(const TcLayout L, const ChunkTab tab, const StageTab stab, const TsTab tst, const float *__restrict__ blob0, const TcArgs a) {
    const float *__restrict__ blob = blob0 + (int64_t)(blockIdx.x % a.replicas) * ((L.total + 31) & ~(int64_t)31);
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[16];         // [0,1] MMA done, [5,6] staging, [7,8] A full, [10..13] weight ring
    __shared__ uint32_t tmem_slot;
    __shared__ float part[2][128];
    const int t = threadIdx.x, warp = t >> 5, row = t & 127, half = (t >> 7) & 1, kb = 16 * half;
    const uint32_t b_s = tc::smem_u32(smem), b_bytes = (uint32_t)a.b_bytes;
    float *stage = reinterpret_cast<float *>(smem + 4 * (size_t)a.b_bytes);
    float *cstw = stage + 4 * kSlabFloats;
    for (int i = t; i < L.n_cst; i += blockDim.x) cstw[i] = __ldg(blob + L.cst + i);
    if (t == 0) { for (int i = 0; i < 14; ++i) tc::mbar_init(bars + i, 1); tc::mbar_init(bars + 7, 256); tc::mbar_init(bars + 8, 256); }
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const float *cst = cstw;
    const uint32_t tmem = tmem_slot, lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int H = L.H, H2 = 2 * L.H;
    const int colX = 0, colY = H2, colA = 2 * H2, colA1 = 0, colA2 = H, colM0 = H2, colM1 = 0;
    const int64_t n_m = min(a.slab, a.n_motifs - a.m_begin);
    const int64_t n_tiles = (n_m + 127) / 128;
    const int64_t total = ((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) * tab.n;

    if (warp == 8) {
        // ================= issuer warp: TMA queues + MMA issue =================
        // weight ring of kRing buffers (chunk j lives in buffer j % kRing, barrier 10 + j % kRing), filled kRing - 1 chunks ahead;
        // all round-robin indices are carried incrementally (no 64-bit division in the loop)
        constexpr int kRing = 4;
        const int lane = t & 31;
        uint64_t *wbar = bars + 10;
        int b_ci = 0, b_buf = 0;                 // next weight chunk to request: index in the tile schedule, ring slot
        int64_t b_seq = 0;
        int s_ci = 0; int64_t s_seq = 0, s_tile = blockIdx.x, su_issue = 0;     // next staging request
        auto issue_b = [&]() {                   // whole warp keeps the counters; lane 0 talks to the TMA
            if (lane == 0) {
                tc::mbar_expect_tx(wbar + b_buf, (uint32_t)tab.bytes[b_ci]);
                tc::tma_load_1d_s(b_s + b_buf * b_bytes, blob + tab.off[b_ci], (uint32_t)tab.bytes[b_ci], wbar + b_buf);
            }
            __syncwarp();
            ++b_seq; b_ci = b_ci + 1 == tab.n ? 0 : b_ci + 1; b_buf = b_buf + 1 == kRing ? 0 : b_buf + 1;
        };
        auto issue_stage = [&]() {               // staged round number su_issue goes to buffer su_issue & 1
            const int ns = stab.ns[s_ci];
            if (ns) {
                const int buf = (int)(su_issue & 1);
                if (lane == 0) {
                    tc::mbar_expect_tx(bars + 5 + buf, (uint32_t)(ns * kSlabFloats * 4));
                    for (int k = 0; k < ns; ++k)
                        tc::tma_load_1d(stage + (buf * 2 + k) * kSlabFloats, a.F + ((s_tile * 3 + stab.pos[s_ci][k]) * 4 + stab.ch[s_ci][k]) * kSlabFloats, kSlabFloats * 4, bars + 5 + buf);
                }
                __syncwarp();
                su_issue++;
            }
            ++s_seq;
            if (++s_ci == tab.n) { s_ci = 0; s_tile += gridDim.x; }
        };
        for (int k = 0; k < kRing - 1 && b_seq < total; ++k) issue_b();       // prime the rings
        for (int k = 0; k < 2 && s_seq < total; ++k) issue_stage();
        int ci = 0, buf = 0;
        uint32_t wpar = 0;                       // parity of ring slot `buf`'s current use
        for (int64_t i = 0; i < total; ++i) {
            const bool tim = a.dbg && blockIdx.x == 0 && lane == 0 && i < 64;
            if (tim) a.dbg[i * 6 + 0] = clock64();
            tc::mbar_wait(bars + 7 + (i & 1), (uint32_t)((i >> 1) & 1));          // fill threads have written A(i) (and read their slabs)
            if (tim) a.dbg[i * 6 + 1] = clock64();
            if (s_seq < total) issue_stage();                                      // staging of round i+2: its buffer was consumed by round i or earlier
            tc::mbar_wait(wbar + buf, wpar);                                       // weight chunk has landed (TMA)
            if (tim) a.dbg[i * 6 + 2] = clock64();
            tc::fence_after_sync();
            const uint32_t leader = tc::elect_one();
            const int n16 = tst.n16[ci], kcols = tst.kcols[ci];
            const uint32_t idesc = tc::idesc_tf32(128, n16);
            const uint32_t lbo_b = (uint32_t)n16 * 16, b_base = b_s + (uint32_t)buf * b_bytes;
            uint64_t bh = tc::smem_desc(b_base, lbo_b, 128), bl = tc::smem_desc(b_base + (uint32_t)n16 * kKC * 4, lbo_b, 128);
            const uint64_t db = (2 * lbo_b) >> 4;
            const uint32_t a_hi = tmem + colA + (uint32_t)(i & 1) * 64, a_lo = a_hi + 32, d = tmem + tst.acc[ci];
            for (int ks = 0; ks < kcols / 8; ++ks) {
                tc::mma_tf32_ts(d, a_hi + 8 * ks, bh, idesc, (uint32_t)(!tst.first[ci] || ks != 0), leader);
                tc::mma_tf32_ts(d, a_lo + 8 * ks, bh, idesc, 1, leader);
                tc::mma_tf32_ts(d, a_hi + 8 * ks, bl, idesc, 1, leader);
                bh += db; bl += db;
            }
            tc::mma_commit(bars + (i & 1), leader);
            if (tim) a.dbg[i * 6 + 3] = clock64();
            __syncwarp();
            if (b_seq < total) {                                                    // ring slot of chunk i-1 is free once MMA(i-1) is done -> chunk i+kRing-1
                if (i >= 1) tc::mbar_wait(bars + ((i - 1) & 1), (uint32_t)(((i - 1) >> 1) & 1));
                if (tim) a.dbg[i * 6 + 4] = clock64();
                if (i >= 1) issue_b();
            }
            ci = ci + 1 == tab.n ? 0 : ci + 1;
            if (++buf == kRing) { buf = 0; wpar ^= 1; }
        }
    } else {
        // ================= fill / epilogue warps =================
        TsPipe p;
        p.stage = stage; p.bars = bars; p.seq = 0; p.waited = 0; p.ri = 0; p.su_use = 0;
        auto both_halves = [&](float v) {
            part[half][row] = v;
            named_sync_fill();
            const float s_ = part[0][row] + part[1][row];
            named_sync_fill();
            return s_;
        };
        auto drain = [&]() { ts_wait_mma(p, p.seq); };      // every issued MMA has completed: accumulators readable
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t ml = tile * 128 + row;
            const bool live = ml < n_m;
            const int64_t gm = a.m_begin + (live ? ml : 0);
            auto score_half = [&]() {
                float sc = 0.f;
                for (int c0 = half * H; c0 < half * H + H; c0 += 16) {
                    float pp[16], q[16];
                    tc::tmem_ld16(tmem + lane_base + colX + c0, pp); tc::tmem_ld16(tmem + lane_base + colY + c0, q);
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        const float4 b1 = lds4(cst + L.w1.b + c0 + i), b2 = lds4(cst + L.w2.b + c0 + i);
                        sc = fmaf(pp[i] + b1.x, q[i] + b2.x, sc); sc = fmaf(pp[i + 1] + b1.y, q[i + 1] + b2.y, sc);
                        sc = fmaf(pp[i + 2] + b1.z, q[i + 2] + b2.z, sc); sc = fmaf(pp[i + 3] + b1.w, q[i + 3] + b2.w, sc);
                    }
                }
                return sc;
            };
            auto copy_slab = [&](int c, int kcols, const float *sg, float *v) { (void)c; (void)kcols;
#pragma unroll
                for (int k = 0; k < 16; k += 4) { const float4 f = lds4(sg + slab_off(row, kb + k)); v[k] = f.x; v[k + 1] = f.y; v[k + 2] = f.z; v[k + 3] = f.w; } };
            // ---- Wp = W1 f2 -> X ; Wq_0 = W2 f_0 -> Y ; score_0 ; Wq_1 = W2 f_1 -> Y ; score_1 (:806-808)
            ts_linear(L.w1, p, stab, tab.n, tmem, colA, lane_base, kb, copy_slab);
            ts_linear(L.w2, p, stab, tab.n, tmem, colA, lane_base, kb, copy_slab);
            drain();
            float s0 = both_halves(score_half());
            tc::fence_before_sync(); named_sync_fill(); tc::fence_after_sync();        // Y is about to be overwritten
            ts_linear(L.w2, p, stab, tab.n, tmem, colA, lane_base, kb, copy_slab);
            drain();
            float s1 = both_halves(score_half());
            tc::fence_before_sync(); named_sync_fill(); tc::fence_after_sync();
            if (L.use_temporal && live) {                                             // temporal weighting (:811-836)
                const int64_t b = gm / a.W;
                const float cut = a.cut[b], sd = __fadd_rn(a.std_[b / a.group], 1e-6f);
                const float d0 = fabsf(__fsub_rn(cut, a.t[gm * 3 + 0])), d1 = fabsf(__fsub_rn(cut, a.t[gm * 3 + 1]));
                s0 = __fmul_rn(s0, __fadd_rn(0.7f, __fmul_rn(0.3f, expf(__fdiv_rn(-d0, sd)))));
                s1 = __fmul_rn(s1, __fadd_rn(0.7f, __fmul_rn(0.3f, expf(__fdiv_rn(-d1, sd)))));
            }
            const float mx = fmaxf(s0, s1), e0 = expf(s0 - mx), e1 = expf(s1 - mx);
            const float al0 = e0 / (e0 + e1), al1 = e1 / (e0 + e1);                  // softmax (:839)
            // ---- sum_k alpha_k (W2 f_k + b2) = W2 (alpha_0 f_0 + alpha_1 f_1) + b2 -> Y (:841)
            ts_linear(L.w2, p, stab, tab.n, tmem, colA, lane_base, kb, [&](int c, int kcols, const float *sg, float *v) { (void)c; (void)kcols;
#pragma unroll
                for (int k = 0; k < 16; k += 4) {
                    const float4 u = lds4(sg + slab_off(row, kb + k)), w = lds4(sg + kSlabFloats + slab_off(row, kb + k));
                    v[k] = fmaf(al0, u.x, al1 * w.x); v[k + 1] = fmaf(al0, u.y, al1 * w.y); v[k + 2] = fmaf(al0, u.z, al1 * w.z); v[k + 3] = fmaf(al0, u.w, al1 * w.w);
                } });
            drain();
            // ---- attention.MLP.0 on f2 + (Y + b2) -> A1 (:842-843)
            ts_linear(L.a0, p, stab, tab.n, tmem, colA, lane_base, kb, [&](int c, int kcols, const float *sg, float *v) { (void)kcols;
                float q[16];
                tc::tmem_ld16(tmem + lane_base + colY + c * kKC + kb, q);
#pragma unroll
                for (int k = 0; k < 16; k += 4) {
                    const float4 f = lds4(sg + slab_off(row, kb + k)), b2 = lds4(cst + L.w2.b + c * kKC + kb + k);
                    v[k] = f.x + (q[k] + b2.x); v[k + 1] = f.y + (q[k + 1] + b2.y); v[k + 2] = f.z + (q[k + 2] + b2.z); v[k + 3] = f.w + (q[k + 3] + b2.w);
                } });
            drain();
            // ---- attention.MLP.3 -> A2
            ts_linear(L.a3, p, stab, tab.n, tmem, colA, lane_base, kb, [&](int c, int kcols, const float *sg, float *v) { (void)kcols; (void)sg;
                float z[16];
                tc::tmem_ld16(tmem + lane_base + colA1 + c * kKC + kb, z);
#pragma unroll
                for (int k = 0; k < 16; k += 4) {
                    const float4 bb = lds4(cst + L.a0.b + c * kKC + kb + k);
                    v[k] = fmaxf(z[k] + bb.x, 0.f); v[k + 1] = fmaxf(z[k + 1] + bb.y, 0.f); v[k + 2] = fmaxf(z[k + 2] + bb.z, 0.f); v[k + 3] = fmaxf(z[k + 3] + bb.w, 0.f);
                } });
            drain();
            // ---- MLP.0 on [attention out | one-hot(category)] -> M0 (:195-200)
            const int cat = (L.if_cat && live && a.cat) ? (int)a.cat[gm] : -1;
            ts_linear(L.m0, p, stab, tab.n, tmem, colA, lane_base, kb, [&](int c, int kcols, const float *sg, float *v) { (void)kcols; (void)sg;
                float z[16];
                tc::tmem_ld16(tmem + lane_base + colA2 + min(c * kKC + kb, H - 16), z);   // columns >= H come from the one-hot
#pragma unroll
                for (int k = 0; k < 16; ++k) { const int j = c * kKC + kb + k; v[k] = j < H ? z[k] + cst[L.a3.b + j] : (j - H == cat ? 1.f : 0.f); } });
            drain();
            // ---- MLP.3 -> M1
            ts_linear(L.m3, p, stab, tab.n, tmem, colA, lane_base, kb, [&](int c, int kcols, const float *sg, float *v) { (void)kcols; (void)sg;
                float z[16];
                tc::tmem_ld16(tmem + lane_base + colM0 + c * kKC + kb, z);
#pragma unroll
                for (int k = 0; k < 16; ++k) { const int j = c * kKC + kb + k; v[k] = j < L.M ? fmaxf(z[k] + cst[L.m0.b + j], 0.f) : 0.f; } });
            drain();
            // ---- MLP.5 + sigmoid
            float z5 = 0.f;
            for (int c0 = half * (H / 2); c0 < half * (H / 2) + H / 2; c0 += 16) {
                float z[16];
                tc::tmem_ld16(tmem + lane_base + colM1 + c0, z);
#pragma unroll
                for (int i = 0; i < 16; ++i) z5 = fmaf(fmaxf(z[i] + cst[L.m3.b + c0 + i], 0.f), cst[L.w5 + c0 + i], z5);
            }
            z5 = both_halves(z5);
            if (live && half == 0) a.scores[gm] = 1.f / (1.f + expf(-(z5 + cst[L.b5])));
            tc::fence_before_sync();
            named_sync_fill();          // all TMEM reads of this tile done before the next tile's MMAs overwrite X
            tc::fence_after_sync();
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}


// ---------------------------------------------------------------------------------------------
// event kernel, TS mode, warp specialised: 16 fill warps (512 threads = 128 event rows x 4 column quarters) + one
// issuer warp, one CTA per SM.  Both orientations' A operands live in TMEM, double buffered:
// TMEM: Z [0,2H)  E/F [2H,4H)  A buffers [4H + 128*b + 64*mb, +64) = hi 32 | lo 32 columns.
// ---------------------------------------------------------------------------------------------
constexpr int kEvThreads = 544;
__device__ __forceinline__ void named_sync_fill512() { asm volatile("bar.sync 1, 512;\n" ::: "memory"); }

struct EvPipe { uint64_t *bars; int64_t seq, waited; };
__device__ __forceinline__ void ev_wait_mma(EvPipe &p, int64_t upto) {
    for (; p.waited < upto; ++p.waited) tc::mbar_wait(p.bars + (p.waited & 1), (uint32_t)((p.waited >> 1) & 1));
    tc::fence_after_sync();
}
// fill(c, kcols, v0, v1): this thread's 8 columns [c*kKC + kq, +8) of the A rows of the MB m-blocks
template <int MB, typename Fill>
__device__ __forceinline__ void ev_linear(const TcLin l, EvPipe &p, uint32_t tmem, int colA, uint32_t lane_base, int kq, Fill fill) {
    const int nch = (l.K8 + kKC - 1) / kKC;
    for (int c = 0; c < nch; ++c) {
        const int kcols = min(kKC, l.K8 - c * kKC);
        const int64_t i = p.seq;
        if (i >= 2) ev_wait_mma(p, i - 1);                     // chunk i-2 done: A buffer (i & 1) is free
        if (kq < kcols) {
            float v[MB][8], h[8], lo[8];
            fill(c, kcols, v);
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
#pragma unroll
                for (int k = 0; k < 8; ++k) tc::split_tf32(v[mb][k], h[k], lo[k]);
                const uint32_t a0 = tmem + lane_base + colA + (uint32_t)(i & 1) * 128 + mb * 64 + kq;
                tc::tmem_st8(a0, h);
                tc::tmem_st8(a0 + 32, lo);
            }
            tc::tmem_st_wait();
        }
        tc::fence_before_sync();
        mbar_arrive(p.bars + 5 + (i & 1));
        p.seq++;
    }
}

__global__ void __launch_bounds__(kEvThreads, 1)
event_ts_kernel(const TcLayout L, const ChunkTab tab, const TsTab tst, const float *__restrict__ blob0, const TcArgs a) {
    const float *__restrict__ blob = blob0 + (int64_t)(blockIdx.x % a.replicas) * ((L.total + 31) & ~(int64_t)31);
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[8];          // [0,1] MMA done per A buffer, [2,3,4] weight ring, [5,6] A full per buffer
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5, row = t & 127, quarter = (t >> 7) & 3, kq = 8 * quarter;
    const uint32_t b_s = tc::smem_u32(smem), b_bytes = (uint32_t)a.b_bytes;
    float *cstw = reinterpret_cast<float *>(smem + 3 * (size_t)a.b_bytes);
    for (int i = t; i < L.n_cst; i += blockDim.x) cstw[i] = __ldg(blob + L.cst + i);
    if (t == 0) { for (int i = 0; i < 5; ++i) tc::mbar_init(bars + i, 1); tc::mbar_init(bars + 5, 512); tc::mbar_init(bars + 6, 512); }
    if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const float *cst = cstw;
    const uint32_t tmem = tmem_slot, lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int H = L.H, D = L.D, Ed = L.Ed;
    const int colZ = 0, colE = 2 * H, colF = 2 * H, colA = 4 * H;
    const int64_t n_rows = 3 * min(a.slab, a.n_motifs - a.m_begin);
    const int64_t n_tiles = (n_rows + 127) / 128;
    const int64_t total = ((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) * tab.n;

    if (warp == 16) {
        // ================= issuer warp =================
        const int lane = t & 31;
        auto issue_b = [&](int64_t seq) {
            const int ci = (int)(seq % tab.n), buf = (int)(seq % 3);
            tc::mbar_expect_tx(bars + 2 + buf, (uint32_t)tab.bytes[ci]);
            tc::tma_load_1d_s(b_s + buf * b_bytes, blob + tab.off[ci], (uint32_t)tab.bytes[ci], bars + 2 + buf);
        };
        if (lane == 0 && total > 0) { issue_b(0); if (total > 1) issue_b(1); }
        __syncwarp();
        for (int64_t i = 0; i < total; ++i) {
            const int ci = (int)(i % tab.n), buf = (int)(i % 3);
            tc::mbar_wait(bars + 5 + (i & 1), (uint32_t)((i >> 1) & 1));          // A(i) written
            tc::mbar_wait(bars + 2 + buf, (uint32_t)((i / 3) & 1));               // weight chunk landed
            tc::fence_after_sync();
            const uint32_t leader = tc::elect_one();
            const int n16 = tst.n16[ci], kcols = tst.kcols[ci], nmb = tst.first[ci] >> 1;     // first: bit 0 = first chunk of the layer, bits 1.. = m-blocks
            const uint32_t idesc = tc::idesc_tf32(128, n16);
            const uint32_t lbo_b = (uint32_t)n16 * 16, b_base = b_s + (uint32_t)buf * b_bytes;
            uint64_t bh = tc::smem_desc(b_base, lbo_b, 128), bl = tc::smem_desc(b_base + (uint32_t)n16 * kKC * 4, lbo_b, 128);
            const uint64_t db = (2 * lbo_b) >> 4;
            const uint32_t acc_flag = (uint32_t)(!(tst.first[ci] & 1));
            for (int ks = 0; ks < kcols / 8; ++ks) {
                for (int mb = 0; mb < nmb; ++mb) {
                    const uint32_t a_hi = tmem + colA + (uint32_t)(i & 1) * 128 + mb * 64 + 8 * ks, d = tmem + tst.acc[ci] + mb * H;
                    tc::mma_tf32_ts(d, a_hi, bh, idesc, acc_flag | (uint32_t)(ks != 0), leader);
                    tc::mma_tf32_ts(d, a_hi + 32, bh, idesc, 1, leader);
                    tc::mma_tf32_ts(d, a_hi, bl, idesc, 1, leader);
                }
                bh += db; bl += db;
            }
            tc::mma_commit(bars + (i & 1), leader);
            __syncwarp();
            if (i + 2 < total) {
                if (i >= 1) tc::mbar_wait(bars + ((i - 1) & 1), (uint32_t)(((i - 1) >> 1) & 1));
                if (lane == 0) issue_b(i + 2);
                __syncwarp();
            }
        }
    } else {
        // ================= fill / epilogue warps =================
        EvPipe p; p.bars = bars; p.seq = 0; p.waited = 0;
        const bool ed_vec = (Ed & 3) == 0, d_vec = (D & 3) == 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int64_t r = tile * 128 + row;
            const bool live = r < n_rows;
            const int64_t ml = live ? r / 3 : 0, gm = a.m_begin + ml;
            const int pos = live ? (int)(r - 3 * ml) : 0;
            int64_t e = 0, ns = 0, nt = 0; float dt = 0.f;
            if (live) {
                e = a.eidx[gm * 3 + pos]; ns = a.nodes[gm * 6 + 2 * pos]; nt = a.nodes[gm * 6 + 2 * pos + 1];
                dt = __fsub_rn(a.t[gm * 3 + 2], a.t[gm * 3 + pos]);                       // explainer.py:326
            }
            const bool e_ok = live && e >= 0 && e < a.n_edge_rows, s_ok = live && ns >= 0 && ns < a.n_node_rows, t_ok = live && nt >= 0 && nt < a.n_node_rows;
            const float *ef = a.edge_feat + e * Ed, *sf = a.node_feat + ns * D, *tf = a.node_feat + nt * D;
            const float *ei = a.eid ? a.eid + gm * 9 + pos * 3 : nullptr;
            auto xval = [&](int j) -> float {                                              // event_features column j (:179)
                if (j < Ed) return e_ok ? __ldg(ef + j) : 0.f;
                if (j < Ed + 3) return (live && ei) ? __ldg(ei + (j - Ed)) : 0.f;
                if (j < L.ev) { const int k = j - Ed - 3; return live ? cos_accurate(__fadd_rn(__fmul_rn(dt, cst[L.freq + k]), cst[L.phase + k])) : 0.f; }   // :55-58
                return 0.f;
            };
            // ---- lin_event (:93) -> E
            ev_linear<1>(L.evt, p, tmem, colA, lane_base, kq, [&](int c, int kcols, float (*v)[8]) { (void)kcols;
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const int j = c * kKC + kq + 4 * g;
                    float4 x;
                    if (ed_vec && j + 3 < Ed) x = e_ok ? ldg4(ef + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    else x = make_float4(xval(j), xval(j + 1), xval(j + 2), xval(j + 3));
                    v[0][4 * g] = x.x; v[0][4 * g + 1] = x.y; v[0][4 * g + 2] = x.z; v[0][4 * g + 3] = x.w;
                } });
            ev_wait_mma(p, p.seq);                                   // E complete
            // ---- event_conv.MLP.0 on src + relu(tgt + event) and tgt + relu(src + event) (:94-95,182-184) -> Z
            ev_linear<2>(L.g0, p, tmem, colA, lane_base, kq, [&](int c, int kcols, float (*v)[8]) { (void)kcols;
                float sv[8], gv[8], ev[8];
#pragma unroll
                for (int k = 0; k < 8; k += 4) {
                    const int j = c * kKC + kq + k;
                    if (d_vec && j + 3 < D) {
                        const float4 s4 = s_ok ? ldg4(sf + j) : make_float4(0.f, 0.f, 0.f, 0.f), g4 = t_ok ? ldg4(tf + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                        sv[k] = s4.x; sv[k + 1] = s4.y; sv[k + 2] = s4.z; sv[k + 3] = s4.w; gv[k] = g4.x; gv[k + 1] = g4.y; gv[k + 2] = g4.z; gv[k + 3] = g4.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) { sv[k + i] = (j + i < D && s_ok) ? __ldg(sf + j + i) : 0.f; gv[k + i] = (j + i < D && t_ok) ? __ldg(tf + j + i) : 0.f; }
                    }
                }
                tc::tmem_ld8(tmem + lane_base + colE + c * kKC + kq, ev);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int j = c * kKC + kq + k;
                    const float e_ = j < D ? ev[k] + cst[L.evt.b + j] : 0.f;
                    v[0][k] = j < D ? sv[k] + fmaxf(gv[k] + e_, 0.f) : 0.f;
                    v[1][k] = j < D ? gv[k] + fmaxf(sv[k] + e_, 0.f) : 0.f;
                } });
            ev_wait_mma(p, p.seq);                                   // Z complete
            // ---- event_conv.MLP.2 (:84) -> F (aliases E, dead)
            ev_linear<2>(L.g2, p, tmem, colA, lane_base, kq, [&](int c, int kcols, float (*v)[8]) { (void)kcols;
#pragma unroll
                for (int mb = 0; mb < 2; ++mb) {
                    float z[8];
                    tc::tmem_ld8(tmem + lane_base + colZ + mb * H + c * kKC + kq, z);
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[mb][k] = fmaxf(z[k] + cst[L.g0.b + c * kKC + kq + k], 0.f);
                } });
            ev_wait_mma(p, p.seq);                                   // F complete
            // ---- updated_feature row (:185): quarter q owns columns [32q, 32q + 32) = column chunk q of the slabs
            {
                const int mrow = (int)(ml & 127);
                float *fc = a.F + (((ml >> 7) * 3 + pos) * 4 + quarter) * kSlabFloats;
                for (int c0 = 0; c0 < 32; c0 += 8) {
                    float v[8];
                    tc::tmem_ld8(tmem + lane_base + colF + quarter * 32 + c0, v);
                    if (live) {
                        const float *bb = cst + L.g2.b + ((quarter * 32 + c0) & (H - 1));
                        *reinterpret_cast<float4 *>(fc + slab_off(mrow, c0)) = make_float4(v[0] + bb[0], v[1] + bb[1], v[2] + bb[2], v[3] + bb[3]);
                        *reinterpret_cast<float4 *>(fc + slab_off(mrow, c0 + 4)) = make_float4(v[4] + bb[4], v[5] + bb[5], v[6] + bb[6], v[7] + bb[7]);
                    }
                }
            }
            tc::fence_before_sync();
            named_sync_fill512();       // all TMEM reads of this tile done before the next tile's MMAs overwrite E
            tc::fence_after_sync();
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace tmb

using namespace tmb;

namespace tmb {

int64_t tc_blob_floats(const tm_encoder_desc &d) { return ((make_tc_layout(d).total + 31) & ~(int64_t)31) * kReplicas; }

int tc_pack(const tm_encoder_desc &d, const tm_encoder_params &p, float *blob) {
    const TcLayout L = make_tc_layout(d);
    memset(blob, 0, sizeof(float) * ((L.total + 31) & ~(int64_t)31) * kReplicas);
    const int H = L.H, D = L.D, M = L.M;
    pack_tc_lin(L, L.evt, L.ev, D, p.lin_event_w, p.lin_event_b, blob);
    pack_tc_lin(L, L.g0, D, H, p.gcn0_w, p.gcn0_b, blob);
    pack_tc_lin(L, L.g2, H, H, p.gcn2_w, p.gcn2_b, blob);
    pack_tc_lin(L, L.w1, 2 * H, 2 * H, p.att_w1_w, p.att_w1_b, blob);
    pack_tc_lin(L, L.w2, 2 * H, 2 * H, p.att_w2_w, p.att_w2_b, blob);
    pack_tc_lin(L, L.a0, 2 * H, H, p.att_mlp0_w, p.att_mlp0_b, blob);
    pack_tc_lin(L, L.a3, H, H, p.att_mlp3_w, p.att_mlp3_b, blob);
    pack_tc_lin(L, L.m0, M, M, p.mlp0_w, p.mlp0_b, blob);
    pack_tc_lin(L, L.m3, M, H, p.mlp3_w, p.mlp3_b, blob);
    float *cst = blob + L.cst;
    for (int k = 0; k < H; ++k) cst[L.w5 + k] = p.mlp5_w[k];
    cst[L.b5] = p.mlp5_b[0];
    for (int k = 0; k < D; ++k) { cst[L.freq + k] = p.basis_freq[k]; cst[L.phase + k] = p.phase[k]; }
    const int64_t stride = (L.total + 31) & ~(int64_t)31;
    for (int r = 1; r < kReplicas; ++r) memcpy(blob + r * stride, blob, sizeof(float) * L.total);
    return TM_OK;
}

// Motifs per slab: a whole number of waves of the motif kernel (2 CTAs per SM x 128 motifs), small enough that the
// slab's updated_feature rows (1.5 KB per motif) stay in the 126 MB L2 between the event and the motif kernel.
int64_t tc_slab_motifs() {
    static int64_t v = 0;
    if (!v) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const char *e = getenv("TEMPME_TC_SLAB_WAVES");
        const int waves = e ? std::max(1, atoi(e)) : 1;
        v = (int64_t)sms * 2 * 128 * waves;
    }
    return v;
}

// optional per-kernel timing (bench.py roofline): CUDA events around every launch of the two kernels
static bool g_prof = false;
static std::vector<cudaEvent_t> g_prof_ev;      // triples: before event kernel, between, after motif kernel
static size_t g_prof_used = 0;

// std_ = per-batch std (already computed); F = workspace for one slab
int tc_encode_score(const tm_encoder_desc &d, const float *d_blob_tc, int64_t B, int64_t W, int64_t group, const int32_t *nodes,
                    const int32_t *eidx, const float *t, const uint8_t *cat, const float *cut, const float *eid, const float *node_feat,
                    int64_t n_node_rows, const float *edge_feat, int64_t n_edge_rows, const float *std_, float *F, float *scores,
                    int device, cudaStream_t st) {
    const TcLayout L = make_tc_layout(d);
    if (L.H != 64) { set_error("tc_encode_score: hid_dim must be 64"); return TM_ERR_UNSUPPORTED; }
    ChunkTab te, tm;
    StageTab stg;
    memset(&stg, 0, sizeof stg);
    te.n = tm.n = 0;
    auto push = [&](ChunkTab &tab, const TcLin &l, int npos, int p0, int p1) {
        const int nch = (l.K8 + kKC - 1) / kKC;
        for (int c = 0; c < nch; ++c) {
            tab.off[tab.n] = l.w + (int64_t)c * 2 * l.N16 * kKC; tab.bytes[tab.n] = 2 * l.N16 * kKC * 4;
            if (&tab == &tm) { stg.ns[tab.n] = (int8_t)npos; stg.pos[tab.n][0] = (int8_t)p0; stg.pos[tab.n][1] = (int8_t)p1; stg.ch[tab.n][0] = stg.ch[tab.n][1] = (int8_t)c; }
            tab.n++;
        }
    };
    const int n_evt = (L.evt.K8 + kKC - 1) / kKC + (L.g0.K8 + kKC - 1) / kKC + 2;
    if (n_evt > 40) { set_error("tc_encode_score: feature dims need more than 40 weight chunks"); return TM_ERR_UNSUPPORTED; }
    push(te, L.evt, 0, 0, 0); push(te, L.g0, 0, 0, 0); push(te, L.g2, 0, 0, 0);
    push(tm, L.w1, 1, 2, 0); push(tm, L.w2, 1, 0, 0); push(tm, L.w2, 1, 1, 0); push(tm, L.w2, 2, 0, 1); push(tm, L.a0, 1, 2, 0);
    push(tm, L.a3, 0, 0, 0); push(tm, L.m0, 0, 0, 0); push(tm, L.m3, 0, 0, 0);
    const int bb_e = 2 * std::max(r16(L.D), L.H) * kKC * 4, bb_m = 2 * std::max(2 * L.H, r16(L.M)) * kKC * 4;
    const size_t cst_b = (size_t)((L.n_cst + 31) & ~31) * 4;
    const int nbuf_e = 2, nbuf_m = 1;
    const size_t smem_e = (size_t)4 * 128 * kKC * 4 + (size_t)nbuf_e * bb_e + cst_b;
    const size_t smem_m = (size_t)2 * 128 * kKC * 4 + (size_t)nbuf_m * bb_m + (size_t)2 * kSlabFloats * 4 + cst_b;
    TsTab tst;
    memset(&tst, 0, sizeof tst);
    {
        const int H2 = 2 * L.H;
        int k = 0;
        auto tsp = [&](const TcLin &l, int acc) {
            const int nch = (l.K8 + kKC - 1) / kKC;
            for (int c = 0; c < nch && k < 40; ++c, ++k) { tst.n16[k] = (int16_t)l.N16; tst.kcols[k] = (int16_t)std::min(kKC, l.K8 - c * kKC); tst.acc[k] = (int16_t)acc; tst.first[k] = (int8_t)(c == 0); }
        };
        tsp(L.w1, 0); tsp(L.w2, H2); tsp(L.w2, H2); tsp(L.w2, H2); tsp(L.a0, 0); tsp(L.a3, L.H); tsp(L.m0, H2); tsp(L.m3, 0);
    }
    TsTab tse;                                                // event kernel chunk table: first = (first chunk of layer) | m-blocks << 1
    memset(&tse, 0, sizeof tse);
    {
        int k = 0;
        auto tsp = [&](const TcLin &l, int acc, int nmb) {
            const int nch = (l.K8 + kKC - 1) / kKC;
            for (int c = 0; c < nch && k < 40; ++c, ++k) { tse.n16[k] = (int16_t)l.N16; tse.kcols[k] = (int16_t)std::min(kKC, l.K8 - c * kKC); tse.acc[k] = (int16_t)acc; tse.first[k] = (int8_t)((c == 0) | (nmb << 1)); }
        };
        tsp(L.evt, 2 * L.H, 1); tsp(L.g0, 0, 2); tsp(L.g2, 2 * L.H, 2);
    }
    const char *es_env = getenv("TEMPME_TC_EVENT");           // "ts": A operand in TMEM, warp-specialised, 1 CTA/SM (needs D <= 128); default: shared-memory operands, 2 CTAs/SM
    const bool use_ts_e = es_env && strcmp(es_env, "ts") == 0 && r16(L.D) <= 2 * L.H;
    const size_t smem_ts_e = (size_t)3 * bb_e + cst_b;
    const char *ms_env = getenv("TEMPME_TC_MOTIF");          // "ts": A operand in TMEM, warp-specialised, 1 CTA/SM; default: both operands from shared memory, 2 CTAs/SM
    const bool use_ts = ms_env && strcmp(ms_env, "ts") == 0;
    const size_t smem_ts = (size_t)4 * bb_m + (size_t)4 * kSlabFloats * 4 + cst_b;
    uint32_t cols_e = 32, cols_m = 32;
    while ((int)cols_e < 2 * L.H + std::max(r16(L.D), 2 * L.H)) cols_e <<= 1;
    while ((int)cols_m < 4 * L.H) cols_m <<= 1;
    static bool attr_set[64] = {false};
    if (device < 64 && !attr_set[device]) {
        TM_CUDA(cudaFuncSetAttribute(event_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        TM_CUDA(cudaFuncSetAttribute(motif_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        TM_CUDA(cudaFuncSetAttribute(motif_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        TM_CUDA(cudaFuncSetAttribute(event_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set[device] = true;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    TcArgs a;
    a.n_motifs = B * W; a.W = W; a.group = group; a.slab = tc_slab_motifs(); a.nodes = nodes; a.eidx = eidx; a.t = t; a.cat = cat; a.cut = cut;
    a.eid = eid; a.node_feat = node_feat; a.edge_feat = edge_feat; a.std_ = std_; a.n_node_rows = n_node_rows; a.n_edge_rows = n_edge_rows;
    a.F = F; a.scores = scores; a.dbg = nullptr;
    { const char *re = getenv("TEMPME_TC_REPLICAS"); a.replicas = re ? std::min(kReplicas, std::max(1, atoi(re))) : kReplicas; }
    static long long *dbg_buf = nullptr;
    const char *tim_env = getenv("TEMPME_TC_TIMING");
    if (tim_env && !dbg_buf) cudaMalloc(&dbg_buf, (2 * 64 * 6 + 2 * 1024) * sizeof(long long));
    const int ctas_e = (cols_e <= 256 && smem_e <= 110 * 1024) ? 2 : 1, ctas_m = (cols_m <= 256 && smem_m <= 110 * 1024) ? 2 : 1;
    // Slabs alternate between two internal streams (each with its own F buffer) so that the ramp-down of one slab's
    // motif kernel overlaps the next slab's event kernel; both fork from / join the caller's stream through events.
    static cudaStream_t s2[64][2];
    static cudaEvent_t ev_fork[64], ev_join[64][2];
    static bool s2_init[64] = {false};
    if (device >= 64) { set_error("tc_encode_score: device index >= 64"); return TM_ERR_UNSUPPORTED; }
    if (!s2_init[device]) {
        for (int k = 0; k < 2; ++k) { TM_CUDA(cudaStreamCreateWithFlags(&s2[device][k], cudaStreamNonBlocking)); TM_CUDA(cudaEventCreateWithFlags(&ev_join[device][k], cudaEventDisableTiming)); }
        TM_CUDA(cudaEventCreateWithFlags(&ev_fork[device], cudaEventDisableTiming));
        s2_init[device] = true;
    }
    const bool two = a.n_motifs > a.slab && !tim_env && !getenv("TEMPME_TC_SERIAL");
    if (two) {
        TM_CUDA(cudaEventRecord(ev_fork[device], st));
        for (int k = 0; k < 2; ++k) TM_CUDA(cudaStreamWaitEvent(s2[device][k], ev_fork[device], 0));
    }
    const int64_t f_floats = ((a.slab + 127) / 128) * 128 * 3 * 2 * (int64_t)L.H;
    cudaStream_t caller = st;
    int64_t islab = 0;
    for (int64_t m0 = 0; m0 < a.n_motifs; m0 += a.slab, ++islab) {
        st = two ? s2[device][islab & 1] : caller;
        a.F = F + (two ? (islab & 1) * f_floats : 0);
        a.m_begin = m0;
        const int64_t nm = std::min(a.slab, a.n_motifs - m0);
        const int64_t tiles_e = (3 * nm + 127) / 128, tiles_m = (nm + 127) / 128;
        cudaEvent_t *pe = nullptr;
        if (g_prof) {
            if (g_prof_used + 3 <= g_prof_ev.size()) {        // pool is created by tm_encoder_profile(1); when exhausted, stop recording
                pe = &g_prof_ev[g_prof_used]; g_prof_used += 3;
                cudaEventRecord(pe[0], st);
            }
        }
        a.tmem_cols = cols_e; a.b_bytes = bb_e; a.nbuf = nbuf_e; a.dbg = (tim_env && m0 == 0) ? dbg_buf : nullptr;
        if (use_ts_e) event_ts_kernel<<<(unsigned)std::min<int64_t>(tiles_e, (int64_t)sms), kEvThreads, smem_ts_e, st>>>(L, te, tse, d_blob_tc, a);
        else
        event_tc_kernel<<<(unsigned)std::min<int64_t>(tiles_e, (int64_t)sms * ctas_e), kTcThreads, smem_e, st>>>(L, te, d_blob_tc, a);
        TM_LAUNCH_CHECK();
        if (pe) cudaEventRecord(pe[1], st);
        a.tmem_cols = cols_m; a.b_bytes = bb_m; a.nbuf = nbuf_m; a.dbg = (tim_env && m0 == 0) ? dbg_buf + 1024 : nullptr;
        if (use_ts) motif_ts_kernel<<<(unsigned)std::min<int64_t>(tiles_m, (int64_t)sms), kTsThreads, smem_ts, st>>>(L, tm, stg, tst, d_blob_tc, a);
        else
        motif_tc_kernel<<<(unsigned)std::min<int64_t>(tiles_m, (int64_t)sms * ctas_m), kTcThreads, smem_m, st>>>(L, tm, stg, d_blob_tc, a);
        TM_LAUNCH_CHECK();
        if (pe) cudaEventRecord(pe[2], st);
    }
    st = caller;
    if (two)
        for (int k = 0; k < 2; ++k) { TM_CUDA(cudaEventRecord(ev_join[device][k], s2[device][k])); TM_CUDA(cudaStreamWaitEvent(st, ev_join[device][k], 0)); }
    if (tim_env) {      // diagnostic only: dump the phase timeline of CTA 0 of the first slab
        cudaStreamSynchronize(st);
        std::vector<long long> h(2 * 64 * 6 + 2 * 1024);
        cudaMemcpy(h.data(), dbg_buf, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        for (int k = 0; k < 2; ++k) {
            const int n = std::min(k ? tm.n : te.n, 64);
            fprintf(stderr, "[tc timing] %s kernel, CTA 0, first tile: chunk: fill | sync | tma-wait | mma-issue | mma-done   (cycles)\n", k ? "motif" : "event");
            if (k && use_ts) {
                fprintf(stderr, "  (TS kernel, issuer warp) chunk: wait A-full | wait weights | issue+commit | wait MMA(i-1) | period\n");
                for (int c = 0; c < n; ++c) {
                    const long long *dd = h.data() + k * 1024 + c * 6, *dn = dd + 6;
                    fprintf(stderr, "  %2d: %6lld %6lld %6lld %6lld   %6lld\n", c, dd[1] - dd[0], dd[2] - dd[1], dd[3] - dd[2], dd[4] - dd[3], c + 1 < n ? dn[0] - dd[0] : 0);
                }
                continue;
            }
            for (int c = 0; c < n; ++c) {
                const long long *dd = h.data() + k * 1024 + c * 6, *d7 = h.data() + k * 1024 + 768 + c * 4;
                fprintf(stderr, "  %2d: %6lld %6lld %6lld %6lld %6lld   round %6lld | warp7: start+%lld fill %lld proxy-fence %lld bar %lld\n", c, dd[1] - dd[0], dd[2] - dd[1], dd[3] - dd[2], dd[4] - dd[3], dd[5] - dd[4], dd[5] - dd[0], d7[0] - dd[0], d7[1] - d7[0], d7[2] - d7[1], d7[3] - d7[2]);
            }
        }
    }
    return TM_OK;
}

}  // namespace tmb

extern "C" int tm_encoder_profile(int enable) {
    tmb::g_prof = enable != 0;
    tmb::g_prof_used = 0;
    if (enable && tmb::g_prof_ev.empty()) {          // event pool, created outside any timed region
        tmb::g_prof_ev.resize(3 * 4096);
        for (auto &e : tmb::g_prof_ev) TM_CUDA(cudaEventCreate(&e));
    }
    return TM_OK;
}

extern "C" int tm_encoder_profile_read(float *h_event_ms, float *h_motif_ms) {
    if (!h_event_ms || !h_motif_ms) { set_error("tm_encoder_profile_read: null output"); return TM_ERR_ARG; }
    double e = 0, m = 0;
    for (size_t i = 0; i + 2 < tmb::g_prof_used + 0 && i + 2 < tmb::g_prof_ev.size() + 0; i += 3) {
        float a = 0, b = 0;
        TM_CUDA(cudaEventSynchronize(tmb::g_prof_ev[i + 2]));
        TM_CUDA(cudaEventElapsedTime(&a, tmb::g_prof_ev[i], tmb::g_prof_ev[i + 1]));
        TM_CUDA(cudaEventElapsedTime(&b, tmb::g_prof_ev[i + 1], tmb::g_prof_ev[i + 2]));
        e += a; m += b;
    }
    *h_event_ms = (float)e; *h_motif_ms = (float)m;
    tmb::g_prof_used = 0;
    return TM_OK;
}
