// tcgen05 (5th-gen tensor core) scorer: TempME.forward (reference models/explainer.py:174-201) as two kernels of
// 3xTF32 GEMM rounds with TMEM accumulators.
//
// Algebra.  Between event_conv.MLP.0's ReLU and attention.MLP.0's ReLU the reference applies only linear maps and the
// two softmax weights, and between attention.MLP.0's ReLU and MLP.0's ReLU only linear maps and a one-hot, so those
// chains are folded on the host (float64, tc_pack) into single matrices.  With h_k = [relu(MLP.0(src side)) |
// relu(MLP.0(tgt side))] of event k (k = 2 is the event next to the root), G = blockdiag(MLP.2, MLP.2), g its bias:
//     Wp   = W1 (G h_2 + g) + b1 = A1 h_2 + c1            Wq_k = W2 (G h_k + g) + b2 = A2 h_k + c2      (:806-807)
//     s_k  = Wp . Wq_k = h_k . (S h_2 + cu) + (d . h_2 + e)      S = A2^T A1, cu = A2^T c1, d = A1^T c2, e = c1 . c2
//     y    = relu(attention.MLP.0(f_2 + sum_k alpha_k Wq_k)) = relu(P h_2 + Q (alpha_0 h_0 + alpha_1 h_1) + cy)   (:841-843)
//     m0   = relu(MLP.0([attention.MLP.3(y) | onehot(cat)])) = relu(R y + cm[cat])                                (:196-200)
// which halves the multiply-adds per motif and takes the per-tile GEMM rounds from 34 to 17 (D = Ed = 32).  The
// folded weights are rounded to fp32 once; scores agree with the unfolded fp32 evaluation to ~2e-7 relative.
//
// One persistent kernel (score_tc_kernel) takes tiles of 128 motifs through three event passes (lin_event over
// [edge features | TimeEncode]; the three edge-identity columns are added on the CUDA cores; position-2 rows have
// dt = 0, so their TimeEncode chunks collapse into a bias; event_conv.MLP.0 + ReLU for the two orientations) and the
// motif rounds ([S; P] h_2 -> scores -> temporal weights, softmax -> + Q mix -> R -> MLP.3 -> MLP.5 + sigmoid).
// Thread (row, half) of the 256 owns TMEM lane `row` and 16 of the 32 columns of every K chunk: it reads the previous
// accumulator row with tcgen05.ld, applies bias / ReLU / mixing in fp32 registers, splits into tf32 hi + lo and stores
// the K-major operand tile in shared memory; weights arrive pre-split and pre-tiled by 1-D bulk TMA.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "tc.cuh"

namespace tmb {

constexpr int kKC = 32;           // K columns per operand chunk
constexpr int kTcThreads = 256;   // 128 rows x 2 column halves
constexpr int kSlabFloats = 128 * kKC;
constexpr uint32_t kATile = 128 * kKC * 4;

struct TcLin { int64_t w; int K8, N16, nch; };   // chunk c at w + c * 2 * N16 * kKC floats: [hi tile | lo tile]
struct TcLayout {
    int D, Ed, H, M, use_temporal, if_cat, D16, M16;
    int nch_edge;                                        // lin_event chunks that hold an edge-feature column
    TcLin evt, g0, sp, q, r, m3;
    int e_b, e_b2, e_wi, e_g0b, e_freq, e_phase, n_cstE; // event-kernel constants (offsets inside cstE)
    int m_cu, m_d, m_e, m_cy, m_m3b, m_w5, m_b5, n_cstM; // motif-kernel constants (offsets inside cstM)
    int64_t cstE, cstM, cm, total;                       // cm: [12][M16] per-category bias of the folded MLP.0 (global, read through L1)
};

__host__ __device__ static inline int r8(int x) { return (x + 7) & ~7; }
__host__ __device__ static inline int r16(int x) { return (x + 15) & ~15; }
__host__ __device__ static inline int64_t chunk_floats(const TcLin &l) { return (int64_t)2 * l.N16 * kKC; }

TcLayout make_tc_layout(const tm_encoder_desc &d) {
    TcLayout L;
    memset(&L, 0, sizeof L);
    L.D = d.node_dim; L.Ed = d.edge_dim; L.H = d.hid_dim; L.use_temporal = d.use_temporal; L.if_cat = d.if_cat;
    L.M = d.if_cat ? d.hid_dim + 12 : d.hid_dim;
    L.D16 = r16(L.D); L.M16 = r16(L.M);
    L.nch_edge = (L.Ed + kKC - 1) / kKC;
    int64_t o = 0;
    auto lin = [&](int K, int N) {
        TcLin l; l.K8 = r8(K); l.N16 = r16(N); l.nch = (l.K8 + kKC - 1) / kKC;
        l.w = o; o += (int64_t)l.nch * chunk_floats(l);
        return l;
    };
    const int H = L.H;
    L.evt = lin(L.Ed + L.D, L.D); L.g0 = lin(L.D, H);
    L.sp = lin(2 * H, 3 * H); L.q = lin(2 * H, H); L.r = lin(H, L.M); L.m3 = lin(L.M, H);
    int c = 0;
    L.e_b = c; c += L.D16; L.e_b2 = c; c += L.D16; L.e_wi = c; c += 3 * L.D16; L.e_g0b = c; c += r16(H);
    L.e_freq = c; c += L.D16; L.e_phase = c; c += L.D16; L.n_cstE = c;
    c = 0;
    L.m_cu = c; c += 2 * H; L.m_d = c; c += 2 * H; L.m_e = c; c += 16; L.m_cy = c; c += H; L.m_m3b = c; c += H;
    L.m_w5 = c; c += H; L.m_b5 = c; c += 16; L.n_cstM = c;
    L.cstE = o; o += L.n_cstE; L.cstM = o; o += L.n_cstM; L.cm = o; o += 12 * L.M16;
    L.total = o;
    return L;
}

// host: weight [N][K] (double) -> per K chunk the [hi | lo] operand tiles in the tc.cuh layout (R = N16 rows)
static void pack_tc_lin(const TcLin &l, int K, int N, const double *w, float *blob) {
    for (int c = 0; c < l.nch; ++c) {
        float *hi = blob + l.w + (int64_t)c * chunk_floats(l), *lo = hi + (int64_t)l.N16 * kKC;
        for (int n = 0; n < N; ++n)
            for (int kk = 0; kk < kKC && c * kKC + kk < K; ++kk) {
                const float x = (float)w[(int64_t)n * K + c * kKC + kk];
                uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u;
                float h; memcpy(&h, &u, 4);
                const int64_t off = ((kk >> 2) * (l.N16 * 16) + (n >> 3) * 128 + (n & 7) * 16 + (kk & 3) * 4) / 4;
                hi[off] = h; lo[off] = x - h;
            }
    }
}

int64_t tc_blob_floats(const tm_encoder_desc &d) { return (make_tc_layout(d).total + 31) & ~(int64_t)31; }

namespace {
using Mat = std::vector<double>;
// C[n x m] = A[n x k] * B[k x m]  (row-major)
Mat mm(const Mat &A, const Mat &B, int n, int k, int m) {
    Mat C((size_t)n * m, 0.0);
    for (int i = 0; i < n; ++i)
        for (int x = 0; x < k; ++x) {
            const double a = A[(size_t)i * k + x];
            if (a == 0.0) continue;
            for (int j = 0; j < m; ++j) C[(size_t)i * m + j] += a * B[(size_t)x * m + j];
        }
    return C;
}
Mat tr(const Mat &A, int n, int m) {
    Mat T((size_t)n * m);
    for (int i = 0; i < n; ++i) for (int j = 0; j < m; ++j) T[(size_t)j * n + i] = A[(size_t)i * m + j];
    return T;
}
Mat dbl(const float *p, size_t n) { Mat v(n); for (size_t i = 0; i < n; ++i) v[i] = p[i]; return v; }
}  // namespace

// Folds the reference's Linear chains (see the header comment) in float64 and writes the tensor-core operand blob.
int tc_pack(const tm_encoder_desc &d, const tm_encoder_params &p, float *blob) {
    const TcLayout L = make_tc_layout(d);
    memset(blob, 0, sizeof(float) * ((L.total + 31) & ~(int64_t)31));
    const int H = L.H, H2 = 2 * L.H, D = L.D, Ed = L.Ed, M = L.M, ev = Ed + 3 + D;
    // ---- lin_event (explainer.py:93): K columns reordered to [edge | time]; the 3 edge-identity columns go to the CUDA cores
    {
        Mat W((size_t)D * (Ed + D));
        for (int n = 0; n < D; ++n) {
            for (int j = 0; j < Ed; ++j) W[(size_t)n * (Ed + D) + j] = p.lin_event_w[(size_t)n * ev + j];
            for (int t = 0; t < D; ++t) W[(size_t)n * (Ed + D) + Ed + t] = p.lin_event_w[(size_t)n * ev + Ed + 3 + t];
        }
        pack_tc_lin(L.evt, Ed + D, D, W.data(), blob);
        float *c = blob + L.cstE;
        for (int n = 0; n < D; ++n) {
            c[L.e_b + n] = p.lin_event_b[n];
            double b2 = p.lin_event_b[n];             // position-2 rows: dt = 0, TimeEncode = cos(phase) in the chunks they skip
            for (int t = 0; t < D; ++t)
                if (Ed + t >= L.nch_edge * kKC) b2 += (double)p.lin_event_w[(size_t)n * ev + Ed + 3 + t] * cos((double)p.phase[t]);
            c[L.e_b2 + n] = (float)b2;
            for (int k = 0; k < 3; ++k) c[L.e_wi + k * L.D16 + n] = p.lin_event_w[(size_t)n * ev + Ed + k];
            c[L.e_freq + n] = p.basis_freq[n]; c[L.e_phase + n] = p.phase[n];
        }
        for (int n = 0; n < H; ++n) c[L.e_g0b + n] = p.gcn0_b[n];
        const Mat G0 = dbl(p.gcn0_w, (size_t)H * D);
        pack_tc_lin(L.g0, D, H, G0.data(), blob);
    }
    // ---- folded motif-level matrices
    const Mat G2 = dbl(p.gcn2_w, (size_t)H * H);
    Mat G((size_t)H2 * H2, 0.0), g(H2);
    for (int i = 0; i < H; ++i) {
        for (int j = 0; j < H; ++j) { G[(size_t)i * H2 + j] = G2[(size_t)i * H + j]; G[(size_t)(H + i) * H2 + H + j] = G2[(size_t)i * H + j]; }
        g[i] = g[H + i] = p.gcn2_b[i];
    }
    const Mat W1 = dbl(p.att_w1_w, (size_t)H2 * H2), W2 = dbl(p.att_w2_w, (size_t)H2 * H2);
    const Mat A1 = mm(W1, G, H2, H2, H2), A2 = mm(W2, G, H2, H2, H2);
    Mat c1 = mm(W1, g, H2, H2, 1), c2 = mm(W2, g, H2, H2, 1);
    for (int i = 0; i < H2; ++i) { c1[i] += p.att_w1_b[i]; c2[i] += p.att_w2_b[i]; }
    const Mat A2t = tr(A2, H2, H2), A1t = tr(A1, H2, H2);
    const Mat S = mm(A2t, A1, H2, H2, H2), cu = mm(A2t, c1, H2, H2, 1), dv = mm(A1t, c2, H2, H2, 1);
    double e0 = 0;
    for (int i = 0; i < H2; ++i) e0 += c1[i] * c2[i];
    const Mat A0 = dbl(p.att_mlp0_w, (size_t)H * H2);
    const Mat P = mm(A0, G, H, H2, H2), Q = mm(A0, A2, H, H2, H2);
    Mat gc(H2);
    for (int i = 0; i < H2; ++i) gc[i] = g[i] + c2[i];
    Mat cy = mm(A0, gc, H, H2, 1);
    for (int i = 0; i < H; ++i) cy[i] += p.att_mlp0_b[i];
    Mat SP((size_t)3 * H * H2);
    memcpy(SP.data(), S.data(), sizeof(double) * S.size());
    memcpy(SP.data() + S.size(), P.data(), sizeof(double) * P.size());
    pack_tc_lin(L.sp, H2, 3 * H, SP.data(), blob);
    pack_tc_lin(L.q, H2, H, Q.data(), blob);
    const Mat A3 = dbl(p.att_mlp3_w, (size_t)H * H);
    Mat M0a((size_t)M * H);
    for (int m = 0; m < M; ++m) for (int j = 0; j < H; ++j) M0a[(size_t)m * H + j] = p.mlp0_w[(size_t)m * M + j];
    const Mat R = mm(M0a, A3, M, H, H);
    pack_tc_lin(L.r, H, M, R.data(), blob);
    const Mat a3b = dbl(p.att_mlp3_b, H);
    const Mat cm0 = mm(M0a, a3b, M, H, 1);
    for (int c = 0; c < 12; ++c)
        for (int m = 0; m < M; ++m)
            blob[L.cm + (int64_t)c * L.M16 + m] = (float)(cm0[m] + p.mlp0_b[m] + (L.if_cat ? (double)p.mlp0_w[(size_t)m * M + H + c] : 0.0));
    const Mat M3 = dbl(p.mlp3_w, (size_t)H * M);
    pack_tc_lin(L.m3, M, H, M3.data(), blob);
    float *c = blob + L.cstM;
    for (int i = 0; i < H2; ++i) { c[L.m_cu + i] = (float)cu[i]; c[L.m_d + i] = (float)dv[i]; }
    c[L.m_e] = (float)e0;
    for (int i = 0; i < H; ++i) { c[L.m_cy + i] = (float)cy[i]; c[L.m_m3b + i] = p.mlp3_b[i]; c[L.m_w5 + i] = p.mlp5_w[i]; }
    c[L.m_b5] = p.mlp5_b[0];
    return TM_OK;
}

// ---------------------------------------------------------------------------------------------
// One GEMM round of a CTA: D[128 x n16] (TMEM column d_col) (+)= A[128 x kcols] * B[n16 x kcols]^T, 3xTF32.
// Single A tile pair (hi, lo) and single weight buffer: the round's weight chunk was requested when the previous
// round's MMAs completed; the co-resident CTAs of the SM cover each other's waits.
// ---------------------------------------------------------------------------------------------
struct TcCtx {
    uint8_t *a;            // A operand: hi tile, then lo tile
    uint32_t a_s, b_s;     // shared-space addresses of the A tiles and of the weight buffer
    uint64_t *bars;        // [0] MMAs done, [1] weight chunk landed
    uint32_t mma_phase, b_phase;
    uint32_t tmem;
    const float *blob;
};

__device__ __forceinline__ void tc_request_b(const TcCtx &x, int64_t off, int bytes) {       // one thread
    tc::mbar_expect_tx(x.bars + 1, (uint32_t)bytes);
    tc::tma_load_1d_s(x.b_s, x.blob + off, (uint32_t)bytes, x.bars + 1);
}

__device__ __forceinline__ void store_a4(const TcCtx &x, int row, int k, float4 v) {
    float4 h, l;
    tc::split_tf32(v.x, h.x, l.x); tc::split_tf32(v.y, h.y, l.y); tc::split_tf32(v.z, h.z, l.z); tc::split_tf32(v.w, h.w, l.w);
    uint8_t *p = x.a + tc::tile_off(128, row, k);
    *reinterpret_cast<float4 *>(p) = h;
    *reinterpret_cast<float4 *>(p + kATile) = l;
}

// fill() has written this thread's share of the A tiles.  next_bytes != 0: weight chunk of the CTA's next round.
__device__ __forceinline__ void tc_mma_round(TcCtx &x, int n16, int kcols, int d_col, bool accumulate, int64_t next_off, int next_bytes) {
    tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) {          // warp 0 (warp-uniform): one elected lane issues
        tc::mbar_wait(x.bars + 1, x.b_phase);
        tc::fence_after_sync();
        const uint32_t leader = tc::elect_one();
        const uint32_t idesc = tc::idesc_tf32(128, n16);
        const uint32_t lbo_a = 128 * 16, lbo_b = (uint32_t)n16 * 16;
        uint64_t ah = tc::smem_desc(x.a_s, lbo_a, 128), al = tc::smem_desc(x.a_s + kATile, lbo_a, 128);
        uint64_t bh = tc::smem_desc(x.b_s, lbo_b, 128), bl = tc::smem_desc(x.b_s + (uint32_t)n16 * kKC * 4, lbo_b, 128);
        const uint64_t da = (2 * lbo_a) >> 4, db = (2 * lbo_b) >> 4;      // descriptor start-address step per K = 8
        const uint32_t dcol = x.tmem + (uint32_t)d_col;
        for (int ks = 0; ks < kcols / 8; ++ks) {
            tc::mma_tf32(dcol, ah, bh, idesc, (uint32_t)(accumulate || ks != 0), leader);
            tc::mma_tf32(dcol, al, bh, idesc, 1, leader);
            tc::mma_tf32(dcol, ah, bl, idesc, 1, leader);
            ah += da; al += da; bh += db; bl += db;
        }
        tc::mma_commit(x.bars, leader);
        __syncwarp();
    }
    x.b_phase ^= 1;
    tc::mbar_wait(x.bars, x.mma_phase);
    x.mma_phase ^= 1;
    tc::fence_after_sync();
    if (threadIdx.x == 0 && next_bytes) tc_request_b(x, next_off, next_bytes);       // the weight buffer is free again
}

struct TcArgs {
    int64_t n_motifs, W, group, m_begin;     // this launch scores motifs [m_begin, m_begin + slab)
    int64_t slab;
    const int32_t *nodes, *eidx;
    const float *t;
    const uint8_t *cat;
    const float *cut, *eid, *node_feat, *edge_feat, *std_;
    int64_t n_node_rows, n_edge_rows;
    float *F;                                // scratch: per CTA 12 h slabs [position][column chunk][piece k/4][128 rows][4]
    float *scores;
    uint32_t tmem_cols;
    int b_bytes;                             // bytes of the weight-chunk buffer
};

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }

// cos(x) for the TimeEncode arguments (they reach 1e8 and beyond, where the library cosf takes its slow path).
// Exact argument reduction in integer arithmetic: |x| = m * 2^e with a 24-bit integer m, so frac(|x| / 2pi) =
// frac(m * frac(2^e / 2pi)); kInv2Pi[e + 44] holds frac(2^e / 2pi) in 0.64 fixed point, of which the top 32 bits of the
// product are kept (error < 2^-32 turn = 1.5e-9 rad).  The turn fraction is split into a quadrant and an angle in
// [-pi/4, pi/4) for the fdlibm single-precision sin/cos kernels.  Max error ~1.5 ulp of 1.0 against the exact cosine
// of the fp32 argument over |x| <= 1e11 (cosf: 1-2 ulp); |x| >= 2^43, inf and nan go to cosf.  The table is read from
// shared memory (a copy of kInv2Pi): lanes index it with different exponents.
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

__device__ __forceinline__ float cos_accurate(float x, const uint2 *tab) {
    const uint32_t bits = __float_as_uint(x) & 0x7fffffffu;
    const int e = (int)(bits >> 23) - 150;                 // |x| = m * 2^e
    if (e > 19) return cosf(x);
    const uint32_t m = e < -44 ? 0u : ((bits & 0x7fffffu) | 0x800000u);        // tiny |x|: angle 0
    const uint2 T = tab[max(e, -44) + 44];                 // {low, high} words of frac(2^e / 2pi)
    const uint32_t fr = m * T.y + __umulhi(m, T.x) + (1u << 29);               // turn fraction + 1/8 turn, 0.32 fixed point
    const int q = (int)(fr >> 30);
    const int r = (int)(fr & 0x3fffffffu) - (1 << 29);                          // angle inside the quadrant, [-1/8, 1/8) turn
    const float th = (float)r * 1.46291807926715968e-9f /* 2 pi / 2^32 */, z = th * th;
    const float cs = fmaf(z, fmaf(z, fmaf(z, fmaf(z, 2.43904487962774090654e-5f, -1.38867637746099294692e-3f), 4.16666233237390631894e-2f), -4.99999997251031003120e-1f), 1.f);
    const float sn = fmaf(th * z, fmaf(z, fmaf(z, fmaf(z, 2.7183114939898219064e-6f, -1.98393348360966317347e-4f), 8.3333293858894631756e-3f), -1.66666666416265235595e-1f), th);
    const float v = (q & 1) ? sn : cs;                     // cos(q pi/2 + th) = {cs, -sn, -cs, sn}[q]
    return ((q + 1) & 2) ? -v : v;
}

// ---------------------------------------------------------------------------------------------
// score_tc_kernel: one persistent launch scores all motifs.  A CTA (256 threads = 128 motifs x 2 column halves, two
// CTAs per SM) takes a tile of 128 motifs through
//   event passes, one per walk position p (row = motif): lin_event -> E, MLP.0 for both orientations -> Zs, Zt,
//     h_p = relu(. + bias) written to the CTA's private 192 KB scratch (12 [128 x 32] slabs, L2 resident);
//   motif rounds: [S; P] h_2 -> scores -> temporal weights, softmax -> + Q mix -> R -> MLP.3 -> MLP.5 + sigmoid.
// While one CTA of the SM is in its (CUDA-core heavy) event passes the other is usually in its (tensor heavy)
// motif rounds.
// TMEM (256 columns): event passes Zs [0,H)  Zt [H,2H)  E [2H, 2H + D16), or E aliasing Zt when MLP.0 has a single
// K chunk (D <= 32): E has then been read completely before the MMA that writes Zt is issued.  Motif rounds:
// U [0,2H) | Y [2H,3H) (one N = 3H accumulator of the [S; P] rounds), later M0 [0,M16) and M1 [2H,3H).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldcg4(const float *p) { return __ldcg(reinterpret_cast<const float4 *>(p)); }

__global__ void __launch_bounds__(kTcThreads, 2)
score_tc_kernel(const TcLayout L, const float *__restrict__ blob, const TcArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float part[2][3][128];
    const int t = threadIdx.x, warp = t >> 5, row = t & 127, half = t >> 7, kb = 16 * half;
    TcCtx x;
    x.a = smem; x.a_s = tc::smem_u32(smem); x.b_s = x.a_s + 2 * kATile; x.bars = bars; x.mma_phase = 0; x.b_phase = 0; x.blob = blob;
    float *cstE = reinterpret_cast<float *>(smem + 2 * kATile + a.b_bytes), *cstM = cstE + L.n_cstE;
    uint2 *ctab = reinterpret_cast<uint2 *>(cstM + L.n_cstM);
    for (int i = t; i < L.n_cstE + L.n_cstM; i += kTcThreads) cstE[i] = __ldg(blob + L.cstE + i);      // cstM follows cstE in the blob
    if (t < 64) ctab[t] = make_uint2((uint32_t)kInv2Pi[t], (uint32_t)(kInv2Pi[t] >> 32));
    if (t == 0) { tc::mbar_init(bars, 1); tc::mbar_init(bars + 1, 1); }
    if (warp == 0) tc::tmem_alloc(&tmem_slot, a.tmem_cols);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot, lane_base = (uint32_t)((warp & 3) * 32) << 16;
    x.tmem = tmem;
    const int H = L.H, H2 = 2 * L.H, D = L.D, Ed = L.Ed, nG = L.g0.nch, nchS = L.sp.nch;
    const int colZ = 0, colE = (nG == 1 && L.D16 <= H) ? H : H2;
    const int colU = 0, colY = H2, colM0 = 0, colM1 = H2;
    const int64_t n_m = a.n_motifs, n_tiles = (n_m + 127) / 128;
    const int bytes_e = (int)chunk_floats(L.evt) * 4, bytes_g = (int)chunk_floats(L.g0) * 4;
    const int bytes_sp = (int)chunk_floats(L.sp) * 4, bytes_q = (int)chunk_floats(L.q) * 4, bytes_r = (int)chunk_floats(L.r) * 4, bytes_m3 = (int)chunk_floats(L.m3) * 4;
    if (t == 0 && blockIdx.x < n_tiles) tc_request_b(x, L.evt.w, bytes_e);
    const bool ed_vec = (Ed & 3) == 0, d_vec = (D & 3) == 0;
    float *Fs = a.F + (int64_t)blockIdx.x * 12 * kSlabFloats;           // this CTA's h slabs: [position][column chunk][piece k/4][128 rows][4]
    const float *F0 = Fs, *F1 = Fs + 4 * kSlabFloats, *F2 = Fs + 8 * kSlabFloats;
    // this thread's 16 columns [kb, kb+16) of row `row` of a [128 x 32] slab (four coalesced 16-byte pieces)
    auto ld16 = [&](const float *slab, float *v) {
#pragma unroll
        for (int g = 0; g < 4; ++g) { const float4 f = ldcg4(slab + ((kb >> 2) + g) * 512 + row * 4); v[4 * g] = f.x; v[4 * g + 1] = f.y; v[4 * g + 2] = f.z; v[4 * g + 3] = f.w; }
    };

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t gm_ = tile * 128 + row;
        const bool live = gm_ < n_m, more = tile + gridDim.x < n_tiles;
        const int64_t gm = live ? gm_ : 0;
        // =========================== event passes ===========================
#pragma unroll 1
        for (int pos = 0; pos < 3; ++pos) {
            const int nE = pos == 2 ? L.nch_edge : L.evt.nch;          // position 2: dt = 0, the pure TimeEncode chunks are in the bias
            int64_t e = 0, ns = 0, nt = 0; float dt = 0.f, ei0 = 0.f, ei1 = 0.f, ei2 = 0.f;
            if (live) {
                e = a.eidx[gm * 3 + pos]; ns = a.nodes[gm * 6 + 2 * pos]; nt = a.nodes[gm * 6 + 2 * pos + 1];
                dt = __fsub_rn(a.t[gm * 3 + 2], a.t[gm * 3 + pos]);                       // explainer.py:326
                if (a.eid) { const float *ei = a.eid + gm * 9 + pos * 3; ei0 = __ldg(ei); ei1 = __ldg(ei + 1); ei2 = __ldg(ei + 2); }
            }
            const bool e_ok = live && e >= 0 && e < a.n_edge_rows, s_ok = live && ns >= 0 && ns < a.n_node_rows, t_ok = live && nt >= 0 && nt < a.n_node_rows;
            const float *ef = a.edge_feat + e * Ed, *sf = a.node_feat + ns * D, *tf = a.node_feat + nt * D;
            auto xval = [&](int j) -> float {                                              // [edge features | TimeEncode] column j (:179, :55-58)
                if (j < Ed) return e_ok ? __ldg(ef + j) : 0.f;
                const int k = j - Ed;
                if (k < D) return live ? cos_accurate(__fadd_rn(__fmul_rn(dt, cstE[L.e_freq + k]), cstE[L.e_phase + k]), ctab) : 0.f;
                return 0.f;
            };
            // ---- lin_event (:93) -> E
            for (int c = 0; c < nE; ++c) {
                const int kcols = min(kKC, L.evt.K8 - c * kKC);
                float4 v[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int k = kb + 4 * g, j = c * kKC + k;
                    if (k >= kcols) { v[g] = make_float4(0.f, 0.f, 0.f, 0.f); continue; }
                    if (ed_vec && j + 3 < Ed) v[g] = e_ok ? ldg4(ef + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    else v[g] = make_float4(xval(j), xval(j + 1), xval(j + 2), xval(j + 3));
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) if (kb + 4 * g < kcols) store_a4(x, row, kb + 4 * g, v[g]);
                const bool last = c + 1 == nE;
                tc_mma_round(x, L.D16, kcols, colE, c != 0, last ? L.g0.w : L.evt.w + (int64_t)(c + 1) * chunk_floats(L.evt), last ? bytes_g : bytes_e);
            }
            // ---- event_conv.MLP.0 on src + relu(tgt + event) (o = 0) and tgt + relu(src + event) (o = 1) (:94-95, :182-184) -> Zs, Zt
            const int eb = pos == 2 ? L.e_b2 : L.e_b;
#pragma unroll 1
            for (int o = 0; o < 2; ++o)
                for (int c = 0; c < nG; ++c) {
                    const int kcols = min(kKC, L.g0.K8 - c * kKC);
                    if (kb < kcols) {
                        float sv[16], gv[16];
#pragma unroll
                        for (int k = 0; k < 16; k += 4) {       // the gathers first (explainer.py:348-351)
                            const int j = c * kKC + kb + k;
                            if (d_vec && j + 3 < D) {
                                const float4 s4 = s_ok ? ldg4(sf + j) : make_float4(0.f, 0.f, 0.f, 0.f), g4 = t_ok ? ldg4(tf + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                                sv[k] = s4.x; sv[k + 1] = s4.y; sv[k + 2] = s4.z; sv[k + 3] = s4.w; gv[k] = g4.x; gv[k + 1] = g4.y; gv[k + 2] = g4.z; gv[k + 3] = g4.w;
                            } else {
#pragma unroll
                                for (int i = 0; i < 4; ++i) { sv[k + i] = (j + i < D && s_ok) ? __ldg(sf + j + i) : 0.f; gv[k + i] = (j + i < D && t_ok) ? __ldg(tf + j + i) : 0.f; }
                            }
                        }
                        float evv[16];
                        tc::tmem_ld16(tmem + lane_base + colE + c * kKC + kb, evv);
#pragma unroll
                        for (int k = 0; k < 16; k += 4) {
                            float z[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int j = c * kKC + kb + k + i;          // j < D16: constants are zero-padded
                                float e_ = evv[k + i] + cstE[eb + j];
                                e_ = fmaf(cstE[L.e_wi + j], ei0, e_); e_ = fmaf(cstE[L.e_wi + L.D16 + j], ei1, e_); e_ = fmaf(cstE[L.e_wi + 2 * L.D16 + j], ei2, e_);
                                const float p_ = o ? gv[k + i] : sv[k + i], q_ = o ? sv[k + i] : gv[k + i];
                                z[i] = (j < D && live) ? p_ + fmaxf(q_ + e_, 0.f) : 0.f;
                            }
                            store_a4(x, row, kb + k, make_float4(z[0], z[1], z[2], z[3]));
                        }
                    }
                    const bool last = c + 1 == nG;
                    int64_t noff; int nbytes;
                    if (!last) { noff = L.g0.w + (int64_t)(c + 1) * chunk_floats(L.g0); nbytes = bytes_g; }
                    else if (o == 0) { noff = L.g0.w; nbytes = bytes_g; }
                    else if (pos < 2) { noff = L.evt.w; nbytes = bytes_e; }
                    else { noff = L.sp.w; nbytes = bytes_sp; }
                    tc_mma_round(x, H, kcols, colZ + o * H, c != 0, noff, nbytes);
                }
            // ---- h_pos = relu(MLP.0 + bias): half o owns orientation o = column chunks 2o, 2o+1
            {
                float *fo = Fs + (pos * 4 + 2 * half) * kSlabFloats + row * 4;
                for (int c0 = 0; c0 < H; c0 += 16) {
                    float v[16];
                    tc::tmem_ld16(tmem + lane_base + colZ + half * H + c0, v);
                    float *fc = fo + (c0 >> 5) * kSlabFloats + ((c0 & 31) >> 2) * 512;
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        const float4 bb = lds4(cstE + L.e_g0b + c0 + i);
                        __stcg(reinterpret_cast<float4 *>(fc + (i >> 2) * 512),
                               make_float4(fmaxf(v[i] + bb.x, 0.f), fmaxf(v[i + 1] + bb.y, 0.f), fmaxf(v[i + 2] + bb.z, 0.f), fmaxf(v[i + 3] + bb.w, 0.f)));
                    }
                }
            }
            tc::fence_before_sync();
            __syncthreads();            // TMEM reads done before the next pass overwrites E / Z; the h slabs are visible to the CTA
            tc::fence_after_sync();
        }
        // =========================== motif rounds ===========================
        // ---- [U | Y] = [S; P] h_2 ; r = d . h_2
        float rp = 0.f;
        {
            float nxt[16];
            ld16(F2, nxt);
            for (int c = 0; c < nchS; ++c) {
                float cur[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) cur[i] = nxt[i];
                if (c + 1 < nchS) ld16(F2 + (c + 1) * kSlabFloats, nxt);
#pragma unroll
                for (int k = 0; k < 16; k += 4) {
                    const float4 dd = lds4(cstM + L.m_d + c * kKC + kb + k);
                    rp = fmaf(dd.x, cur[k], rp); rp = fmaf(dd.y, cur[k + 1], rp); rp = fmaf(dd.z, cur[k + 2], rp); rp = fmaf(dd.w, cur[k + 3], rp);
                    store_a4(x, row, kb + k, make_float4(cur[k], cur[k + 1], cur[k + 2], cur[k + 3]));
                }
                const bool last = c + 1 == nchS;
                tc_mma_round(x, 3 * H, kKC, colU, c != 0, last ? L.q.w : L.sp.w + (int64_t)(c + 1) * chunk_floats(L.sp), last ? bytes_q : bytes_sp);
            }
        }
        // ---- s_k = h_k . (U + cu) + r  (:806-808 after folding)
        float s0 = 0.f, s1 = 0.f;
#pragma unroll 1
        for (int c = 0; c < nchS; c += 2) {
            float p0[2][16], p1[2][16];
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) { ld16(F0 + (c + cc) * kSlabFloats, p0[cc]); ld16(F1 + (c + cc) * kSlabFloats, p1[cc]); }
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                float u[16];
                tc::tmem_ld16(tmem + lane_base + colU + (c + cc) * kKC + kb, u);
#pragma unroll
                for (int k = 0; k < 16; k += 4) {
                    const float4 cu = lds4(cstM + L.m_cu + (c + cc) * kKC + kb + k);
                    const float u0 = u[k] + cu.x, u1 = u[k + 1] + cu.y, u2 = u[k + 2] + cu.z, u3 = u[k + 3] + cu.w;
                    s0 = fmaf(p0[cc][k], u0, s0); s0 = fmaf(p0[cc][k + 1], u1, s0); s0 = fmaf(p0[cc][k + 2], u2, s0); s0 = fmaf(p0[cc][k + 3], u3, s0);
                    s1 = fmaf(p1[cc][k], u0, s1); s1 = fmaf(p1[cc][k + 1], u1, s1); s1 = fmaf(p1[cc][k + 2], u2, s1); s1 = fmaf(p1[cc][k + 3], u3, s1);
                }
            }
        }
        float n0[16], n1[16];                       // mix operands of the first Q round, in flight across the reduction
        ld16(F0, n0); ld16(F1, n1);
        part[half][0][row] = s0; part[half][1][row] = s1; part[half][2][row] = rp;
        __syncthreads();
        {
            const float r_ = part[0][2][row] + part[1][2][row] + cstM[L.m_e];
            s0 = part[0][0][row] + part[1][0][row] + r_;
            s1 = part[0][1][row] + part[1][1][row] + r_;
        }
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
        // ---- Y += Q (alpha_0 h_0 + alpha_1 h_1)   (:841-843 after folding)
        for (int c = 0; c < nchS; ++c) {
            float c0[16], c1[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { c0[i] = n0[i]; c1[i] = n1[i]; }
            if (c + 1 < nchS) { ld16(F0 + (c + 1) * kSlabFloats, n0); ld16(F1 + (c + 1) * kSlabFloats, n1); }
#pragma unroll
            for (int k = 0; k < 16; k += 4)
                store_a4(x, row, kb + k, make_float4(fmaf(al0, c0[k], al1 * c1[k]), fmaf(al0, c0[k + 1], al1 * c1[k + 1]),
                                                     fmaf(al0, c0[k + 2], al1 * c1[k + 2]), fmaf(al0, c0[k + 3], al1 * c1[k + 3])));
            const bool last = c + 1 == nchS;
            tc_mma_round(x, H, kKC, colY, true, last ? L.r.w : L.q.w + (int64_t)(c + 1) * chunk_floats(L.q), last ? bytes_r : bytes_q);
        }
        // ---- M0 = R relu(Y + cy)   (attention.MLP.3 and MLP.0 folded)
        for (int c = 0; c < L.r.nch; ++c) {
            float z[16];
            tc::tmem_ld16(tmem + lane_base + colY + c * kKC + kb, z);
#pragma unroll
            for (int k = 0; k < 16; k += 4) {
                const float4 bb = lds4(cstM + L.m_cy + c * kKC + kb + k);
                store_a4(x, row, kb + k, make_float4(fmaxf(z[k] + bb.x, 0.f), fmaxf(z[k + 1] + bb.y, 0.f), fmaxf(z[k + 2] + bb.z, 0.f), fmaxf(z[k + 3] + bb.w, 0.f)));
            }
            const bool last = c + 1 == L.r.nch;
            tc_mma_round(x, L.M16, kKC, colM0, c != 0, last ? L.m3.w : L.r.w + (int64_t)(c + 1) * chunk_floats(L.r), last ? bytes_m3 : bytes_r);
        }
        // ---- M1 = MLP.3 relu(M0 + cm[category])   (:199)
        const float *cmr = blob + L.cm + (int64_t)((L.if_cat && live && a.cat) ? min((int)a.cat[gm], 11) : 0) * L.M16;
        for (int c = 0; c < L.m3.nch; ++c) {
            const int kcols = min(kKC, L.m3.K8 - c * kKC);
            if (kb < kcols) {
                float z[16];
                tc::tmem_ld16(tmem + lane_base + colM0 + c * kKC + kb, z);      // columns < M16 (kcols is a multiple of 8, M16 of 16)
#pragma unroll
                for (int k = 0; k < 16; k += 4) {
                    const float4 bb = ldg4(cmr + c * kKC + kb + k);
                    const int j = c * kKC + kb + k;
                    store_a4(x, row, kb + k, make_float4(j < L.M ? fmaxf(z[k] + bb.x, 0.f) : 0.f, j + 1 < L.M ? fmaxf(z[k + 1] + bb.y, 0.f) : 0.f,
                                                         j + 2 < L.M ? fmaxf(z[k + 2] + bb.z, 0.f) : 0.f, j + 3 < L.M ? fmaxf(z[k + 3] + bb.w, 0.f) : 0.f));
                }
            }
            const bool last = c + 1 == L.m3.nch;
            tc_mma_round(x, H, kcols, colM1, c != 0, last ? L.evt.w : L.m3.w + (int64_t)(c + 1) * chunk_floats(L.m3), last ? (more ? bytes_e : 0) : bytes_m3);
        }
        // ---- MLP.5 + sigmoid (:199-200)
        float z5 = 0.f;
        for (int c0 = half * (H / 2); c0 < half * (H / 2) + H / 2; c0 += 16) {
            float z[16];
            tc::tmem_ld16(tmem + lane_base + colM1 + c0, z);
#pragma unroll
            for (int i = 0; i < 16; ++i) z5 = fmaf(fmaxf(z[i] + cstM[L.m_m3b + c0 + i], 0.f), cstM[L.m_w5 + c0 + i], z5);
        }
        part[half][0][row] = z5;         // the score reduction's reads of part[] ended before the Q rounds' barriers
        tc::fence_before_sync();
        __syncthreads();                 // also: all TMEM reads of this tile done before the next tile's MMAs overwrite it
        tc::fence_after_sync();
        if (live && half == 0) a.scores[gm] = 1.f / (1.f + expf(-(part[0][0][row] + part[1][0][row] + cstM[L.m_b5])));
        // the next write to part[] comes after the barriers of the next tile's rounds
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, a.tmem_cols);
}

// Motifs whose h rows the workspace holds: two resident CTAs per SM x 128 motifs (tm_encoder_workspace_floats)
int64_t tc_slab_motifs() {
    static int64_t v = 0;
    if (!v) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        v = (int64_t)sms * 2 * 128;
    }
    return v;
}

// optional kernel timing (bench.py roofline): CUDA events around the launch
static bool g_prof = false;
static std::vector<cudaEvent_t> g_prof_ev;      // pairs: before, after
static size_t g_prof_used = 0;

// std_ = per-batch std (already computed); F = scratch for the h slabs of the resident CTAs
int tc_encode_score(const tm_encoder_desc &d, const float *d_blob_tc, int64_t B, int64_t W, int64_t group, const int32_t *nodes,
                    const int32_t *eidx, const float *t, const uint8_t *cat, const float *cut, const float *eid, const float *node_feat,
                    int64_t n_node_rows, const float *edge_feat, int64_t n_edge_rows, const float *std_, float *F, float *scores,
                    int device, cudaStream_t st) {
    const TcLayout L = make_tc_layout(d);
    if (L.H != 64) { set_error("tc_encode_score: hid_dim must be 64"); return TM_ERR_UNSUPPORTED; }
    if (device < 0 || device >= 64) { set_error("tc_encode_score: device index out of range"); return TM_ERR_UNSUPPORTED; }
    const int H = L.H;
    const bool alias_e = L.g0.nch == 1 && L.D16 <= H;
    uint32_t cols = 32;
    while ((int)cols < std::max(3 * H, alias_e ? 2 * H : 2 * H + L.D16)) cols <<= 1;
    if (cols > 512) { set_error("tc_encode_score: node_dim too large for the TMEM layout"); return TM_ERR_UNSUPPORTED; }
    int64_t bb = 0;
    for (const TcLin *l : {&L.evt, &L.g0, &L.sp, &L.q, &L.r, &L.m3}) bb = std::max(bb, chunk_floats(*l) * 4);
    const size_t need = (size_t)2 * kATile + (size_t)bb + (size_t)(L.n_cstE + L.n_cstM) * 4 + 64 * 8;
    if (need > 220 * 1024) { set_error("tc_encode_score: feature dims too large for one weight chunk in shared memory"); return TM_ERR_UNSUPPORTED; }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    // resident CTAs per SM: TMEM columns and shared memory (registers: __launch_bounds__(256, 2))
    const int ctas = std::max(1, std::min<int>(2, std::min<int>(512 / cols, (int)((228 * 1024) / (need + 1024 + 4096)))));
    // dynamic shared memory padded so that no more than `ctas` CTAs fit an SM (TMEM columns are not part of the occupancy
    // calculation: a CTA beyond 512 / cols would spin in tcgen05.alloc while holding its other resources)
    const size_t smem = std::max(need, std::min((size_t)228 * 1024 / (ctas + 1), (size_t)227 * 1024 - 4096));
    static bool attr_set[64] = {false};
    if (!attr_set[device]) {
        TM_CUDA(cudaFuncSetAttribute(score_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 4096));
        TM_CUDA(cudaFuncSetAttribute(score_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr_set[device] = true;
    }
    TcArgs a;
    a.n_motifs = B * W; a.W = W; a.group = group; a.m_begin = 0; a.slab = 0; a.nodes = nodes; a.eidx = eidx; a.t = t; a.cat = cat; a.cut = cut;
    a.eid = eid; a.node_feat = node_feat; a.edge_feat = edge_feat; a.std_ = std_; a.n_node_rows = n_node_rows; a.n_edge_rows = n_edge_rows;
    a.F = F; a.scores = scores; a.tmem_cols = cols; a.b_bytes = (int)bb;
    const int64_t tiles = (a.n_motifs + 127) / 128, cap = (int64_t)sms * ctas;
    const int64_t per_cta = (tiles + cap - 1) / cap;
    const unsigned grid = (unsigned)((tiles + per_cta - 1) / per_cta);          // every CTA gets the same number of tiles (+-1); grid <= 2 * sms
    if (getenv("TEMPME_TC_DEBUG")) fprintf(stderr, "[tc] score kernel: %u CTAs (%d per SM), smem %zu B (needs %zu), %u TMEM columns, %lld tiles\n", grid, ctas, smem, need, cols, (long long)tiles);
    cudaEvent_t *pe = nullptr;
    if (g_prof && g_prof_used + 2 <= g_prof_ev.size()) {        // pool is created by tm_encoder_profile(1); when exhausted, stop recording
        pe = &g_prof_ev[g_prof_used]; g_prof_used += 2;
        cudaEventRecord(pe[0], st);
    }
    score_tc_kernel<<<grid, kTcThreads, smem, st>>>(L, d_blob_tc, a);
    TM_LAUNCH_CHECK();
    if (pe) cudaEventRecord(pe[1], st);
    return TM_OK;
}

}  // namespace tmb

extern "C" int tm_encoder_profile(int enable) {
    tmb::g_prof = enable != 0;
    tmb::g_prof_used = 0;
    if (enable && tmb::g_prof_ev.empty()) {          // event pool, created outside any timed region
        tmb::g_prof_ev.resize(2 * 4096);
        for (auto &e : tmb::g_prof_ev) TM_CUDA(cudaEventCreate(&e));
    }
    return TM_OK;
}

// h_event_ms: accumulated milliseconds of score_tc_kernel since the last read; h_motif_ms: 0 (the event-level and the
// motif-level phases are one kernel)
extern "C" int tm_encoder_profile_read(float *h_event_ms, float *h_motif_ms) {
    if (!h_event_ms || !h_motif_ms) { tmb::set_error("tm_encoder_profile_read: null output"); return TM_ERR_ARG; }
    double e = 0;
    for (size_t i = 0; i + 1 < tmb::g_prof_used && i + 1 < tmb::g_prof_ev.size(); i += 2) {
        float a = 0;
        TM_CUDA(cudaEventSynchronize(tmb::g_prof_ev[i + 1]));
        TM_CUDA(cudaEventElapsedTime(&a, tmb::g_prof_ev[i], tmb::g_prof_ev[i + 1]));
        e += a;
    }
    *h_event_ms = (float)e; *h_motif_ms = 0.f;
    tmb::g_prof_used = 0;
    return TM_OK;
}
