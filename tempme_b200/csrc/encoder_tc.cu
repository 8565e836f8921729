// tcgen05 (5th-gen tensor core) scorer: TempME.forward (reference models/explainer.py:174-201) as one persistent kernel of
// 3xTF32 GEMM rounds with TMEM accumulators (and the dependency gate of retrieve_edge_imp_node at the end of the file).
//
// Algebra.  Between event_conv.MLP.0's ReLU and attention.MLP.0's ReLU the reference applies only linear maps and the
// two softmax weights, and between attention.MLP.0's ReLU and MLP.0's ReLU only linear maps and a one-hot, so those
// chains are folded on the host (float64, tc_pack) into single matrices.  With h_k = [relu(MLP.0(src side)) |
// relu(MLP.0(tgt side))] of event k (k = 2 is the event next to the root), G = blockdiag(MLP.2, MLP.2), g its bias:
//     Wp   = W1 (G h_2 + g) + b1 = A1 h_2 + c1            Wq_k = W2 (G h_k + g) + b2 = A2 h_k + c2      (:806-807)
//     s_k  = Wp . Wq_k = h_k . (S h_2 + cu) + (d . h_2 + e)      S = A2^T A1, cu = A2^T c1, d = A1^T c2, e = c1 . c2
//     y    = relu(attention.MLP.0(f_2 + sum_k alpha_k Wq_k)) = relu(P h_2 + Q (alpha_0 h_0 + alpha_1 h_1) + cy)   (:841-843)
//     m0   = relu(MLP.0([attention.MLP.3(y) | onehot(cat)])) = relu(R y + cm[cat])                                (:196-200)
// which halves the multiply-adds per motif and, with the chunk pairs below, takes the per-tile GEMM rounds from 34 to 16
// (D = Ed = 32).  The folded weights are rounded to fp32 once; scores agree with the unfolded fp32 evaluation to ~2e-7 relative.
//
// One persistent kernel (score_tc_kernel) takes tiles of 128 motifs through three event passes (lin_event over
// [edge features | TimeEncode]; the three edge-identity columns are added on the CUDA cores; position-2 rows have
// dt = 0, so their TimeEncode chunks collapse into a bias; event_conv.MLP.0 + ReLU for the two orientations) and the
// motif rounds ([S; P] h_2 -> scores -> temporal weights, softmax -> + Q mix -> R -> MLP.3 -> MLP.5 + sigmoid).
// Thread (row, part) owns TMEM lane `row` and CW of the 32 columns of every K chunk (CW = 16: 256 threads): it reads the previous
// accumulator row with tcgen05.ld, applies bias / ReLU / mixing in fp32 registers, splits into tf32 hi + lo and writes its columns of
// the A operand into TMEM (tcgen05.st; TS-mode MMA) -- or, with TEMPME_TC_A=smem, the K-major operand tile in shared memory.  Weights
// arrive pre-split and pre-tiled by 1-D bulk TMA; feature rows by tensor-map TMA (tile::gather4) one pass ahead.  Dual rounds use a
// second A buffer: two K chunks, or two row blocks sharing a chunk, per barrier.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "tc.cuh"
#include "timeenc.cuh"

namespace tmb {

constexpr int kKC = 32;           // K columns per operand chunk
constexpr int kSlabFloats = 128 * kKC;
constexpr uint32_t kATile = 128 * kKC * 4;

struct TcLin { int64_t w; int K8, N16, nch; };   // chunk c at w + c * 2 * N16 * kKC floats: [hi tile | lo tile]
struct TcLayout {
    int D, Ed, H, M, use_temporal, if_cat, D16, M16;
    int nch_edge;                                        // lin_event chunks that hold an edge-feature column
    TcLin evt, g0, sp, q, r, m3;
    TcLin evtT;                                          // lin_event restricted to its TimeEncode columns (edge-projection mode: the edge columns
                                                         //   are applied once per edge id by tc_project_edges, not once per walk event)
    int64_t we;                                          // lin_event's edge columns, plain fp32 [D][Ed] (operand of the projection kernel)
    int e_b, e_b2, e_b2p, e_wi, e_g0b, e_freq, e_phase, n_cstE; // event-kernel constants (offsets inside cstE)
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
    L.evtT = lin(L.D, L.D);
    L.we = o; o += ((int64_t)L.D * L.Ed + 31) & ~(int64_t)31;
    int c = 0;
    L.e_b = c; c += L.D16; L.e_b2 = c; c += L.D16; L.e_b2p = c; c += L.D16; L.e_wi = c; c += 3 * L.D16; L.e_g0b = c; c += r16(H);
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
        {   // edge-projection mode: TimeEncode columns only on the tensor cores, edge columns as a plain matrix for tc_project_edges
            Mat Wt((size_t)D * D);
            for (int n = 0; n < D; ++n) {
                for (int t = 0; t < D; ++t) Wt[(size_t)n * D + t] = p.lin_event_w[(size_t)n * ev + Ed + 3 + t];
                for (int j = 0; j < Ed; ++j) blob[L.we + (int64_t)n * Ed + j] = p.lin_event_w[(size_t)n * ev + j];
            }
            pack_tc_lin(L.evtT, D, D, Wt.data(), blob);
        }
        float *c = blob + L.cstE;
        for (int n = 0; n < D; ++n) {
            c[L.e_b + n] = p.lin_event_b[n];
            double b2p = p.lin_event_b[n];            // position-2 rows in edge-projection mode: every TimeEncode column is cos(phase)
            for (int t = 0; t < D; ++t) b2p += (double)p.lin_event_w[(size_t)n * ev + Ed + 3 + t] * cos((double)p.phase[t]);
            c[L.e_b2p + n] = (float)b2p;
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
    uint8_t *a;            // A operand in shared memory (SS mode): hi tile, then lo tile
    uint32_t a_s, b_s;     // shared-space addresses of the A tiles and of the weight buffer
    uint32_t a_col;        // A operand in TMEM (TS mode): hi at columns [a_col, a_col + 32), lo at [a_col + 32, a_col + 64)
    uint32_t a_col2;       // second A buffer of the dual rounds (event passes: the 64 columns below a_col; Q / R rounds: columns [0, 64))
    uint64_t *bars;        // [0] MMAs done, [1] weight chunk landed
    uint32_t mma_phase, b_phase;
    uint32_t tmem;
    const float *blob;
    long long *dbg;        // TEMPME_TC_TIMING: thread 0 of CTA 0 stamps 5 clocks per round (entry, after sync, weights landed, MMAs issued, MMAs done)
    int dbg_i;
};

__device__ __forceinline__ void tc_request_b(const TcCtx &x, int64_t off, int bytes) {       // one thread
    tc::mbar_expect_tx(x.bars + 1, (uint32_t)bytes);
    tc::tma_load_1d_s(x.b_s, x.blob + off, (uint32_t)bytes, x.bars + 1);
}

// A-fill of one round.  SS mode: 16-byte pieces of the hi / lo operand tiles in shared memory.  TS mode: the thread's CW columns are
// collected in registers and written to its TMEM lane with tcgen05.st (commit); kk = column offset inside the thread's CW columns.
template <int CW, bool TS>
struct AFill {
    float hi[TS ? CW : 4], lo[TS ? CW : 4];
    __device__ __forceinline__ void put4(const TcCtx &x, int row, int kb, int kk, float4 v) {
        float4 h, l;
        tc::split_tf32(v.x, h.x, l.x); tc::split_tf32(v.y, h.y, l.y); tc::split_tf32(v.z, h.z, l.z); tc::split_tf32(v.w, h.w, l.w);
        if (TS) {
            hi[kk] = h.x; hi[kk + 1] = h.y; hi[kk + 2] = h.z; hi[kk + 3] = h.w;
            lo[kk] = l.x; lo[kk + 1] = l.y; lo[kk + 2] = l.z; lo[kk + 3] = l.w;
        } else {
            uint8_t *p = x.a + tc::tile_off(128, row, kb + kk);
            *reinterpret_cast<float4 *>(p) = h;
            *reinterpret_cast<float4 *>(p + kATile) = l;
        }
    }
    // all CW columns have been put; second: the event passes' second A buffer (the 64 TMEM columns below the primary one)
    __device__ __forceinline__ void commit(const TcCtx &x, uint32_t lane_base, int kb, bool second = false) { commit_at(x, lane_base, kb, second ? x.a_col2 : x.a_col); }
    __device__ __forceinline__ void commit_at(const TcCtx &x, uint32_t lane_base, int kb, uint32_t col) {
        if (TS) {
            const uint32_t a0 = x.tmem + lane_base + col + (uint32_t)kb;
            if (CW == 16) { tc::tmem_st16(a0, hi); tc::tmem_st16(a0 + kKC, lo); }
            else { tc::tmem_st8(a0, hi); tc::tmem_st8(a0 + kKC, lo); }
            tc::tmem_st_wait();
        }
    }
};

// fill() has written this thread's share of the A tiles.  next_bytes != 0: weight chunk of the CTA's next round.
struct NoMid { __device__ __forceinline__ void operator()() const {} };
// mid(): run by every warp between the issue of the round's MMAs and the wait for their completion (work that overlaps the tensor pipe)
// Dual rounds (TS mode, two A buffers): kDualK: D (+)= [A | A2] * [chunk | next chunk]^T (kcols + kc2 K columns);
// kDualM: D (+)= A * chunk^T and D2 (+)= A2 * chunk^T (two row blocks share the chunk).
enum { kSingle = 0, kDualK = 1, kDualM = 2 };
struct Dual { int mode, kc2, d2, kc3, a3; };      // kc3 != 0 (with kDualK): a third K chunk from the A buffer at TMEM column a3
template <bool TS, typename Mid = NoMid>
__device__ __forceinline__ void tc_mma_round(TcCtx &x, int n16, int kcols, int d_col, bool accumulate, int64_t next_off, int next_bytes, Mid mid = Mid(),
                                             Dual dual = Dual{kSingle, 0, 0, 0, 0}) {
#ifdef TM_TC_TIMING
    const bool tim = x.dbg && threadIdx.x == 0 && x.dbg_i < 128;
#else
    constexpr bool tim = false;        // clock stamps compiled out (build with -DTM_TC_TIMING, TEMPME_BUILD_TIMING=1)
#endif
    if (tim) x.dbg[x.dbg_i * 5 + 0] = clock64();
    if (!TS) tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();
    if (tim) x.dbg[x.dbg_i * 5 + 1] = clock64();
    if (threadIdx.x < 32) {          // warp 0 (warp-uniform): one elected lane issues
        tc::mbar_wait(x.bars + 1, x.b_phase);
        if (tim) x.dbg[x.dbg_i * 5 + 2] = clock64();
        tc::fence_after_sync();
        const uint32_t leader = tc::elect_one();
        const uint32_t idesc = tc::idesc_tf32(128, n16);
        const uint32_t lbo_a = 128 * 16, lbo_b = (uint32_t)n16 * 16;
        uint64_t ah = tc::smem_desc(x.a_s, lbo_a, 128), al = tc::smem_desc(x.a_s + kATile, lbo_a, 128);
        uint64_t bh = tc::smem_desc(x.b_s, lbo_b, 128), bl = tc::smem_desc(x.b_s + (uint32_t)n16 * kKC * 4, lbo_b, 128);
        const uint64_t da = (2 * lbo_a) >> 4, db = (2 * lbo_b) >> 4;      // descriptor start-address step per K = 8
        if (TS) {
            // The operands of every MMA are derived from warp-uniform bases taken through a warp reduction (REDUX writes a uniform register):
            // ptxas then steps addresses and descriptors on the uniform datapath instead of moving each operand of each MMA from the
            // elected lane's vector registers (tools/mma_rate.py: with lean issue code an MMA costs 128 N / 256 cycles, i.e. 16 at N = 32).
            const uint32_t tm_u = __reduce_or_sync(0xffffffffu, x.tmem), bs_u = __reduce_or_sync(0xffffffffu, x.b_s);
            const uint32_t dcol = tm_u + (uint32_t)d_col, ta = tm_u + x.a_col, ta2 = tm_u + x.a_col2;
            // low words of the weight descriptors: [hi tile | lo tile] per chunk, chunks back to back; a K = 8 step advances the start address
            const uint32_t hiw = tc::smem_desc_hi(128), tile = ((uint32_t)n16 * kKC * 4) >> 4, dbw = (2 * lbo_b) >> 4;
            const uint32_t b0 = tc::smem_desc_lo(bs_u, lbo_b);
            const int nks = kcols >> 3;                     // K chunks have at most kKC / 8 = 4 steps: unrolled, so the offsets are immediates
#pragma unroll
            for (int ks = 0; ks < kKC / 8; ++ks) {
                if (ks < nks) {
                    tc::mma3_tf32_ts(dcol, ta + 8 * ks, ta + kKC + 8 * ks, b0 + ks * dbw, b0 + tile + ks * dbw, hiw, idesc, (uint32_t)(accumulate || ks != 0), leader);
                    if (dual.mode == kDualM)
                        tc::mma3_tf32_ts(tm_u + (uint32_t)dual.d2, ta2 + 8 * ks, ta2 + kKC + 8 * ks, b0 + ks * dbw, b0 + tile + ks * dbw, hiw, idesc,
                                         (uint32_t)(accumulate || ks != 0), leader);
                }
            }
            if (dual.mode == kDualK) {                      // second K chunk: A2 with the chunk that follows in the weight buffer
                const uint32_t b2 = b0 + 2 * tile;
                const int nk2 = dual.kc2 >> 3;
#pragma unroll
                for (int ks = 0; ks < kKC / 8; ++ks)
                    if (ks < nk2) tc::mma3_tf32_ts(dcol, ta2 + 8 * ks, ta2 + kKC + 8 * ks, b2 + ks * dbw, b2 + tile + ks * dbw, hiw, idesc, 1, leader);
                if (dual.kc3) {                             // third K chunk
                    const uint32_t t3 = tm_u + (uint32_t)dual.a3, b3 = b0 + 4 * tile;
                    const int nk3 = dual.kc3 >> 3;
#pragma unroll
                    for (int ks = 0; ks < kKC / 8; ++ks)
                        if (ks < nk3) tc::mma3_tf32_ts(dcol, t3 + 8 * ks, t3 + kKC + 8 * ks, b3 + ks * dbw, b3 + tile + ks * dbw, hiw, idesc, 1, leader);
                }
            }
        } else {
            const uint32_t dcol = x.tmem + (uint32_t)d_col;
            for (int ks = 0; ks < kcols / 8; ++ks) {
                tc::mma_tf32(dcol, ah, bh, idesc, (uint32_t)(accumulate || ks != 0), leader);
                tc::mma_tf32(dcol, al, bh, idesc, 1, leader);
                tc::mma_tf32(dcol, ah, bl, idesc, 1, leader);
                ah += da; al += da; bh += db; bl += db;
            }
        }
        tc::mma_commit(x.bars, leader);
        if (tim) x.dbg[x.dbg_i * 5 + 3] = clock64();
        __syncwarp();
    }
    x.b_phase ^= 1;
    mid();
    tc::mbar_wait(x.bars, x.mma_phase);
    if (tim) x.dbg[x.dbg_i * 5 + 4] = clock64();
#ifdef TM_TC_TIMING
    if (x.dbg) x.dbg_i++;
#endif
    x.mma_phase ^= 1;
    tc::fence_after_sync();
    if (threadIdx.x == 0 && next_bytes) tc_request_b(x, next_off, next_bytes);       // the weight buffer is free again
}

__device__ __forceinline__ void mbar_arrive(uint64_t *mbar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" :: "r"(tc::smem_u32(mbar)) : "memory"); }
// Feature rows are gathered by tensor-map TMA, four rows per instruction (tile::gather4), into 128-byte staging rows with the
// 128-byte swizzle: chunk c (16 bytes) of a staging row sits at chunk c ^ (bits 7-9 of the row's shared-memory address), so 16-byte
// reads down a column of rows are conflict-free.  One table's staging = 128 rows x 128 bytes (tm_selftest_gather4 pins the layout).
constexpr int kStageRow = 32;     // floats per staged row piece
constexpr int kStageTable = 128 * kStageRow;
__device__ __forceinline__ const float *stage_piece(const float *tab, int row, int k) {      // columns [k, k+4) of the chunk, k % 4 == 0
    const float *p = tab + row * kStageRow;
    return p + (((k >> 2) ^ (int)((tc::smem_u32(p) >> 7) & 7u)) << 2);
}
__device__ __forceinline__ void tma_gather4(void *dst_smem, const CUtensorMap *map, int col, int r0, int r1, int r2, int r3, uint64_t *mbar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n"
                 :: "r"(tc::smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(tc::smem_u32(mbar)) : "memory");
}

constexpr int kMaxPeers = 7;
// Launch constants derived from the layout on the host: the kernel reads them from the parameter bank where it needs them instead of
// deriving them in its prologue and carrying them in registers through every round
struct TcDer {
    int H2, nG, nchS, colE, colY, colM0, colM1, colA3, nsl, dq, de, m3one, cpr, ed_vec, d_vec, jofs;
    int ev_K8, ev_nch, ne01, ne2;            // lin_event operand in use (edge-projection mode: its TimeEncode columns only): padded K, chunks, rounds at positions 0 / 1 and 2
    int bytes_e, bytes_g, bytes_sp, bytes_q, bytes_r, bytes_m3;
    int fw01_bytes, fw2_bytes;               // first weight request of a pass at positions 0 / 1 and at position 2
    int64_t ev_w, ev_chunk, fw01_off, fw2_off;
    int64_t n_rows, n_tiles;                 // tile rows (motifs, or first-hop slots of walk groups) and tiles of 128 rows
};
struct TcArgs {
    TcDer k;
    int64_t n_motifs, W, group;              // B * W motifs; W walks per root; roots per reference batch (index of std_)
    const int32_t *nodes, *eidx;
    const float *t;
    const uint8_t *cat;
    const float *cut, *eid, *node_feat, *edge_feat, *std_;
    int64_t n_node_rows, n_edge_rows;
    float *F;                                // scratch: per CTA 12 h slabs [position][column chunk][piece k/4][128 rows][4]
    float *Es;                               // drain mode (node_dim > 32): per CTA nG slabs [MLP.0 chunk][piece][128 rows][4] that hold lin_event's output
                                             //   (+ bias + edge-identity terms) between the lin_event rounds and the MLP.0 rounds, so that E and Zs | Zt
                                             //   can share TMEM columns and two tiles fit an SM; nullptr: E stays in TMEM
    float *scores;
    float *y_out;                            // optional [n_motifs, H]: relu(attention.MLP.0(.)), the input of attention.MLP.3 (enhance path)
    float *peer[kMaxPeers];                  // score gather fused into the kernel: the same [n_motifs] segment in each peer GPU's gathered buffer (NVLink peer stores)
    int n_peer;
    uint32_t tmem_cols;
    int b_bytes;                             // bytes of the weight-chunk buffer
    int stage_off;                           // byte offset of the node-feature staging (src rows, then tgt rows); 0: gather with plain loads
    int stage_edge_off;                      // byte offset of the edge-feature staging; 0: gather with plain loads
    unsigned long long *tile_counter;        // dynamic tile scheduler: CTA b takes tile b first, then gridDim.x + atomicAdd(counter, 1)
    int proj;                                // edge-projection mode: edge_feat is the table P = lin_event[:, :Ed] . edge features [n_edge_rows][D] (tc_project_edges);
                                             //   lin_event runs over its TimeEncode columns only and P's rows are gathered like a third node-feature table
    int nodes_vec;                           // nodes is 8-byte aligned: both endpoints of an event with one load
    int eid_u8;                              // eid points to byte counts (tm_edge_identity_u8) instead of floats
    int discard;                             // discard.global.L2 on the h slabs once a tile has consumed them (no write-back of the scratch)
    int share;                               // SHARE instantiations: consecutive walks per first-hop slot (find_k_walks: w = i1 * N2 + j, share = N2 >= 2, divides n_motifs)
    float *Ys;                               // SHARE: per CTA H / 32 slabs that hold P h_2 + cy of the tile's slots between the sub-tiles
    int sp_stage_safe;                       // the row staging lies behind the [S; P] weight chunk: staged rows may be in flight during the [S; P] rounds
    int dual;                                // two A buffers: bit 0 both orientations of MLP.0 per round + Q / R chunk pairs, bit 1 lin_event chunk pairs, bit 2 MLP.3 in one round
    long long *dbg;                          // TEMPME_TC_TIMING: 128 x 5 clock stamps of CTA 0
};

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }

// ---------------------------------------------------------------------------------------------
// score_tc_kernel: one persistent launch scores all motifs.  A CTA (256 threads = 128 motifs x 2 column halves, two
// CTAs per SM) takes a tile of 128 motifs through
//   event passes, one per walk position p (row = motif): lin_event -> E, MLP.0 for both orientations -> Zs, Zt,
//     h_p = relu(. + bias) written to the CTA's private 192 KB scratch (12 [128 x 32] slabs, L2 resident);
//   motif rounds: [S; P] h_2 -> scores -> temporal weights, softmax -> + Q mix -> R -> MLP.3 -> MLP.5 + sigmoid.
// While one CTA of the SM is in its (CUDA-core heavy) event passes the other is usually in its (tensor heavy)
// motif rounds.
// TMEM (256 columns at D <= 32): event passes Zs [0,H)  Zt [H,2H)  E [2H, 2H + D16), or E aliasing Zt when MLP.0 has a single
// K chunk (D <= 32): E has then been read completely before the MMAs that write Zt are issued; A2 | A = the top 128 columns.
// Motif rounds: U [0,2H) | Y [2H,3H) (one N = 3H accumulator of the [S; P] rounds) | A; then A2 = [0,64) over the dead U,
// M0 [H, H + M16) and M1 [0,H) (single-buffer mode: M0 [0,M16), M1 [2H,3H)).
// Tiles are handed out dynamically (CTA b starts with tile b, then takes gridDim.x + atomicAdd(counter)).
// Walk groups (SHARE instantiations, see the template's comment): a tile is 128 first-hop slots; order per tile = position-2 pass, [S; P] rounds,
// U + cu / P h_2 + cy parked in the CTA's scratch, then per sub-tile the passes of positions 0 and 1 (s_k taken in their epilogues) and the motif
// rounds with Y = Q mix + the parked P h_2 + cy.  The TMEM layouts of the phases are the ones above.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldcg4(const float *p) { return __ldcg(reinterpret_cast<const float4 *>(p)); }
// The CTA's L2 scratch (h slabs, U / Y, drained E): written and read back within microseconds while feature rows and walk tensors stream
// through L2.  Its lines carry the evict_last priority (createpolicy), so that the streams push each other out first (-11 % DRAM writes at
// cfg5, same kernel time; TM_SCRATCH_PLAIN restores plain .cg accesses).
#ifndef TM_SCRATCH_PLAIN
__device__ __forceinline__ uint64_t scratch_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ float4 scr_ld4(const float *p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.cg.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;\n" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ void scr_st4(float *p, float4 v, uint64_t pol) {
    asm volatile("st.global.cg.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;\n" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
#else
__device__ __forceinline__ uint64_t scratch_policy() { return 0; }
__device__ __forceinline__ float4 scr_ld4(const float *p, uint64_t) { return __ldcg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void scr_st4(float *p, float4 v, uint64_t) { __stcg(reinterpret_cast<float4 *>(p), v); }
#endif

template <int CW> __device__ __forceinline__ void tmem_ldw(uint32_t taddr, float *v);
template <> __device__ __forceinline__ void tmem_ldw<16>(uint32_t taddr, float *v) { tc::tmem_ld16(taddr, v); }
template <> __device__ __forceinline__ void tmem_ldw<8>(uint32_t taddr, float *v) { tc::tmem_ld8(taddr, v); }

// CW = columns of a K chunk per thread: 16 -> 256 threads, 8 -> 512 threads; TS = A operand in TMEM (else shared memory)
// SHARE: walk groups.  The s = a.share walks of a first-hop slot are consecutive (find_k_walks: w = i1 * N2 + j, utils/graph.py:290-300) and
// carry the same event next to the root: same edge id, endpoints and edge-identity counts, dt = 0.  A tile is then 128 SLOTS: the
// position-2 pass and the [S; P] rounds run once per slot, U + cu and P h_2 + cy leave TMEM for the CTA's L2 scratch (U over the dead h_2
// slabs), and s sub-tiles (motif = slot * s + j) follow with the passes of positions 0 / 1 and the motif rounds.  Every tile verifies the
// premise on its own operands (a CTA-wide vote); where it does not hold the position-2 work is simply repeated per sub-tile.
template <int CW, bool TS, bool DRAIN, bool SHARE>
__global__ void __launch_bounds__(128 * (kKC / CW), 2)
score_tc_kernel(const TcLayout L, const float *__restrict__ blob, const TcArgs a, const __grid_constant__ CUtensorMap tm_node,
                const __grid_constant__ CUtensorMap tm_edge) {
    constexpr int kParts = kKC / CW, kThreads = 128 * kParts;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[4];              // [0] MMAs done, [1] weight chunk landed, [2] staged node-feature rows landed, [3] staged edge-feature rows landed
    __shared__ uint32_t tmem_slot;
    __shared__ float part[kParts][3][128];
    __shared__ long long s_next_tile;
    const int t = threadIdx.x, warp = t >> 5, row = t & 127, prt = t >> 7, kb = CW * prt;
    TcCtx x;
    x.a = smem; x.a_s = tc::smem_u32(smem); x.b_s = x.a_s + (TS ? 0 : 2 * kATile); x.bars = bars; x.mma_phase = 0; x.b_phase = 0; x.blob = blob;
    x.dbg = blockIdx.x == 0 ? a.dbg : nullptr; x.dbg_i = 0;
    float *cstE = reinterpret_cast<float *>(smem + (TS ? 0 : 2 * kATile) + a.b_bytes), *cstM = cstE + L.n_cstE;
    uint2 *ctab = reinterpret_cast<uint2 *>(cstM + L.n_cstM);
    for (int i = t; i < L.n_cstE + L.n_cstM; i += kThreads) cstE[i] = __ldg(blob + L.cstE + i);      // cstM follows cstE in the blob
    cos_table_to_smem(ctab);
    if (t == 0) { tc::mbar_init(bars, 1); tc::mbar_init(bars + 1, 1); tc::mbar_init(bars + 2, kThreads / 32); tc::mbar_init(bars + 3, kThreads / 32); }
    if (warp == 0) tc::tmem_alloc(&tmem_slot, a.tmem_cols);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot, lane_base = (uint32_t)((warp & 3) * 32) << 16;
    x.tmem = tmem; x.a_col = a.tmem_cols - 2 * kKC; x.a_col2 = x.a_col - 2 * kKC;
    AFill<CW, TS> af;
    const int H = L.H, D = L.D, Ed = L.Ed;
    constexpr bool drain = DRAIN;                          // host: a.Es != nullptr exactly for the DRAIN instantiations
    // Launch constants (TcDer, filled by tc_encode_score): read from the parameter bank at their uses.
    // colE: E aliases Zt when MLP.0 has a single K chunk; dq / de: dual rounds (MLP.0 orientations, Q, R chunk pairs / lin_event chunk pairs);
    // m3one: MLP.3 in one round -- its two or three K chunks from A, A2 and a third buffer behind M0; M1 then lands over M0's first chunks
    // (hid_dim 32: U [0,64) | Y [64,96); M0 then starts at column 64, clear of the second A buffer [0,64))
#define H2 (a.k.H2)
#define nG (a.k.nG)
#define nchS (a.k.nchS)
#define colE (a.k.colE)
#define colY (a.k.colY)
#define colM0 (a.k.colM0)
#define colM1 (a.k.colM1)
#define colA3 (a.k.colA3)
#define nsl (a.k.nsl)
#define dq (a.k.dq != 0)
#define de (a.k.de != 0)
#define m3one (a.k.m3one != 0)
#define cpr (a.k.cpr)
#define ed_vec (a.k.ed_vec != 0)
#define d_vec (a.k.d_vec != 0)
#define jofs (a.k.jofs)
#define bytes_e (a.k.bytes_e)
#define bytes_g (a.k.bytes_g)
#define bytes_sp (a.k.bytes_sp)
#define bytes_q (a.k.bytes_q)
#define bytes_r (a.k.bytes_r)
#define bytes_m3 (a.k.bytes_m3)
#define n_rows (a.k.n_rows)
#define n_tiles (a.k.n_tiles)
    constexpr int colZ = 0, colU = 0;
    // SPILL: the [S; P] rounds run once per slot too and U / Y wait in the scratch; TM_SHARE_SP_PER_WALK (A/B) shares the position-2 pass only and
    // keeps the motif rounds of the per-walk evaluation ([S; P] from the h_2 slabs for every sub-tile, U and Y in TMEM)
#ifdef TM_SHARE_SP_PER_WALK
    constexpr bool SPILL = false;
#else
    constexpr bool SPILL = SHARE;
#endif
    const int sh = SHARE ? a.share : 1;                    // sub-tiles per tile
    const bool proj = a.proj != 0;
    // lin_event rounds of the pass at position pos (position 2: dt = 0, the pure TimeEncode chunks are in the bias; edge-projection mode: no round
    // at all), and the weights of a pass's first round
    auto n_evt_rounds = [&](int pos) { return pos == 2 ? a.k.ne2 : a.k.ne01; };
    auto first_w = [&](int pos, int64_t &off, int &bytes) {
        if (pos == 2) { off = a.k.fw2_off; bytes = a.k.fw2_bytes; } else { off = a.k.fw01_off; bytes = a.k.fw01_bytes; }
    };
    constexpr int kFirstPos = SHARE ? 2 : 0;               // a tile starts with this position's pass
    if (t == 0 && blockIdx.x < n_tiles) { int64_t o_; int b_; first_w(kFirstPos, o_, b_); tc_request_b(x, o_, b_); }
    float *Fs = a.F + (int64_t)blockIdx.x * 3 * nsl * kSlabFloats;      // this CTA's h slabs: [position][column chunk][piece k/4][128 rows][4]
    const float *F0 = Fs, *F1 = Fs + nsl * kSlabFloats, *F2 = Fs + 2 * nsl * kSlabFloats;
    float *Es = drain ? a.Es + (int64_t)blockIdx.x * nG * kSlabFloats : nullptr;
    float *Us = Fs + 2 * nsl * kSlabFloats;                // SHARE: U + cu over the h_2 slabs (same thread-to-address map as their reader's), P h_2 + cy behind
    float *Ys = SHARE ? a.Ys + (int64_t)blockIdx.x * (H / kKC) * kSlabFloats : nullptr;
    const uint64_t spol = scratch_policy();
    // this thread's CW columns [kb, kb+CW) of row `row` of a [128 x 32] slab (coalesced 16-byte pieces)
    auto ldw = [&](const float *slab, float *v) {
#pragma unroll
        for (int g = 0; g < CW / 4; ++g) { const float4 f = scr_ld4(slab + ((kb >> 2) + g) * 512 + row * 4, spol); v[4 * g] = f.x; v[4 * g + 1] = f.y; v[4 * g + 2] = f.z; v[4 * g + 3] = f.w; }
    };

    // Per-pass operands of this thread's row are fetched one pass ahead: the indices, dt and the edge-identity counts with
    // plain loads; the two endpoints' feature rows (one 128-byte piece per K chunk) by 1-D bulk TMA into padded staging
    // rows (thread part 0 requests the source row, part 1 the target row; every thread arrives on bars[2] once per request).
    struct PassIdx { int32_t e, ns, nt; float dt, ei0, ei1, ei2; };
    auto load_idx = [&](int64_t g, bool lv, int pos) {
        PassIdx q; q.e = -1; q.ns = -1; q.nt = -1; q.dt = 0.f; q.ei0 = q.ei1 = q.ei2 = 0.f;
        if (lv) {
            q.e = a.eidx[g * 3 + pos];
            if (a.nodes_vec) { const int2 nd = *reinterpret_cast<const int2 *>(a.nodes + g * 6 + 2 * pos); q.ns = nd.x; q.nt = nd.y; }      // both endpoints: one 8-byte load
            else { q.ns = a.nodes[g * 6 + 2 * pos]; q.nt = a.nodes[g * 6 + 2 * pos + 1]; }
            q.dt = __fsub_rn(a.t[g * 3 + 2], a.t[g * 3 + pos]);                          // explainer.py:326
            if (a.eid) {
                if (a.eid_u8) {                   // [motif][position][4] bytes: three counts and a pad byte, one 4-byte load
                    const uchar4 c4 = __ldg(reinterpret_cast<const uchar4 *>(a.eid) + g * 3 + pos);
                    q.ei0 = (float)c4.x; q.ei1 = (float)c4.y; q.ei2 = (float)c4.z;
                }
                else { const float *ei = a.eid + g * 9 + pos * 3; q.ei0 = __ldg(ei); q.ei1 = __ldg(ei + 1); q.ei2 = __ldg(ei + 2); }
            }
        }
        return q;
    };
    const bool stage_nodes = a.stage_off != 0;
    float *stg = reinterpret_cast<float *>(smem + a.stage_off);
    uint32_t n_phase = 0;
    // One mbarrier arrival per warp.  A requesting warp gathers its 32 rows with eight gather4 instructions (lanes 0-7, the row
    // indices come from the owning lanes by shuffle; rows that are out of range fetch row 0 and are ignored by the reader).
    auto request_rows = [&](uint64_t *bar, bool warp_requests, const CUtensorMap *map, int idx, float *tab, int col) {
        const int lane = t & 31, l4 = (lane & 7) * 4;
        const int r0 = __shfl_sync(0xffffffffu, idx, l4), r1 = __shfl_sync(0xffffffffu, idx, l4 + 1), r2 = __shfl_sync(0xffffffffu, idx, l4 + 2),
                  r3 = __shfl_sync(0xffffffffu, idx, l4 + 3);
        if (lane == 0) { if (warp_requests) tc::mbar_expect_tx(bar, 8u * 4u * kStageRow * 4u); else mbar_arrive(bar); }
        __syncwarp();
        if (warp_requests && lane < 8) tma_gather4(tab + ((warp & 3) * 32 + l4) * kStageRow, map, col, r0, r1, r2, r3, bar);
    };
    auto request_nodes = [&](const PassIdx &q, int c) {       // chunk c of both endpoints' rows (:348-351): part 0 source rows, part 1 target rows
        if (!stage_nodes) return;
        const int idx = prt == 0 ? q.ns : q.nt;
        request_rows(bars + 2, prt < 2, &tm_node, (idx >= 0 && idx < a.n_node_rows) ? idx : 0, stg + (prt & 1) * kStageTable, c * kKC);
    };
    const bool stage_edges = a.stage_edge_off != 0;
    float *stg_e = reinterpret_cast<float *>(smem + a.stage_edge_off);
    uint32_t e_phase = 0;
    auto request_edges = [&](const PassIdx &q, int c) {       // chunk c of the edge-feature rows (:332-338); part 0 requests
        if (!stage_edges) return;
        request_rows(bars + 3, prt == 0, &tm_edge, (q.e >= 0 && q.e < a.n_edge_rows) ? q.e : 0, stg_e, c * kKC);
    };
    PassIdx pcur, pnext;
    {
        const int64_t r0_ = (int64_t)blockIdx.x * 128 + row;
        const bool lv = blockIdx.x < n_tiles && r0_ < n_rows;
        pcur = load_idx(lv ? r0_ * sh : 0, lv, kFirstPos);
        pnext = pcur;
        if (blockIdx.x < n_tiles) { request_nodes(pcur, 0); request_edges(pcur, 0); }
    }
    // SHARE: do the sh walks of this thread's slot (first motif g0_, walk 0's position-2 operands in p0) carry the same event next to the root?
    // Thread part 0 compares the ids, part 1 the edge-identity counts; the loads of up to four walks are in flight together.
    auto group_eq = [&](const int64_t g0_, const PassIdx &p0) -> int {
        int eq = 1;
        if (prt == 0) {
#pragma unroll 1
            for (int j = 1; j < sh; j += 4) {
                int32_t e_[4], s_[4], t_[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int64_t g = g0_ + min(j + i, sh - 1);
                    e_[i] = a.eidx[g * 3 + 2];
                    s_[i] = a.nodes[g * 6 + 4]; t_[i] = a.nodes[g * 6 + 5];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) eq &= (int)(e_[i] == p0.e) & (int)(s_[i] == p0.ns) & (int)(t_[i] == p0.nt);
            }
        } else if (prt == 1 && a.eid) {
#pragma unroll 1
            for (int j = 1; j < sh; j += 4) {
                float c_[4][3];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int64_t gg = g0_ + min(j + i, sh - 1);
                    if (a.eid_u8) { const uchar4 c4 = __ldg(reinterpret_cast<const uchar4 *>(a.eid) + gg * 3 + 2); c_[i][0] = (float)c4.x; c_[i][1] = (float)c4.y; c_[i][2] = (float)c4.z; }
                    else { const float *ei = a.eid + gg * 9 + 6; c_[i][0] = __ldg(ei); c_[i][1] = __ldg(ei + 1); c_[i][2] = __ldg(ei + 2); }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) eq &= (int)(c_[i][0] == p0.ei0) & (int)(c_[i][1] == p0.ei1) & (int)(c_[i][2] == p0.ei2);
            }
        }
        return eq;
    };
    int eq = 1;                                 // this row's verdict for the tile that starts next (voted at the closing barrier of its position-2 pass)
    if (SHARE) {
        const int64_t r0_ = (int64_t)blockIdx.x * 128 + row;
        if (blockIdx.x < n_tiles && r0_ < n_rows) eq = group_eq(r0_ * sh, pcur);
    }
    float sk0 = 0.f, sk1 = 0.f;                 // SHARE: this thread's share of h_0 . (U + cu) and h_1 . (U + cu), accumulated by the passes' epilogues

    // =========================== event pass ===========================
    // One walk position of the tile's rows: lin_event -> E, MLP.0 for both orientations -> Zs, Zt, h_pos = relu(. + bias) -> the CTA's h slabs.
    // pi: the rows' operands (their first staged chunks are in flight); has_next: another event pass follows directly -- its operands are in
    // pnext and its first chunks are requested behind this pass's last rounds; (after_off, after_bytes): weights of the round that follows the
    // pass.  Returns the CTA-wide AND of pred, taken at the pass's closing barrier.
    auto event_pass = [&](const int pos, const PassIdx pi, const bool live, const bool has_next, const int64_t after_off, const int after_bytes, const int pred) -> int {
        x.a_col2 = x.a_col - 2 * kKC;
        const int nE = n_evt_rounds(pos);
        const bool have_E = nE > 0;
        const bool e_ok = pi.e >= 0 && pi.e < a.n_edge_rows, s_ok = pi.ns >= 0 && pi.ns < a.n_node_rows, t_ok = pi.nt >= 0 && pi.nt < a.n_node_rows;
        const float *ef = a.edge_feat + (int64_t)max(pi.e, 0) * Ed, *sf = a.node_feat + (int64_t)max(pi.ns, 0) * D, *tf = a.node_feat + (int64_t)max(pi.nt, 0) * D;
        auto xval = [&](int j) -> float {                                              // [edge features | TimeEncode] column j (:179, :55-58)
            if (j < Ed) return e_ok ? __ldg(ef + j) : 0.f;
            const int k = j - Ed;
            if (k < D) return live ? cos_accurate(__fadd_rn(__fmul_rn(pi.dt, cstE[L.e_freq + k]), cstE[L.e_phase + k]), ctab) : 0.f;
            return 0.f;
        };
        // this thread's columns of chunk cc are all TimeEncode columns (straight-line code, the cosines interleave)
        auto time_chunk = [&](int cc) {
            const int j0 = cc * kKC + kb + jofs;
            return j0 >= Ed && ((j0 - Ed) & 3) == 0 && j0 + CW <= Ed + L.D16 && kb + CW <= min(kKC, a.k.ev_K8 - cc * kKC);
        };
        auto cos_chunk = [&](int cc, float *w) {
            const int j0 = cc * kKC + kb + jofs;
#pragma unroll
            for (int g = 0; g < CW / 4; ++g) {
                const float4 fq = lds4(cstE + L.e_freq + (j0 - Ed) + 4 * g), ph = lds4(cstE + L.e_phase + (j0 - Ed) + 4 * g);      // zero-padded to D16
                w[4 * g] = cos_accurate(__fadd_rn(__fmul_rn(pi.dt, fq.x), ph.x), ctab); w[4 * g + 1] = cos_accurate(__fadd_rn(__fmul_rn(pi.dt, fq.y), ph.y), ctab);
                w[4 * g + 2] = cos_accurate(__fadd_rn(__fmul_rn(pi.dt, fq.z), ph.z), ctab); w[4 * g + 3] = cos_accurate(__fadd_rn(__fmul_rn(pi.dt, fq.w), ph.w), ctab);
            }
        };
        // ---- lin_event (:93) -> E.  fill_evt(cc, second): this thread's columns of chunk cc of [edge | TimeEncode] into an A buffer
        auto fill_evt = [&](int cc, bool second) {
            const int kcols = min(kKC, a.k.ev_K8 - cc * kKC);
            const int j0 = cc * kKC + kb + jofs;                // this thread's columns [j0, j0 + CW) of [edge | TimeEncode]
            if (ed_vec && j0 + CW <= Ed) {                      // all edge features (warp-uniform)
#pragma unroll
                for (int g = 0; g < CW / 4; ++g)
                    af.put4(x, row, kb, 4 * g, !e_ok ? make_float4(0.f, 0.f, 0.f, 0.f) : stage_edges ? lds4(stage_piece(stg_e, row, kb + 4 * g)) : ldg4(ef + j0 + 4 * g));
                af.commit(x, lane_base, kb, second);
            } else if (time_chunk(cc)) {                        // all TimeEncode
                float w[CW];
                cos_chunk(cc, w);
                // no masks: columns >= D have zero frequency / phase (cos = 1) and zero weights; rows past the last motif are never stored
#pragma unroll
                for (int g = 0; g < CW / 4; ++g) af.put4(x, row, kb, 4 * g, make_float4(w[4 * g], w[4 * g + 1], w[4 * g + 2], w[4 * g + 3]));
                af.commit(x, lane_base, kb, second);
            } else {                                            // mixed columns
#pragma unroll
                for (int g = 0; g < CW / 4; ++g) {
                    const int k = kb + 4 * g, j = cc * kKC + k + jofs;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (k < kcols) {
                        if (ed_vec && j + 3 < Ed) v = !e_ok ? make_float4(0.f, 0.f, 0.f, 0.f) : stage_edges ? lds4(stage_piece(stg_e, row, k)) : ldg4(ef + j);
                        else v = make_float4(xval(j), xval(j + 1), xval(j + 2), xval(j + 3));
                    }
                    if (TS || k < kcols) af.put4(x, row, kb, 4 * g, v);
                }
                if (kb < kcols) af.commit(x, lane_base, kb, second);
            }
        };
        for (int c = 0; c < nE; c += cpr) {
            const int cnt = min(cpr, nE - c);
            const bool has_edge = !proj && c < L.nch_edge;      // pairs are only formed when at most the first chunk holds edge columns
            if (stage_edges && has_edge) { tc::mbar_wait(bars + 3, e_phase); e_phase ^= 1; }        // chunk c of the edge rows has landed
            fill_evt(c, false);
            if (cnt == 2) fill_evt(c + 1, true);
            const int kc0 = min(kKC, a.k.ev_K8 - c * kKC), kc1 = cnt == 2 ? min(kKC, a.k.ev_K8 - (c + 1) * kKC) : 0;
            const int left = nE - c - cnt;
            tc_mma_round<TS>(x, L.D16, kc0, colE, c != 0, left > 0 ? a.k.ev_w + (int64_t)(c + cnt) * a.k.ev_chunk : L.g0.w,
                             left > 0 ? min(cpr, left) * bytes_e : bytes_g,
                             [&]() {                            // the staged chunk has been consumed: the next one lands behind the MMAs
                                 if (has_edge) {
                                     if (c + 1 < L.nch_edge) request_edges(pi, c + 1);
                                     else if (has_next) request_edges(pnext, 0);
                                 }
                             }, Dual{cnt == 2 ? kDualK : kSingle, kc1, 0, 0, 0});
        }
        // ---- event_conv.MLP.0 on src + relu(tgt + event) (o = 0) and tgt + relu(src + event) (o = 1) (:94-95, :182-184) -> Zs, Zt
        const int eb = pos == 2 ? (proj ? L.e_b2p : L.e_b2) : L.e_b;
        // lin_event's bias and the three edge-identity columns (applied on the CUDA cores) for this thread's columns of MLP.0 chunk c
        auto finish_e = [&](float *ee, int c) {
#pragma unroll
            for (int k = 0; k < CW; k += 4) {
                const int j0 = c * kKC + kb + k;                 // < D16: constants are zero-padded
                const float4 bb = lds4(cstE + eb + j0), w0 = lds4(cstE + L.e_wi + j0), w1 = lds4(cstE + L.e_wi + L.D16 + j0), w2 = lds4(cstE + L.e_wi + 2 * L.D16 + j0);
                ee[k] = fmaf(w2.x, pi.ei2, fmaf(w1.x, pi.ei1, fmaf(w0.x, pi.ei0, ee[k] + bb.x)));
                ee[k + 1] = fmaf(w2.y, pi.ei2, fmaf(w1.y, pi.ei1, fmaf(w0.y, pi.ei0, ee[k + 1] + bb.y)));
                ee[k + 2] = fmaf(w2.z, pi.ei2, fmaf(w1.z, pi.ei1, fmaf(w0.z, pi.ei0, ee[k + 2] + bb.z)));
                ee[k + 3] = fmaf(w2.w, pi.ei2, fmaf(w1.w, pi.ei1, fmaf(w0.w, pi.ei0, ee[k + 3] + bb.w)));
            }
        };
        if (drain && have_E) {
            // E leaves TMEM: every thread parks the columns it will consume in the MLP.0 rounds in its own rows of the CTA's E scratch
            // (L2; written and read back by the same thread), then Zs | Zt take over the columns
#pragma unroll 1
            for (int c = 0; c < nG; ++c) {
                if (kb < min(kKC, L.g0.K8 - c * kKC)) {
                    float ee[CW];
                    tmem_ldw<CW>(tmem + lane_base + colE + c * kKC + kb, ee);
                    finish_e(ee, c);
                    float *ec = Es + c * kSlabFloats + (kb >> 2) * 512 + row * 4;
#pragma unroll
                    for (int g = 0; g < CW / 4; ++g) scr_st4(ec + g * 512, make_float4(ee[4 * g], ee[4 * g + 1], ee[4 * g + 2], ee[4 * g + 3]), spol);
                }
            }
            tc::fence_before_sync();
            __syncthreads();
            tc::fence_after_sync();
        }
        const bool with_u = SPILL && pos < 2;
        float4 ua[4], ub[4];                                    // SHARE: U + cu of the epilogue's first two column groups
#pragma unroll
        for (int i = 0; i < 4; ++i) ua[i] = ub[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        auto load_u = [&](const int g, float4 *u) {
            const int c0 = (H2 / kParts) * prt + g * 16;
            const float *uc = Us + (c0 >> 5) * kSlabFloats + ((c0 & 31) >> 2) * 512 + row * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) u[i] = scr_ld4(uc + i * 512, spol);
        };
        auto g0_chunk = [&](const int c, const bool last) __attribute__((always_inline)) {      // one K chunk of MLP.0
            const int kcols = min(kKC, L.g0.K8 - c * kKC);
            if (stage_nodes) { tc::mbar_wait(bars + 2, n_phase); n_phase ^= 1; }          // chunk c of the rows has landed
            if (proj && stage_edges) { tc::mbar_wait(bars + 3, e_phase); e_phase ^= 1; }  // and chunk c of the projected edge rows
            float sv[CW], gv[CW], ee[CW];                       // endpoints' features and lin_event output + bias + edge-identity terms
            auto load_g = [&]() {
#pragma unroll
                for (int k = 0; k < CW; k += 4) {
                    const int j = c * kKC + kb + k;
                    float4 s4, g4;
                    if (stage_nodes) {
                        s4 = s_ok ? lds4(stage_piece(stg, row, kb + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        g4 = t_ok ? lds4(stage_piece(stg + kStageTable, row, kb + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    } else if (d_vec && j + 3 < D) {
                        s4 = s_ok ? ldg4(sf + j) : make_float4(0.f, 0.f, 0.f, 0.f); g4 = t_ok ? ldg4(tf + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    } else {
                        float w[8];
#pragma unroll
                        for (int i = 0; i < 4; ++i) { w[i] = (j + i < D && s_ok) ? __ldg(sf + j + i) : 0.f; w[4 + i] = (j + i < D && t_ok) ? __ldg(tf + j + i) : 0.f; }
                        s4 = make_float4(w[0], w[1], w[2], w[3]); g4 = make_float4(w[4], w[5], w[6], w[7]);
                    }
                    sv[k] = s4.x; sv[k + 1] = s4.y; sv[k + 2] = s4.z; sv[k + 3] = s4.w; gv[k] = g4.x; gv[k + 1] = g4.y; gv[k + 2] = g4.z; gv[k + 3] = g4.w;
                }
                if (!have_E) {
#pragma unroll
                    for (int i = 0; i < CW; ++i) ee[i] = 0.f;
                    finish_e(ee, c);
                } else if (drain) ldw(Es + c * kSlabFloats, ee);
                else { tmem_ldw<CW>(tmem + lane_base + colE + c * kKC + kb, ee); finish_e(ee, c); }
                if (proj && e_ok) {                             // + lin_event[:, :Ed] . edge features, precomputed per edge id
#pragma unroll
                    for (int k = 0; k < CW; k += 4) {
                        const int j = c * kKC + kb + k;
                        float4 p4;
                        if (stage_edges) p4 = lds4(stage_piece(stg_e, row, kb + k));       // columns past D arrive as zeros (tensor-map bounds)
                        else if (d_vec && j + 3 < D) p4 = ldg4(a.edge_feat + (int64_t)pi.e * D + j);
                        else { const float *pr = a.edge_feat + (int64_t)pi.e * D; p4 = make_float4(j < D ? __ldg(pr + j) : 0.f, j + 1 < D ? __ldg(pr + j + 1) : 0.f, j + 2 < D ? __ldg(pr + j + 2) : 0.f, j + 3 < D ? __ldg(pr + j + 3) : 0.f); }
                        ee[k] += p4.x; ee[k + 1] += p4.y; ee[k + 2] += p4.z; ee[k + 3] += p4.w;
                    }
                }
            };
            auto put_z = [&](int o, bool second) {              // orientation o: p + relu(q + event)
#pragma unroll
                for (int k = 0; k < CW; k += 4) {
                    float z[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float p_ = o ? gv[k + i] : sv[k + i], q_ = o ? sv[k + i] : gv[k + i];
                        z[i] = p_ + fmaxf(q_ + ee[k + i], 0.f);     // columns >= D: features, lin_event output and constants are all zero (padding)
                    }
                    af.put4(x, row, kb, k, make_float4(z[0], z[1], z[2], z[3]));
                }
                af.commit(x, lane_base, kb, second);
            };
            auto next_nodes = [&]() {                           // the staged chunk has been consumed
                if (c + 1 < nG) { request_nodes(pi, c + 1); if (proj) request_edges(pi, c + 1); }
                else if (has_next) { request_nodes(pnext, 0); if (proj) request_edges(pnext, 0); }
            };
            int64_t noff; int nbytes;                           // weights after this chunk's last round
            if (!last) { noff = L.g0.w + (int64_t)(c + 1) * chunk_floats(L.g0); nbytes = bytes_g; }
            else { noff = after_off; nbytes = after_bytes; }
            if (a.dual) {                                       // both orientations in one round: Zs from the first A buffer, Zt from the second
                if (kb < kcols) { load_g(); put_z(0, false); put_z(1, true); }
                tc_mma_round<TS>(x, H, kcols, colZ, c != 0, noff, nbytes, next_nodes, Dual{kDualM, 0, colZ + H, 0, 0});
            } else {
#pragma unroll 1
                for (int o = 0; o < 2; ++o) {
                    if (kb < kcols) { load_g(); put_z(o, false); }
                    if (o == 0) tc_mma_round<TS>(x, H, kcols, colZ, c != 0, L.g0.w + (int64_t)c * chunk_floats(L.g0), bytes_g);
                    else tc_mma_round<TS>(x, H, kcols, colZ + H, c != 0, noff, nbytes, next_nodes);
                }
            }
        };
#pragma unroll 1
        for (int c = 0; c < nG; ++c) g0_chunk(c, c + 1 == nG);
        // ---- h_pos = relu(MLP.0 + bias): thread part p owns columns [4 CW p, 4 CW (p + 1)) of [Zs | Zt]
        // (SHARE, positions 0 / 1: U + cu is known already, so s_pos = h_pos . (U + cu) is taken here and the motif rounds do not read h back for
        // it.  U comes from the L2 scratch 32 columns at a time; groups beyond the first two exist at hid_dim 64 only -- one uniform branch
        // around their loads and uses, so that the values stay in registers.)
        float sk = 0.f;
        auto ep_group = [&](const int g, const float4 *u) __attribute__((always_inline)) {
            const int c0 = (H2 / kParts) * prt + g * 16;
            float v[16];
            tc::tmem_ld16(tmem + lane_base + colZ + c0, v);
            float *fc = Fs + (pos * nsl + (c0 >> 5)) * kSlabFloats + ((c0 & 31) >> 2) * 512 + row * 4;
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 bb = lds4(cstE + L.e_g0b + (c0 & (H - 1)) + i);
                const float4 hv = make_float4(fmaxf(v[i] + bb.x, 0.f), fmaxf(v[i + 1] + bb.y, 0.f), fmaxf(v[i + 2] + bb.z, 0.f), fmaxf(v[i + 3] + bb.w, 0.f));
                scr_st4(fc + (i >> 2) * 512, hv, spol);
                if (SPILL) {
                    const float4 uu = u[i >> 2];
                    sk = fmaf(hv.x, uu.x, sk); sk = fmaf(hv.y, uu.y, sk); sk = fmaf(hv.z, uu.z, sk); sk = fmaf(hv.w, uu.w, sk);
                }
            }
        };
        {
            if (with_u) { load_u(0, ua); load_u(1, ub); }
            ep_group(0, ua); ep_group(1, ub);
            if (H2 / kParts > 32) {
                if (with_u) { load_u(2, ua); load_u(3, ub); }
                ep_group(2, ua); ep_group(3, ub);
            }
        }
        if (SPILL) { if (pos == 0) sk0 = sk; else if (pos == 1) sk1 = sk; }
        tc::fence_before_sync();
        const int vote = __syncthreads_and(pred);   // TMEM reads done before the next rounds overwrite E / Z; the h slabs are visible to the CTA
        tc::fence_after_sync();
        return vote;
    };

    // =========================== [U | Y] = [S; P] h_2 ; returns this thread's share of r = d . h_2 ===========================
    auto sp_rounds = [&](const int64_t after_off, const int after_bytes) -> float {
        float rp = 0.f;
        float nxt[CW];
        ldw(F2, nxt);
        for (int c = 0; c < nchS; ++c) {
            float cur[CW];
#pragma unroll
            for (int i = 0; i < CW; ++i) cur[i] = nxt[i];
            if (c + 1 < nchS) ldw(F2 + (c + 1) * kSlabFloats, nxt);
#pragma unroll
            for (int k = 0; k < CW; k += 4) {
                const float4 dd = lds4(cstM + L.m_d + c * kKC + kb + k);
                rp = fmaf(dd.x, cur[k], rp); rp = fmaf(dd.y, cur[k + 1], rp); rp = fmaf(dd.z, cur[k + 2], rp); rp = fmaf(dd.w, cur[k + 3], rp);
                af.put4(x, row, kb, k, make_float4(cur[k], cur[k + 1], cur[k + 2], cur[k + 3]));
            }
            af.commit(x, lane_base, kb);
            const bool last = c + 1 == nchS;
            tc_mma_round<TS>(x, 3 * H, kKC, colU, c != 0, last ? after_off : L.sp.w + (int64_t)(c + 1) * chunk_floats(L.sp), last ? after_bytes : bytes_sp);
        }
        return rp;
    };
    // SHARE: U + cu and P h_2 + cy leave TMEM (every thread parks the columns it will read back itself: no barrier on the scratch)
    // U: the columns whose h the thread finishes in the event passes' epilogues (part p: [2H / parts * p, ...)), at the addresses of the same
    // (row, column) of the h_2 slabs; Y: the columns of the thread's A-fills (put_y).
    auto spill_uy = [&]() {
#pragma unroll 1
        for (int c0 = (H2 / kParts) * prt; c0 < (H2 / kParts) * (prt + 1); c0 += 16) {
            float v[16];
            tc::tmem_ld16(tmem + lane_base + colU + c0, v);
            float *dst = Us + (c0 >> 5) * kSlabFloats + ((c0 & 31) >> 2) * 512 + row * 4;
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 bb = lds4(cstM + L.m_cu + c0 + i);
                scr_st4(dst + (i >> 2) * 512, make_float4(v[i] + bb.x, v[i + 1] + bb.y, v[i + 2] + bb.z, v[i + 3] + bb.w), spol);
            }
        }
#pragma unroll 1
        for (int c = 0; c < H / kKC; ++c) {
            float v[CW];
            tmem_ldw<CW>(tmem + lane_base + colY + c * kKC + kb, v);
            float *dst = Ys + c * kSlabFloats + (kb >> 2) * 512 + row * 4;
#pragma unroll
            for (int g = 0; g < CW / 4; ++g) {
                const float4 bb = lds4(cstM + L.m_cy + c * kKC + kb + 4 * g);
                scr_st4(dst + g * 512, make_float4(v[4 * g] + bb.x, v[4 * g + 1] + bb.y, v[4 * g + 2] + bb.z, v[4 * g + 3] + bb.w), spol);
            }
        }
        tc::fence_before_sync();
        __syncthreads();                 // TMEM reads done before the next pass's MMAs overwrite U | Y
        tc::fence_after_sync();
    };

    // =========================== motif rounds of one (sub-)tile ===========================
    // scores -> temporal weights, softmax -> Q mix -> R -> MLP.3 -> MLP.5 + sigmoid.  (nxt_g, nxt_live, nxt_pos): the rows of the event pass
    // that follows (has_next); their operands are loaded into pcur behind the Q rounds and their first staged chunks requested behind MLP.3.
    auto motif_rounds = [&](const int64_t gm, const bool live, const float rp, const bool has_next, const int64_t nxt_g, const bool nxt_live, const int nxt_pos, const bool new_tile) {
        // ---- s_k = h_k . (U + cu) + r  (:806-808 after folding)
        float s0 = SPILL ? sk0 : 0.f, s1 = SPILL ? sk1 : 0.f;       // SHARE: taken by the event passes' epilogues
#pragma unroll 1
        for (int c = 0; c < (SPILL ? 0 : nchS); c += 2) {
            float p0[2][CW], p1[2][CW];
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) { ldw(F0 + (c + cc) * kSlabFloats, p0[cc]); ldw(F1 + (c + cc) * kSlabFloats, p1[cc]); }
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                float u[CW];
                tmem_ldw<CW>(tmem + lane_base + colU + (c + cc) * kKC + kb, u);
#pragma unroll
                for (int k = 0; k < CW; k += 4) {
                    const float4 cu = lds4(cstM + L.m_cu + (c + cc) * kKC + kb + k);
                    const float u0 = u[k] + cu.x, u1 = u[k + 1] + cu.y, u2 = u[k + 2] + cu.z, u3 = u[k + 3] + cu.w;
                    s0 = fmaf(p0[cc][k], u0, s0); s0 = fmaf(p0[cc][k + 1], u1, s0); s0 = fmaf(p0[cc][k + 2], u2, s0); s0 = fmaf(p0[cc][k + 3], u3, s0);
                    s1 = fmaf(p1[cc][k], u0, s1); s1 = fmaf(p1[cc][k + 1], u1, s1); s1 = fmaf(p1[cc][k + 2], u2, s1); s1 = fmaf(p1[cc][k + 3], u3, s1);
                }
            }
        }
        float n0[CW], n1[CW];                       // mix operands of the first Q round, in flight across the reduction
        ldw(F0, n0); ldw(F1, n1);
        part[prt][0][row] = s0; part[prt][1][row] = s1; part[prt][2][row] = rp;
        __syncthreads();
        {
            float r_ = cstM[L.m_e];
            s0 = 0.f; s1 = 0.f;
#pragma unroll
            for (int q = 0; q < kParts; ++q) { s0 += part[q][0][row]; s1 += part[q][1][row]; r_ += part[q][2][row]; }
            s0 += r_; s1 += r_;
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
        // ---- Y (+)= Q (alpha_0 h_0 + alpha_1 h_1)   (:841-843 after folding; SHARE: P h_2 + cy is added from the scratch afterwards)
        auto put_mix = [&](const float *c0, const float *c1, bool second) {
#pragma unroll
            for (int k = 0; k < CW; k += 4)
                af.put4(x, row, kb, k, make_float4(fmaf(al0, c0[k], al1 * c1[k]), fmaf(al0, c0[k + 1], al1 * c1[k + 1]),
                                                   fmaf(al0, c0[k + 2], al1 * c1[k + 2]), fmaf(al0, c0[k + 3], al1 * c1[k + 3])));
            af.commit(x, lane_base, kb, second);
        };
        float yb[2][CW];                                        // SHARE: P h_2 + cy of this thread's put_y columns, in flight across the last Q round
        auto load_yb = [&]() {
#pragma unroll
            for (int c = 0; c < 2; ++c) ldw(Ys + min(c, H / kKC - 1) * kSlabFloats, yb[c]);     // hid_dim 32: one chunk (the copy in yb[1] is not read)
        };
        if (dq) {                                               // two K chunks per round
            x.a_col2 = colU;
            auto q_round = [&](const int c, const bool LAST) __attribute__((always_inline)) {
                float m0[CW], m1[CW];
                ldw(F0 + (c + 1) * kSlabFloats, m0); ldw(F1 + (c + 1) * kSlabFloats, m1);
                put_mix(n0, n1, false);
                if (!LAST) { ldw(F0 + (c + 2) * kSlabFloats, n0); ldw(F1 + (c + 2) * kSlabFloats, n1); }
                put_mix(m0, m1, true);
                tc_mma_round<TS>(x, H, kKC, colY, !SPILL || c != 0, LAST ? L.r.w : L.q.w + (int64_t)(c + 2) * chunk_floats(L.q), LAST ? min(2, L.r.nch) * bytes_r : 2 * bytes_q,
                                 NoMid(), Dual{kDualK, kKC, 0, 0, 0});
            };
#pragma unroll 1
            for (int c = 0; c < nchS; c += 2) q_round(c, c + 2 >= nchS);
        } else {
            for (int c = 0; c < nchS; ++c) {
                float c0[CW], c1[CW];
#pragma unroll
                for (int i = 0; i < CW; ++i) { c0[i] = n0[i]; c1[i] = n1[i]; }
                if (c + 1 < nchS) { ldw(F0 + (c + 1) * kSlabFloats, n0); ldw(F1 + (c + 1) * kSlabFloats, n1); }
                put_mix(c0, c1, false);
                const bool last = c + 1 == nchS;
                tc_mma_round<TS>(x, H, kKC, colY, !SPILL || c != 0, last ? L.r.w : L.q.w + (int64_t)(c + 1) * chunk_floats(L.q), last ? bytes_r : bytes_q);
            }
        }
        // The tile's h slabs have been consumed (their last readers were the Q rounds' fills, completed before the rounds' barriers):
        // drop the dirty L2 lines instead of letting them be written back to HBM.  A 128-byte line = 8 rows x one 16-byte piece, read
        // only by this warp; the lane of row 8j discards it.  (SHARE: the slabs of position 2 hold U for the sub-tiles to come.)
        if (a.discard && (t & 7) == 0) {
#pragma unroll 1
            for (int sI = 0; sI < (SHARE ? 2 : 3) * nsl; ++sI)
#pragma unroll
                for (int g = 0; g < CW / 4; ++g)
                    asm volatile("discard.global.L2 [%0], 128;\n" :: "l"(Fs + sI * kSlabFloats + ((kb >> 2) + g) * 512 + row * 4) : "memory");
        }
        // ---- M0 = R relu(Y + cy)   (attention.MLP.3 and MLP.0 folded); the next pass's operands start to arrive
        pcur = load_idx(nxt_live ? nxt_g : 0, nxt_live, nxt_pos);
        if (SPILL) load_yb();
        auto put_y = [&](const int c, bool second) {           // c: 0 or 1 (hid_dim <= 64)
            float z[CW];
            tmem_ldw<CW>(tmem + lane_base + colY + c * kKC + kb, z);
#pragma unroll
            for (int k = 0; k < CW; k += 4) {
                float4 bb;
                if (SPILL) bb = c == 0 ? make_float4(yb[0][k], yb[0][k + 1], yb[0][k + 2], yb[0][k + 3]) : make_float4(yb[1][k], yb[1][k + 1], yb[1][k + 2], yb[1][k + 3]);
                else bb = lds4(cstM + L.m_cy + c * kKC + kb + k);
                const float4 yv = make_float4(fmaxf(z[k] + bb.x, 0.f), fmaxf(z[k + 1] + bb.y, 0.f), fmaxf(z[k + 2] + bb.z, 0.f), fmaxf(z[k + 3] + bb.w, 0.f));
                af.put4(x, row, kb, k, yv);
                if (a.y_out && live) *reinterpret_cast<float4 *>(a.y_out + gm * H + c * kKC + kb + k) = yv;
            }
            af.commit(x, lane_base, kb, second);
        };
        if (dq) {                                               // both K chunks in one round; M0 lands at colM0 (Y has been read completely)
            put_y(0, false);
            if (L.r.nch > 1) put_y(1, true);
            tc_mma_round<TS>(x, L.M16, kKC, colM0, false, L.m3.w, (m3one ? L.m3.nch : 1) * bytes_m3, NoMid(),
                             Dual{L.r.nch > 1 ? kDualK : kSingle, L.r.nch > 1 ? kKC : 0, 0, 0, 0});
        } else {
            for (int c = 0; c < L.r.nch; ++c) {
                put_y(c, false);
                const bool last = c + 1 == L.r.nch;
                tc_mma_round<TS>(x, L.M16, kKC, colM0, c != 0, last ? L.m3.w : L.r.w + (int64_t)(c + 1) * chunk_floats(L.r), last ? bytes_m3 : bytes_r);
            }
        }
        // ---- M1 = MLP.3 relu(M0 + cm[category])   (:199)
        int64_t noff = 0; int nbytes = 0;                       // weights of the pass that follows
        if (has_next) first_w(nxt_pos, noff, nbytes);
        const float *cmr = blob + L.cm + (int64_t)((L.if_cat && live && a.cat) ? min((int)a.cat[gm], 11) : 0) * L.M16;
        auto fill_m3 = [&](int c, uint32_t col) {
            const int kcols = min(kKC, L.m3.K8 - c * kKC);
            if (kb < kcols) {
                float z[CW];
                tmem_ldw<CW>(tmem + lane_base + colM0 + c * kKC + kb, z);      // columns < M16 (kcols is a multiple of 8, M16 of 16)
#pragma unroll
                for (int k = 0; k < CW; k += 4) {
                    const float4 bb = ldg4(cmr + c * kKC + kb + k);
                    // columns in [M, M16): R's rows and cm are zero-padded, relu(0) = 0
                    af.put4(x, row, kb, k, make_float4(fmaxf(z[k] + bb.x, 0.f), fmaxf(z[k + 1] + bb.y, 0.f), fmaxf(z[k + 2] + bb.z, 0.f), fmaxf(z[k + 3] + bb.w, 0.f)));
                }
                af.commit_at(x, lane_base, kb, col);
            }
        };
        if (m3one) {
            const int nc = L.m3.nch;
            fill_m3(0, x.a_col);
            if (nc > 1) fill_m3(1, x.a_col2);
            if (nc > 2) fill_m3(2, (uint32_t)colA3);
            tc_mma_round<TS>(x, H, min(kKC, L.m3.K8), colM1, false, noff, nbytes,
                             [&]() {
                                 if (has_next) { request_nodes(pcur, 0); request_edges(pcur, 0); }
                                 if (SHARE && new_tile) eq = nxt_live ? group_eq(nxt_g, pcur) : 1;       // the next tile's verdict, behind the MMAs
                             },
                             Dual{nc > 1 ? kDualK : kSingle, nc > 1 ? min(kKC, L.m3.K8 - kKC) : 0, 0, nc > 2 ? L.m3.K8 - 2 * kKC : 0, colA3});
        } else
        for (int c = 0; c < L.m3.nch; ++c) {
            const int kcols = min(kKC, L.m3.K8 - c * kKC);
            fill_m3(c, x.a_col);
            const bool last = c + 1 == L.m3.nch;
            // the staging overlaps the weight buffer of the larger motif-round chunks; MLP.3's are small and R's MMAs have completed
            tc_mma_round<TS>(x, H, kcols, colM1, c != 0, last ? noff : L.m3.w + (int64_t)(c + 1) * chunk_floats(L.m3), last ? nbytes : bytes_m3,
                             [&]() {
                                 if (c == 0 && has_next) { request_nodes(pcur, 0); request_edges(pcur, 0); }
                                 if (SHARE && new_tile && c == 0) eq = nxt_live ? group_eq(nxt_g, pcur) : 1;
                             });
        }
        // ---- MLP.5 + sigmoid (:199-200)
        float z5 = 0.f;
        for (int c0 = prt * (H / kParts); c0 < (prt + 1) * (H / kParts); c0 += 16) {
            float z[16];
            tc::tmem_ld16(tmem + lane_base + colM1 + c0, z);
#pragma unroll
            for (int i = 0; i < 16; ++i) z5 = fmaf(fmaxf(z[i] + cstM[L.m_m3b + c0 + i], 0.f), cstM[L.m_w5 + c0 + i], z5);
        }
        part[prt][0][row] = z5;         // the score reduction's reads of part[] ended before the Q rounds' barriers
        tc::fence_before_sync();
        __syncthreads();                 // also: all TMEM reads of this (sub-)tile done before the next MMAs overwrite it
        tc::fence_after_sync();
        if (live && prt == 0) {
            float zz = cstM[L.m_b5];
#pragma unroll
            for (int q = 0; q < kParts; ++q) zz += part[q][0][row];
            const float sc = 1.f / (1.f + expf(-zz));
            a.scores[gm] = sc;
            for (int p = 0; p < a.n_peer; ++p) a.peer[p][gm] = sc;       // warp stores over NVLink; visible to the peers at kernel end
        }
        // the next write to part[] comes after the barriers of the next rounds
    };

    float rp = 0.f;
    for (int64_t tile = blockIdx.x; tile < n_tiles;) {
        const int64_t gr = tile * 128 + row;        // tile row: a motif, or (SHARE) a first-hop slot with its sh consecutive motifs
        const bool live = gr < n_rows;
        // the CTA's next tile (dynamic: SMs that run a single CTA, or faster ones, take more tiles); read after the event passes' barriers
        if (t == 0) s_next_tile = (long long)gridDim.x + (long long)atomicAdd(a.tile_counter, 1ull);
        bool shared = false;
        int64_t next_tile = 0;
#pragma unroll 1
        for (int j = 0; j < sh; ++j) {
            const int64_t gm = live ? gr * sh + j : 0;
            // passes of this (sub-)tile: positions 0, 1, 2, then the [S; P] rounds -- or (SHARE) position 2 and the [S; P] rounds, once per slot
            // where the vote allows, then positions 0 and 1
#pragma unroll 1
            for (int k = (SHARE && j > 0 && shared) ? 1 : 0; k < 3; ++k) {
                const int pos = SHARE ? (k == 0 ? 2 : k - 1) : k;
                if (k < 2) pnext = load_idx(gm, live, SHARE ? k : k + 1);
                int64_t aoff; int abytes;
                if (SPILL ? k == 0 : k == 2) { aoff = L.sp.w; abytes = bytes_sp; }
                else if (SPILL && k == 2) { aoff = L.q.w; abytes = (dq ? 2 : 1) * bytes_q; }
                else first_w(SHARE ? k : k + 1, aoff, abytes);
                // SHARE, k == 0: position 0's rows are requested behind the pass's last round where the staging lies clear of the [S; P] weights
                const int vote = event_pass(pos, pcur, live, SPILL ? (k == 1 || (k == 0 && a.sp_stage_safe)) : k < 2, aoff, abytes, (SHARE && j == 0 && k == 0) ? eq : 1);
                if (SHARE && j == 0 && k == 0) shared = vote != 0;
                if (SPILL && k == 0) {
                    first_w(0, aoff, abytes);
                    rp = sp_rounds(aoff, abytes);
                    pcur = pnext;
                    if (!a.sp_stage_safe) { request_nodes(pcur, 0); request_edges(pcur, 0); }       // otherwise nothing was in flight during the [S; P] rounds
                    spill_uy();
                } else if (k < 2) pcur = pnext;
            }
            if (!SPILL) rp = sp_rounds(L.q.w, (dq ? 2 : 1) * bytes_q);
            bool has_next, nxt_live; int64_t nxt_g; int nxt_pos;
            const bool new_tile = !(SHARE && j + 1 < sh);
            if (!new_tile) { has_next = true; nxt_live = live; nxt_g = gr * sh + j + 1; nxt_pos = shared ? 0 : 2; }
            else {
                next_tile = s_next_tile;
                has_next = next_tile < n_tiles;
                const int64_t rn = next_tile * 128 + row;
                nxt_live = has_next && rn < n_rows; nxt_g = rn * sh; nxt_pos = kFirstPos;
            }
            motif_rounds(gm, live, rp, has_next, nxt_g, nxt_live, nxt_pos, new_tile);
        }
        tile = next_tile;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, a.tmem_cols);
#undef H2
#undef nG
#undef nchS
#undef colE
#undef colY
#undef colM0
#undef colM1
#undef colA3
#undef nsl
#undef dq
#undef de
#undef m3one
#undef cpr
#undef ed_vec
#undef d_vec
#undef jofs
#undef bytes_e
#undef bytes_g
#undef bytes_sp
#undef bytes_q
#undef bytes_r
#undef bytes_m3
#undef n_rows
#undef n_tiles
}

// Motifs whose h rows the workspace holds: two resident CTAs per SM x 128 motifs (tm_encoder_workspace_floats)
int64_t tc_slab_motifs(int device) {
    static int64_t v[64] = {0};
    if (device < 0) cudaGetDevice(&device);
    const int slot = device >= 0 && device < 64 ? device : 0;
    if (!v[slot]) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
        v[slot] = (int64_t)sms * 2 * 128;
    }
    return v[slot];
}

// optional kernel timing (bench.py roofline): CUDA events around the launch
static bool g_prof = false;
static std::vector<cudaEvent_t> g_prof_ev;      // pairs: before, after
static size_t g_prof_used = 0;

// 2-D tensor map over a row-major fp32 feature table [rows x dim] with a box of 32 columns x 1 row (tile::gather4 fetches four such
// rows), 128-byte swizzle, zero fill outside the table
bool make_gather_map(CUtensorMap *map, const float *table, int64_t rows, int dim, int swizzle128) {
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return false;
        encode = (EncodeFn)fn;
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows}, gstride[1] = {(cuuint64_t)dim * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kKC, 1}, estr[2] = {1, 1};
    return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)table, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---------------------------------------------------------------------------------------------
// Edge projection: P[e][n] = sum_j lin_event.weight[n][j] * edge_feat[e][j], j < Ed -- the part of event_conv.lin_event (explainer.py:93)
// that depends on the edge id alone.  The reference recomputes it for every walk event; here it is a table, rebuilt whenever the
// explainer's weights change (one pass over the edge table: n_edge_rows * Ed * D multiply-adds in fp32 FMA, then the scorer gathers P's
// rows where it used to gather the raw feature rows).  Tile of 64 edges per block: features and results staged in shared memory.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
edge_project_kernel(const float *__restrict__ We, const float *__restrict__ efeat, int64_t rows, int Ed, int D, float *__restrict__ P) {
    extern __shared__ float xs[];                    // X [64][Ed + 1] | Y [64][D + 1]
    float *ys = xs + 64 * (Ed + 1);
    const int64_t r0 = (int64_t)blockIdx.x * 64;
    for (int i = threadIdx.x; i < 64 * Ed; i += 256) {
        const int r = i / Ed, j = i - r * Ed;
        xs[r * (Ed + 1) + j] = r0 + r < rows ? __ldg(efeat + (r0 + r) * Ed + j) : 0.f;
    }
    __syncthreads();
    const int r = threadIdx.x & 63;
    const float *x = xs + r * (Ed + 1);
    for (int n = threadIdx.x >> 6; n < D; n += 4) {  // the lanes of a warp share n (broadcast weight loads) and own consecutive rows
        const float *w = We + (int64_t)n * Ed;
        float acc = 0.f;
        for (int j = 0; j < Ed; ++j) acc = fmaf(__ldg(w + j), x[j], acc);
        ys[r * (D + 1) + n] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * D; i += 256) {
        const int rr = i / D, n = i - rr * D;
        if (r0 + rr < rows) P[(r0 + rr) * D + n] = ys[rr * (D + 1) + n];
    }
}

int tc_project_edges(const tm_encoder_desc &d, const float *d_blob_tc, const float *edge_feat, int64_t n_edge_rows, float *P, cudaStream_t st) {
    const TcLayout L = make_tc_layout(d);
    const size_t smem = sizeof(float) * 64 * ((size_t)L.Ed + 1 + L.D + 1);
    if (smem > 200 * 1024) { set_error("tm_encoder_project_edges: edge_dim + node_dim too large for one tile in shared memory"); return TM_ERR_UNSUPPORTED; }
    static bool attr_set[64] = {false};
    int dev = 0;
    TM_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !attr_set[dev]) {
        TM_CUDA(cudaFuncSetAttribute(edge_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set[dev] = true;
    }
    if (n_edge_rows <= 0) return TM_OK;
    edge_project_kernel<<<(unsigned)((n_edge_rows + 63) / 64), 256, smem, st>>>(d_blob_tc + L.we, edge_feat, n_edge_rows, L.Ed, L.D, P);
    TM_LAUNCH_CHECK();
    return TM_OK;
}

// std_ = per-batch std (already computed); F = scratch for the h slabs of the resident CTAs
int tc_encode_score(const tm_encoder_desc &d, const float *d_blob_tc, int64_t B, int64_t W, int64_t group, const int32_t *nodes,
                    const int32_t *eidx, const float *t, const uint8_t *cat, const float *cut, const float *eid, const float *node_feat,
                    int64_t n_node_rows, const float *edge_feat, int64_t n_edge_rows, const float *std_, float *F, float *scores,
                    float *y_out, float *const *peer_scores, int n_peers, int device, cudaStream_t st) {
    const TcLayout L = make_tc_layout(d);
    if (n_peers < 0 || n_peers > kMaxPeers || (n_peers > 0 && !peer_scores)) { set_error("tc_encode_score: at most %d peer outputs", kMaxPeers); return TM_ERR_ARG; }
    if (L.H != 64 && L.H != 32) { set_error("tc_encode_score: hid_dim must be 64 or 32 (the defaults of temp_exp_main.py / enhance_main.py)"); return TM_ERR_UNSUPPORTED; }
    if (device < 0 || device >= 64) { set_error("tc_encode_score: device index out of range"); return TM_ERR_UNSUPPORTED; }
    const int H = L.H;
    const bool alias_e = L.g0.nch == 1 && L.D16 <= H;
    const char *ts_env = getenv("TEMPME_TC_A");                  // "smem": A operand in shared memory (SS); default: in TMEM (TS), the last 64 columns
    const bool ts = !(ts_env && strcmp(ts_env, "smem") == 0);
    uint32_t cols = 32;
    // event passes with two A buffers (pairs of lin_event chunks, both MLP.0 orientations per round): TS mode, at most the first
    // lin_event chunk holds edge columns (one staged edge chunk per round)
    const bool dual = ts && !getenv("TEMPME_TC_NO_DUAL");
    // drain mode (node_dim > 32, where E cannot alias Zt): lin_event's output leaves TMEM through an L2 scratch before the MLP.0 rounds, so E
    // [0, D16) and Zs | Zt [0, 2H) share columns and the tile needs 256 columns instead of 512 -- two tiles per SM instead of one
    const bool drain = ts && dual && !alias_e && L.D16 + 2 * kKC <= 256 && 2 * H + 4 * kKC <= 256 && !getenv("TEMPME_TC_NO_DRAIN");
    const bool proj = d.edge_projected != 0;                 // d_edge_feat is the projected table [n_edge_rows][D] (tc_project_edges)
    const TcLin &EV = proj ? L.evtT : L.evt;
    const bool dual_e = dual && (proj || L.nch_edge <= 1) && !drain;   // pairs of lin_event chunks: one staged edge chunk per round at most; drain mode has one A buffer beside E
    // motif rounds: U | Y + one A buffer; event passes: Zs | Zt (| E) + one or two A buffers (the A buffers are the top columns)
    const int ev_cols = drain ? std::max(2 * H + 4 * kKC, L.D16 + 2 * kKC) : (alias_e ? 2 * H : 2 * H + L.D16) + (ts ? (dual ? 4 : 2) * kKC : 0);
    while ((int)cols < std::max(3 * H + (ts ? 2 * kKC : 0), ev_cols)) cols <<= 1;
    if (cols > 512) { set_error("tc_encode_score: node_dim too large for the TMEM layout"); return TM_ERR_UNSUPPORTED; }
    int64_t bb = 0;
    for (const TcLin *l : {&EV, &L.g0, &L.sp, &L.q, &L.r, &L.m3}) bb = std::max(bb, chunk_floats(*l) * 4);
    if (dual_e) bb = std::max(bb, (int64_t)std::min(2, EV.nch) * chunk_floats(EV) * 4);              // lin_event chunks arrive in pairs
    if (dual) bb = std::max(bb, 2 * std::max(chunk_floats(L.q), chunk_floats(L.r)) * 4);             // so do Q's and R's
    // MLP.3 in one round: all of its (two or three) K chunks in the weight buffer; the third A buffer sits between M0 and the A buffers
    const bool m3one = dual && !getenv("TEMPME_TC_NO_M3ONE") && L.m3.nch <= 3 && (L.m3.nch < 3 || (L.m3.K8 - 2 * kKC <= 16 && H == 64 && H + L.M16 + 3 * kKC / 2 <= 3 * H));
    if (m3one) bb = std::max(bb, (int64_t)L.m3.nch * chunk_floats(L.m3) * 4);
    // node-feature rows are gathered by bulk TMA into a staging area that may overlap the tail of the weight buffer: rows
    // are in flight only while lin_event / MLP.0 / MLP.3 chunks are being loaded, so it starts behind the largest of those
    CUtensorMap tm_node, tm_edge;
    memset(&tm_node, 0, sizeof tm_node); memset(&tm_edge, 0, sizeof tm_edge);
    const bool stage_nodes = L.D % 4 == 0 && ((uintptr_t)node_feat & 15) == 0 && !getenv("TEMPME_TC_NO_STAGING") && make_gather_map(&tm_node, node_feat, n_node_rows, L.D, 1);
    const int edge_cols = proj ? L.D : L.Ed;                 // row length of the table behind d_edge_feat
    const bool stage_edges = edge_cols % 4 == 0 && ((uintptr_t)edge_feat & 15) == 0 && !getenv("TEMPME_TC_NO_STAGING") && !getenv("TEMPME_TC_NO_EDGE_STAGING") &&
                             make_gather_map(&tm_edge, edge_feat, n_edge_rows, edge_cols, 1);
    const int64_t stage_rel = (std::max(std::max((dual_e ? std::min(2, EV.nch) : 1) * chunk_floats(EV), chunk_floats(L.g0)), (m3one ? L.m3.nch : 1) * chunk_floats(L.m3)) * 4 + 1023) & ~(int64_t)1023;
    const int64_t stage_edge_rel = stage_rel + (stage_nodes ? (int64_t)2 * kStageTable * 4 : 0);
    bb = std::max(bb, stage_edge_rel + (stage_edges ? (int64_t)kStageTable * 4 : 0));
    const size_t a_bytes = ts ? 0 : (size_t)2 * kATile;
    const size_t need = a_bytes + (size_t)bb + (size_t)(L.n_cstE + L.n_cstM) * 4 + 152 * 8;
    if (need > 220 * 1024) { set_error("tc_encode_score: feature dims too large for one weight chunk in shared memory"); return TM_ERR_UNSUPPORTED; }
    using ScoreK = void (*)(const TcLayout, const float *, const TcArgs, const CUtensorMap, const CUtensorMap);
    constexpr int kKernels = 7;
    static const ScoreK kern[kKernels] = {score_tc_kernel<8, false, false, false>, score_tc_kernel<16, false, false, false>, score_tc_kernel<8, true, false, false>,
                                          score_tc_kernel<16, true, false, false>, score_tc_kernel<16, true, true, false>,
                                          score_tc_kernel<16, true, false, true>, score_tc_kernel<16, true, true, true>};      // [5], [6]: walk groups (SHARE)
    static bool attr_set[64] = {false};
    if (!attr_set[device]) {
        for (ScoreK k : kern) {
            TM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 8192));
            TM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        }
        attr_set[device] = true;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const char *cw_env = getenv("TEMPME_TC_CW");                 // columns per thread per K chunk: 16 (256 threads, default) or 8 (512 threads)
    const int cw = cw_env && atoi(cw_env) == 8 && H == 64 && !drain ? 8 : 16;
    // walk groups: d.walk_fanout consecutive walks share their first-hop event (verified per tile by the kernel); needs the default
    // operand path (A in TMEM, 256 threads) and whole groups
    const int share = (d.walk_fanout >= 2 && d.walk_fanout <= 64 && ts && dual && cw == 16 && W % d.walk_fanout == 0 && !getenv("TEMPME_TC_NO_SHARE")) ? d.walk_fanout : 1;
    const int kv = share > 1 ? (drain ? 6 : 5) : drain ? 4 : (cw == 16) + 2 * ts;         // drain mode has its own instantiations (CW = 16, A operand in TMEM)
    static size_t static_smem[kKernels] = {0};
    if (!static_smem[0])
        for (int v = 0; v < kKernels; ++v) { cudaFuncAttributes fa; TM_CUDA(cudaFuncGetAttributes(&fa, kern[v])); static_smem[v] = fa.sharedSizeBytes; }
    // resident CTAs per SM: TMEM columns and shared memory (registers: __launch_bounds__(threads, 2))
    const int ctas = std::max(1, std::min<int>(2, std::min<int>(512 / cols, (int)((228 * 1024) / (need + 1024 + static_smem[kv])))));
    // dynamic shared memory padded so that no more than `ctas` CTAs fit an SM (TMEM columns are not part of the occupancy
    // calculation: a CTA beyond 512 / cols would spin in tcgen05.alloc while holding its other resources)
    const size_t smem = std::max(need, std::min((size_t)228 * 1024 / (ctas + 1), (size_t)227 * 1024 - 8192));

    TcArgs a;
    a.n_motifs = B * W; a.W = W; a.group = group; a.nodes = nodes; a.eidx = eidx; a.t = t; a.cat = cat; a.cut = cut;
    a.eid = eid; a.node_feat = node_feat; a.edge_feat = edge_feat; a.std_ = std_; a.n_node_rows = n_node_rows; a.n_edge_rows = n_edge_rows;
    a.n_peer = n_peers;
    for (int p = 0; p < kMaxPeers; ++p) a.peer[p] = p < n_peers ? peer_scores[p] : nullptr;
    a.F = F; a.scores = scores; a.y_out = y_out; a.tmem_cols = cols; a.b_bytes = (int)bb; a.dbg = nullptr;
    a.stage_off = stage_nodes ? (int)(a_bytes + stage_rel) : 0;
    a.stage_edge_off = stage_edges ? (int)(a_bytes + stage_edge_rel) : 0;
    a.dual = (dual ? 1 : 0) | (dual_e ? 2 : 0) | (m3one ? 4 : 0);
    a.proj = proj ? 1 : 0;
    a.eid_u8 = d.edge_identity_u8 != 0 ? 1 : 0;
    a.nodes_vec = ((uintptr_t)nodes & 7) == 0 ? 1 : 0;
    if (a.eid_u8 && ((uintptr_t)eid & 3) != 0) { set_error("tc_encode_score: byte edge-identity counts must be 4-byte aligned"); return TM_ERR_ARG; }
    a.discard = getenv("TEMPME_TC_DISCARD") ? 1 : 0;          // A/B knob, off: the discards removed the scratch write-back but cost 2.6 % of kernel time (profiles/README.md r02b)
    static long long *dbg_buf = nullptr;
#ifdef TM_TC_TIMING
    const char *tim_env = getenv("TEMPME_TC_TIMING");          // diagnostic: per-round clock stamps of CTA 0
#else
    const char *tim_env = nullptr;                             // needs a -DTM_TC_TIMING build (TEMPME_BUILD_TIMING=1 python -m tempme_b200.build)
#endif
    if (tim_env) {
        if (!dbg_buf) TM_CUDA(cudaMalloc(&dbg_buf, 128 * 5 * sizeof(long long)));
        TM_CUDA(cudaMemsetAsync(dbg_buf, 0, 128 * 5 * sizeof(long long), st));
        a.dbg = dbg_buf;
    }
    a.share = share;
    {
        TcDer &k = a.k;
        const bool dq_ = dual, de_ = dual_e;
        k.H2 = 2 * H; k.nG = L.g0.nch; k.nchS = L.sp.nch;
        k.colE = drain ? 0 : (L.g0.nch == 1 && L.D16 <= H) ? H : 2 * H;
        k.colY = 2 * H; k.colM0 = dq_ ? std::max(H, 2 * kKC) : 0; k.colM1 = m3one ? k.colM0 : dq_ ? 0 : 2 * H; k.colA3 = k.colM0 + L.M16;
        k.nsl = 2 * H / kKC; k.dq = dq_; k.de = de_; k.m3one = m3one; k.cpr = de_ ? 2 : 1;
        k.ed_vec = (L.Ed & 3) == 0; k.d_vec = (L.D & 3) == 0; k.jofs = proj ? L.Ed : 0;
        k.ev_K8 = EV.K8; k.ev_nch = EV.nch; k.ev_w = EV.w; k.ev_chunk = chunk_floats(EV);
        k.ne01 = proj ? EV.nch : L.evt.nch; k.ne2 = proj ? 0 : L.nch_edge;
        k.bytes_e = (int)chunk_floats(EV) * 4; k.bytes_g = (int)chunk_floats(L.g0) * 4; k.bytes_sp = (int)chunk_floats(L.sp) * 4;
        k.bytes_q = (int)chunk_floats(L.q) * 4; k.bytes_r = (int)chunk_floats(L.r) * 4; k.bytes_m3 = (int)chunk_floats(L.m3) * 4;
        auto fw = [&](int ne, int64_t &off, int &bytes) { if (ne > 0) { off = EV.w; bytes = std::min(k.cpr, ne) * k.bytes_e; } else { off = L.g0.w; bytes = k.bytes_g; } };
        fw(k.ne01, k.fw01_off, k.fw01_bytes); fw(k.ne2, k.fw2_off, k.fw2_bytes);
        k.n_rows = a.n_motifs / share; k.n_tiles = (k.n_rows + 127) / 128;
    }
    a.sp_stage_safe = stage_rel >= chunk_floats(L.sp) * 4 ? 1 : 0;
    const int64_t tiles = (a.n_motifs / share + 127) / 128, cap = (int64_t)sms * ctas;
    const unsigned grid = (unsigned)std::min(tiles, cap);
    a.tile_counter = reinterpret_cast<unsigned long long *>(F + (int64_t)grid * 3 * (2 * H / kKC) * kSlabFloats);          // behind the h scratch (the workspace holds twice as much)
    a.Ys = F + (int64_t)grid * 3 * (2 * H / kKC) * kSlabFloats + 64;                                                      // SHARE: H / 32 slabs per CTA, a sixth of the h scratch
    a.Es = drain ? a.Ys + (int64_t)grid * (H / kKC) * kSlabFloats : nullptr;                                              // drain mode: nG slabs per CTA behind both
    const size_t scratch_bytes = (size_t)((drain ? a.Es + (int64_t)grid * L.g0.nch * kSlabFloats : a.Ys + (int64_t)grid * (H / kKC) * kSlabFloats) - F) * sizeof(float);
    TM_CUDA(cudaMemsetAsync(a.tile_counter, 0, sizeof(unsigned long long), st));
    if (getenv("TEMPME_TC_DEBUG")) fprintf(stderr, "[tc] score kernel: %u CTAs (%d per SM), smem %zu B (needs %zu), %u TMEM columns, %lld tiles%s%s\n", grid, ctas, smem, need, cols, (long long)tiles, share > 1 ? " of walk groups" : "", drain ? (proj ? ", E drained through L2, projected edge table" : ", E drained through L2") : (proj ? ", projected edge table" : ""));
    cudaEvent_t *pe = nullptr;
    if (g_prof && g_prof_used + 2 <= g_prof_ev.size()) {        // pool is created by tm_encoder_profile(1); when exhausted, stop recording
        pe = &g_prof_ev[g_prof_used]; g_prof_used += 2;
        cudaEventRecord(pe[0], st);
    }
    // The CTAs' scratch (h slabs, U / Y, drained E) is written and read back within microseconds, but streams of feature rows and walk tensors
    // pass through L2 in between: TEMPME_TC_L2_PERSIST_MB=<n> sets n MB of L2 aside for persisting lines and marks the scratch as such
    // (access policy window on this launch), so that its reads hit and its dirty lines are not written back to HBM.
    static int persist_mb[64];
    static bool persist_init[64] = {false};
    if (!persist_init[device]) {
        const char *pm = getenv("TEMPME_TC_L2_PERSIST_MB");
        int want = pm ? atoi(pm) : 0, max_persist = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device);
        want = std::min(want, max_persist >> 20);
        if (want > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)want << 20) != cudaSuccess) { cudaGetLastError(); want = 0; }
        persist_mb[device] = want;
        persist_init[device] = true;
        if (getenv("TEMPME_TC_DEBUG")) fprintf(stderr, "[tc] L2 set-aside %d MB (device maximum %d MB)\n", want, max_persist >> 20);
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128 * (kKC / cw)); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (persist_mb[device] > 0) {
        int max_win = 0;
        cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, device);
        const size_t win = std::min(scratch_bytes, (size_t)max_win);
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = F;
        attr[0].val.accessPolicyWindow.num_bytes = win;
        attr[0].val.accessPolicyWindow.hitRatio = std::min(1.0f, (float)((double)((size_t)persist_mb[device] << 20) / (double)std::max<size_t>(win, 1)));
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.attrs = attr; cfg.numAttrs = 1;
    }
    TM_CUDA(cudaLaunchKernelEx(&cfg, kern[kv], L, d_blob_tc, a, tm_node, tm_edge));
    TM_LAUNCH_CHECK();
    if (pe) cudaEventRecord(pe[1], st);
    if (tim_env) {
        TM_CUDA(cudaStreamSynchronize(st));
        std::vector<long long> h(128 * 5);
        TM_CUDA(cudaMemcpy(h.data(), dbg_buf, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[tc timing] CTA 0: round: fill+gap | fence+sync | weights wait | MMA issue | MMA done wait || round total   (cycles)\n");
        for (int r = 1; r < 128 && h[r * 5 + 4]; ++r)
            fprintf(stderr, "  %3d: %6lld %6lld %6lld %6lld %6lld || %6lld\n", r, h[r * 5] - h[(r - 1) * 5 + 4], h[r * 5 + 1] - h[r * 5], h[r * 5 + 2] - h[r * 5 + 1],
                    h[r * 5 + 3] - h[r * 5 + 2], h[r * 5 + 4] - h[r * 5 + 3], h[r * 5 + 4] - h[(r - 1) * 5 + 4]);
    }
    return TM_OK;
}

// =================================================================================================
// Dependency gate of retrieve_edge_imp_node (reference models/explainer.py:367-386): per walk event
//   walk_imp = score * (0.5 + 0.5 sigmoid(edge_dependency_gcn([edge features | TimeEncode(raw t)])))
// edge_dependency_gcn = Linear(Ed + D, H), ReLU, Linear(H, H/2), ReLU, Linear(H/2, 1) (:143-151, Dropout = identity in eval).
// Tile = 128 events, 256 threads, same round machinery as the scorer (TS mode).  TMEM: G1 [0,H)  G2 [H, H + H/2)  A [192,256).
// =================================================================================================
struct GateLayout {
    int D, Ed, H, H2, D16;                 // H2 = H / 2
    TcLin l1, l2;
    int b1, b2, w3, b3, freq, phase, n_cst;
    int64_t cst, total;
};

GateLayout make_gate_layout(const tm_gate_desc &d) {
    GateLayout G;
    memset(&G, 0, sizeof G);
    G.D = d.time_dim; G.Ed = d.edge_dim; G.H = d.hid_dim; G.H2 = d.hid_dim / 2; G.D16 = r16(G.D);
    int64_t o = 0;
    auto lin = [&](int K, int N) {
        TcLin l; l.K8 = r8(K); l.N16 = r16(N); l.nch = (l.K8 + kKC - 1) / kKC;
        l.w = o; o += (int64_t)l.nch * chunk_floats(l);
        return l;
    };
    G.l1 = lin(G.Ed + G.D, G.H); G.l2 = lin(G.H, G.H2);
    int c = 0;
    G.b1 = c; c += r16(G.H); G.b2 = c; c += r16(G.H2); G.w3 = c; c += r16(G.H2); G.b3 = c; c += 16;
    G.freq = c; c += G.D16; G.phase = c; c += G.D16; G.n_cst = c;
    G.cst = o; o += c; G.total = o;
    return G;
}

int64_t tc_gate_blob_floats(const tm_gate_desc &d) { return (make_gate_layout(d).total + 31) & ~(int64_t)31; }

int tc_gate_pack(const tm_gate_desc &d, const tm_gate_params &p, float *blob) {
    const GateLayout G = make_gate_layout(d);
    memset(blob, 0, sizeof(float) * ((G.total + 31) & ~(int64_t)31));
    const Mat W1 = dbl(p.w0, (size_t)G.H * (G.Ed + G.D)), W2 = dbl(p.w3, (size_t)G.H2 * G.H);
    pack_tc_lin(G.l1, G.Ed + G.D, G.H, W1.data(), blob);
    pack_tc_lin(G.l2, G.H, G.H2, W2.data(), blob);
    float *c = blob + G.cst;
    for (int i = 0; i < G.H; ++i) c[G.b1 + i] = p.b0[i];
    for (int i = 0; i < G.H2; ++i) { c[G.b2 + i] = p.b3[i]; c[G.w3 + i] = p.w6[i]; }
    c[G.b3] = p.b6[0];
    for (int i = 0; i < G.D; ++i) { c[G.freq + i] = p.basis_freq[i]; c[G.phase + i] = p.phase[i]; }
    return TM_OK;
}

struct GateArgs {
    int64_t n_events;                        // B * W * 3 walk events, flat
    const int32_t *eidx;
    const float *t, *scores, *edge_feat;
    int64_t n_edge_rows;
    float *out;                              // walk_imp [n_events]
    uint32_t tmem_cols;
    int b_bytes;
};

__global__ void __launch_bounds__(256, 2)
gate_tc_kernel(const GateLayout G, const float *__restrict__ blob, const GateArgs a) {
    constexpr int CW = 16;
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float part[2][128];
    const int t = threadIdx.x, warp = t >> 5, row = t & 127, prt = t >> 7, kb = CW * prt;
    TcCtx x;
    x.a = smem; x.a_s = tc::smem_u32(smem); x.b_s = x.a_s; x.bars = bars; x.mma_phase = 0; x.b_phase = 0; x.blob = blob;
    x.dbg = nullptr; x.dbg_i = 0;
    float *cst = reinterpret_cast<float *>(smem + a.b_bytes);
    uint2 *ctab = reinterpret_cast<uint2 *>(cst + G.n_cst);
    for (int i = t; i < G.n_cst; i += 256) cst[i] = __ldg(blob + G.cst + i);
    cos_table_to_smem(ctab);
    if (t == 0) { tc::mbar_init(bars, 1); tc::mbar_init(bars + 1, 1); }
    if (warp == 0) tc::tmem_alloc(&tmem_slot, a.tmem_cols);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot, lane_base = (uint32_t)((warp & 3) * 32) << 16;
    x.tmem = tmem; x.a_col = a.tmem_cols - 2 * kKC; x.a_col2 = x.a_col - 2 * kKC;          // TMEM: G1 [0,H)  G2 [H, H + H/2)  A2 [128,192)  A [192,256)
    AFill<CW, true> af;
    const int D = G.D, Ed = G.Ed, H = G.H, colG1 = 0, colG2 = H;
    const int64_t n_tiles = (a.n_events + 127) / 128;
    const int bytes1 = (int)chunk_floats(G.l1) * 4, bytes2 = (int)chunk_floats(G.l2) * 4;
    if (t == 0 && blockIdx.x < n_tiles) tc_request_b(x, G.l1.w, min(2, G.l1.nch) * bytes1);
    const bool ed_vec = (Ed & 3) == 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t r = tile * 128 + row;
        const bool live = r < a.n_events, more = tile + gridDim.x < n_tiles;
        const int32_t e = live ? a.eidx[r] : -1;
        const float tt = live ? a.t[r] : 0.f;                                   // the raw timestamp (explainer.py:371-372)
        const bool e_ok = e >= 0 && e < a.n_edge_rows;
        const float *ef = a.edge_feat + (int64_t)max(e, 0) * Ed;
        auto xval = [&](int j) -> float {                                       // [edge features | TimeEncode] column j (:375)
            if (j < Ed) return e_ok ? __ldg(ef + j) : 0.f;
            const int k = j - Ed;
            return (k < D && live) ? cos_accurate(__fadd_rn(__fmul_rn(tt, cst[G.freq + k]), cst[G.phase + k]), ctab) : 0.f;
        };
        // ---- Linear(Ed + D, H) -> G1, two K chunks per round
        auto fill1 = [&](int c, bool second) {
            const int kcols = min(kKC, G.l1.K8 - c * kKC), j0 = c * kKC + kb;
            if (kb >= kcols) return;
            if (ed_vec && j0 + CW <= Ed) {
#pragma unroll
                for (int g = 0; g < CW / 4; ++g) af.put4(x, row, kb, 4 * g, e_ok ? ldg4(ef + j0 + 4 * g) : make_float4(0.f, 0.f, 0.f, 0.f));
            } else if (j0 >= Ed && ((j0 - Ed) & 3) == 0 && j0 + CW <= Ed + G.D16) {
                float w[CW];
#pragma unroll
                for (int g = 0; g < CW / 4; ++g) {
                    const float4 fq = lds4(cst + G.freq + (j0 - Ed) + 4 * g), ph = lds4(cst + G.phase + (j0 - Ed) + 4 * g);
                    w[4 * g] = cos_accurate(__fadd_rn(__fmul_rn(tt, fq.x), ph.x), ctab); w[4 * g + 1] = cos_accurate(__fadd_rn(__fmul_rn(tt, fq.y), ph.y), ctab);
                    w[4 * g + 2] = cos_accurate(__fadd_rn(__fmul_rn(tt, fq.z), ph.z), ctab); w[4 * g + 3] = cos_accurate(__fadd_rn(__fmul_rn(tt, fq.w), ph.w), ctab);
                }
#pragma unroll
                for (int g = 0; g < CW / 4; ++g) {
                    const int k = j0 - Ed + 4 * g;
                    af.put4(x, row, kb, 4 * g, make_float4((k < D && live) ? w[4 * g] : 0.f, (k + 1 < D && live) ? w[4 * g + 1] : 0.f,
                                                           (k + 2 < D && live) ? w[4 * g + 2] : 0.f, (k + 3 < D && live) ? w[4 * g + 3] : 0.f));
                }
            } else {
#pragma unroll
                for (int g = 0; g < CW / 4; ++g) { const int j = j0 + 4 * g; af.put4(x, row, kb, 4 * g, make_float4(xval(j), xval(j + 1), xval(j + 2), xval(j + 3))); }
            }
            af.commit(x, lane_base, kb, second);
        };
        for (int c = 0; c < G.l1.nch; c += 2) {
            const int cnt = min(2, G.l1.nch - c), left = G.l1.nch - c - cnt;
            fill1(c, false);
            if (cnt == 2) fill1(c + 1, true);
            const int kc0 = min(kKC, G.l1.K8 - c * kKC), kc1 = cnt == 2 ? min(kKC, G.l1.K8 - (c + 1) * kKC) : 0;
            tc_mma_round<true>(x, r16(H), kc0, colG1, c != 0, left > 0 ? G.l1.w + (int64_t)(c + 2) * chunk_floats(G.l1) : G.l2.w,
                               left > 0 ? min(2, left) * bytes1 : min(2, G.l2.nch) * bytes2, NoMid(), Dual{cnt == 2 ? kDualK : kSingle, kc1, 0, 0, 0});
        }
        // ---- ReLU, Linear(H, H/2) -> G2, two K chunks per round
        auto fill2 = [&](int c, bool second) {
            const int kcols = min(kKC, G.l2.K8 - c * kKC);
            if (kb >= kcols) return;
            float z[CW];
            tc::tmem_ld16(tmem + lane_base + colG1 + c * kKC + kb, z);
#pragma unroll
            for (int k = 0; k < CW; k += 4) {
                const float4 bb = lds4(cst + G.b1 + c * kKC + kb + k);
                af.put4(x, row, kb, k, make_float4(fmaxf(z[k] + bb.x, 0.f), fmaxf(z[k + 1] + bb.y, 0.f), fmaxf(z[k + 2] + bb.z, 0.f), fmaxf(z[k + 3] + bb.w, 0.f)));
            }
            af.commit(x, lane_base, kb, second);
        };
        for (int c = 0; c < G.l2.nch; c += 2) {
            const int cnt = min(2, G.l2.nch - c), left = G.l2.nch - c - cnt;
            fill2(c, false);
            if (cnt == 2) fill2(c + 1, true);
            const int kc0 = min(kKC, G.l2.K8 - c * kKC), kc1 = cnt == 2 ? min(kKC, G.l2.K8 - (c + 1) * kKC) : 0;
            tc_mma_round<true>(x, r16(G.H2), kc0, colG2, c != 0, left > 0 ? G.l2.w + (int64_t)(c + 2) * chunk_floats(G.l2) : G.l1.w,
                               left > 0 ? min(2, left) * bytes2 : (more ? min(2, G.l1.nch) * bytes1 : 0), NoMid(), Dual{cnt == 2 ? kDualK : kSingle, kc1, 0, 0, 0});
        }
        // ---- ReLU, Linear(H/2, 1), sigmoid gate (:379-386)
        float g_ = 0.f;
        for (int c0 = prt * 16; c0 < r16(G.H2); c0 += 32) {          // part p takes the 16-column groups p, p + 2, ... (hid_dim 32: one group)
            float z[16];
            tc::tmem_ld16(tmem + lane_base + colG2 + c0, z);
#pragma unroll
            for (int i = 0; i < 16; ++i) g_ = fmaf(fmaxf(z[i] + cst[G.b2 + c0 + i], 0.f), cst[G.w3 + c0 + i], g_);
        }
        part[prt][row] = g_;
        tc::fence_before_sync();
        __syncthreads();                 // also: all TMEM reads of this tile done before the next tile's MMAs overwrite it
        tc::fence_after_sync();
        if (live && prt == 0) {
            const float dep = part[0][row] + part[1][row] + cst[G.b3];
            const float gate = 1.f / (1.f + expf(-dep));
            a.out[r] = __fmul_rn(a.scores[r / 3], __fadd_rn(0.5f, __fmul_rn(0.5f, gate)));
        }
        __syncthreads();
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, a.tmem_cols);
}

int tc_gate_launch(const tm_gate_desc &d, const float *d_blob, int64_t n_events, const int32_t *eidx, const float *t, const float *scores,
                   const float *edge_feat, int64_t n_edge_rows, float *out, int device, cudaStream_t st) {
    const GateLayout G = make_gate_layout(d);
    if ((G.H != 64 && G.H != 32) || G.Ed < 1 || G.D < 1 || G.D > 256) { set_error("tm_edge_importance: gate needs hid_dim 64 or 32 and time_dim in [1,256]"); return TM_ERR_UNSUPPORTED; }
    const int64_t bb = std::max(std::min(2, G.l1.nch) * chunk_floats(G.l1), std::min(2, G.l2.nch) * chunk_floats(G.l2)) * 4;       // chunks arrive in pairs
    const size_t need = (size_t)bb + (size_t)G.n_cst * 4 + 152 * 8;
    static bool attr_set[64] = {false};
    if (device >= 0 && device < 64 && !attr_set[device]) {
        TM_CUDA(cudaFuncSetAttribute(gate_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        TM_CUDA(cudaFuncSetAttribute(gate_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr_set[device] = true;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    GateArgs a;
    a.n_events = n_events; a.eidx = eidx; a.t = t; a.scores = scores; a.edge_feat = edge_feat; a.n_edge_rows = n_edge_rows; a.out = out;
    a.tmem_cols = 256; a.b_bytes = (int)bb;
    const size_t smem = std::max(need, (size_t)228 * 1024 / 3);       // two CTAs per SM (256 TMEM columns each), never three
    const int64_t tiles = (n_events + 127) / 128, cap = (int64_t)sms * 2, per = (tiles + cap - 1) / cap;
    gate_tc_kernel<<<(unsigned)((tiles + per - 1) / per), 256, smem, st>>>(G, d_blob, a);
    TM_LAUNCH_CHECK();
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
