// Fused motif encoder / scorer: TempME.forward in eval mode (reference models/explainer.py:174-201):
// feature gathers (:318-352) + TimeEncode (:45-59) + event_gcn x2 (:79-96, lin_event evaluated once)
// + TemporalAwareAttention (:768-846) / Attention (:12-43) + category one-hot (:308-315) + MLP + sigmoid.
// One CTA owns a tile of T motifs; every activation of the tile stays in shared memory from the
// gathers to the score, weights stream from L2 in a transposed, 32-column padded layout.
// Arithmetic is fp32 throughout (the reference's type); TimeEncode keeps the reference's
// mul-then-add rounding (no FMA contraction) and the accurate cosf because its arguments reach 1e8.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace tmb {

constexpr int kEncThreads = 256;
constexpr int kEncWarps = kEncThreads / 32;

struct Lin { int64_t w, b; int K, Kp, N, Np; };   // Wt [Kp][Np] at w, bias [Np] at b (float offsets in the blob)

struct EncLayout {
    int D, Ed, H, M, ev, evp, Dp, use_temporal, if_cat;
    Lin evt, g0, g2, w1, w2, a0, a3, m0, m3, m5;
    int64_t freq, phase, total;
};

__host__ __device__ static inline int r4(int x) { return (x + 3) & ~3; }
__host__ __device__ static inline int r32(int x) { return (x + 31) & ~31; }

static EncLayout make_layout(const tm_encoder_desc &d) {
    EncLayout L;
    memset(&L, 0, sizeof L);
    L.D = d.node_dim; L.Ed = d.edge_dim; L.H = d.hid_dim; L.use_temporal = d.use_temporal; L.if_cat = d.if_cat;
    L.M = d.if_cat ? d.hid_dim + 12 : d.hid_dim;
    L.ev = L.Ed + 3 + L.D; L.evp = r4(L.ev); L.Dp = r4(L.D);
    int64_t o = 0;
    auto lin = [&](int K, int N) { Lin l; l.K = K; l.Kp = r4(K); l.N = N; l.Np = r32(N); l.w = o; o += (int64_t)l.Kp * l.Np; l.b = o; o += l.Np; return l; };
    L.evt = lin(L.ev, L.D); L.g0 = lin(L.D, L.H); L.g2 = lin(L.H, L.H);
    L.w1 = lin(2 * L.H, 2 * L.H); L.w2 = lin(2 * L.H, 2 * L.H); L.a0 = lin(2 * L.H, L.H); L.a3 = lin(L.H, L.H);
    L.m0 = lin(L.M, L.M); L.m3 = lin(L.M, L.H); L.m5 = lin(L.H, 1);
    L.freq = o; o += r32(L.D); L.phase = o; o += r32(L.D);
    L.total = o;
    return L;
}

// ---- batch-global std of |cut_time - t_k|, k = 0,1 over one reference batch (explainer.py:826-828)
__global__ void __launch_bounds__(256)
time_std_kernel(int64_t B, int64_t W, int64_t group, const float *__restrict__ t, const float *__restrict__ cut, float *__restrict__ o_std) {
    __shared__ double sh[8];
    __shared__ double sh_mean;
    const int64_t b0 = (int64_t)blockIdx.x * group, b1 = min(B, b0 + group);
    const int64_t cnt = (b1 - b0) * W * 2;
    auto block_sum = [&](double v) {
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
        __syncthreads();
        double s = 0;
        for (int i = 0; i < 8; ++i) s += sh[i];
        return s;
    };
    auto value = [&](int64_t i) {  // i over [rows, W, 2]
        const int64_t bw = i >> 1, b = b0 + bw / W;
        return (double)fabsf(__fsub_rn(cut[b], t[(b0 * W + bw) * 3 + (i & 1)]));
    };
    double s = 0;
    for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) s += value(i);
    s = block_sum(s);
    if (threadIdx.x == 0) sh_mean = s / (double)cnt;
    __syncthreads();
    const double mean = sh_mean;
    double q = 0;
    for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) { const double d = value(i) - mean; q += d * d; }
    q = block_sum(q);
    if (threadIdx.x == 0) o_std[blockIdx.x] = (float)sqrt(q / (double)(cnt - 1));  // unbiased; cnt == 1 -> NaN like torch
}

// ---- warp GEMM over shared-memory rows: out[r][:] = A[r][:] * Wt + bias for the rows of this warp
template <int CN, typename RowPtr, typename Epi>
__device__ __forceinline__ void warp_gemm(int rows, const Lin l, const float *__restrict__ blob, RowPtr rowptr, Epi epi, int warp, int lane) {
    constexpr int RM = 4;
    const float *__restrict__ Wt = blob + l.w;
    const float *__restrict__ bias = blob + l.b;
    for (int r0 = warp * RM; r0 < rows; r0 += kEncWarps * RM) {
        float acc[RM][CN];
        const float *ap[RM];
#pragma unroll
        for (int r = 0; r < RM; ++r) {
            ap[r] = rowptr(min(r0 + r, rows - 1));
#pragma unroll
            for (int c = 0; c < CN; ++c) acc[r][c] = 0.f;
        }
        for (int k = 0; k < l.Kp; k += 4) {
            float4 a[RM];
#pragma unroll
            for (int r = 0; r < RM; ++r) a[r] = *reinterpret_cast<const float4 *>(ap[r] + k);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                float w[CN];
#pragma unroll
                for (int c = 0; c < CN; ++c) w[c] = __ldg(Wt + (size_t)(k + kk) * l.Np + lane + 32 * c);
#pragma unroll
                for (int r = 0; r < RM; ++r) {
                    const float av = kk == 0 ? a[r].x : kk == 1 ? a[r].y : kk == 2 ? a[r].z : a[r].w;
#pragma unroll
                    for (int c = 0; c < CN; ++c) acc[r][c] = fmaf(av, w[c], acc[r][c]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RM; ++r)
            if (r0 + r < rows) {
#pragma unroll
                for (int c = 0; c < CN; ++c) {
                    const int col = lane + 32 * c;
                    if (col < l.N) epi(r0 + r, col, acc[r][c] + __ldg(bias + col));
                }
            }
    }
}

struct EncArgs {
    int64_t n_motifs, W, group;
    const int32_t *nodes, *eidx;
    const float *t;
    const uint8_t *cat;
    const float *cut, *eid, *node_feat, *edge_feat, *std_;
    int64_t n_node_rows, n_edge_rows;
    float *scores;
    int T;
};

template <int CN_D, int CN_H, int CN_2H, int CN_M>
__global__ void __launch_bounds__(kEncThreads, 1)
encode_kernel(const EncLayout L, const float *__restrict__ blob, const EncArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int T = a.T, R3 = 3 * T, H = L.H, H2 = 2 * L.H, Dp = L.Dp, evp = L.evp, Mp = r4(L.M);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // persistent zone
    float *F = smem;                                  // [3T][2H]   updated_feature (explainer.py:185)
    float *U = F + (size_t)R3 * H2;                   // union zone
    // phase 1 views
    float *X = U;                                     // [3T][evp]  event features (:179)
    float *S = X + (size_t)R3 * evp;                  // [3T][Dp]   src node feats -> src + relu(tgt + evt)
    float *G = S + (size_t)R3 * Dp;                   // [3T][Dp]   tgt node feats -> tgt + relu(src + evt)
    float *E = G + (size_t)R3 * Dp;                   // [3T][Dp]   lin_event output
    float *Z = E + (size_t)R3 * Dp;                   // [6T][H]    relu(MLP.0(.)) for both orientations
    // phase 2 views (alias phase 1)
    float *Q = U;                                     // [3T][2H]   rows < T: W1 f2 ; rows T + 2m + k: W2 f_k
    float *O = Q + (size_t)R3 * H2;                   // [T][2H]
    float *A1 = O + (size_t)T * H2;                   // [T][H]
    float *HC = A1 + (size_t)T * H;                   // [T][Mp]    [attention out | one-hot]
    float *M0 = HC + (size_t)T * Mp;                  // [T][Mp]
    float *M1 = M0 + (size_t)T * Mp;                  // [T][H]
    const float *__restrict__ freq = blob + L.freq;
    const float *__restrict__ phase = blob + L.phase;
    const int64_t n_tiles = (a.n_motifs + T - 1) / T;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t m0 = tile * T;
        __syncthreads();
        // ---------------- P0: gathers + TimeEncode, one warp per event row ----------------
        for (int r = warp; r < R3; r += kEncWarps) {
            const int m = r / 3, pos = r - 3 * m;
            const int64_t gm = m0 + m;
            float *x = X + (size_t)r * evp, *s = S + (size_t)r * Dp, *g = G + (size_t)r * Dp;
            if (gm >= a.n_motifs) {
                for (int k = lane; k < evp; k += 32) x[k] = 0.f;
                for (int k = lane; k < Dp; k += 32) { s[k] = 0.f; g[k] = 0.f; }
                continue;
            }
            const int64_t e = a.eidx[gm * 3 + pos];
            const int64_t ns = a.nodes[gm * 6 + 2 * pos], nt = a.nodes[gm * 6 + 2 * pos + 1];
            const float dt = __fsub_rn(a.t[gm * 3 + 2], a.t[gm * 3 + pos]);          // :326
            const bool e_ok = e >= 0 && e < a.n_edge_rows, s_ok = ns >= 0 && ns < a.n_node_rows, t_ok = nt >= 0 && nt < a.n_node_rows;
            const float *ef = a.edge_feat + e * L.Ed, *sf = a.node_feat + ns * L.D, *tf = a.node_feat + nt * L.D;
            for (int k = lane; k < L.Ed; k += 32) x[k] = e_ok ? __ldg(ef + k) : 0.f;  // :338
            if (lane < 3) x[L.Ed + lane] = a.eid ? __ldg(a.eid + gm * 9 + pos * 3 + lane) : 0.f;   // :177
            for (int k = lane; k < L.D; k += 32)                                       // TimeEncode :55-58
                x[L.Ed + 3 + k] = cosf(__fadd_rn(__fmul_rn(dt, __ldg(freq + k)), __ldg(phase + k)));
            for (int k = L.ev + lane; k < evp; k += 32) x[k] = 0.f;
            for (int k = lane; k < Dp; k += 32) {                                      // :348-351
                s[k] = (k < L.D && s_ok) ? __ldg(sf + k) : 0.f;
                g[k] = (k < L.D && t_ok) ? __ldg(tf + k) : 0.f;
            }
        }
        __syncthreads();
        // ---------------- P1: event = lin_event(event_features)  (:93) ----------------
        warp_gemm<CN_D>(R3, L.evt, blob, [&](int r) { return X + (size_t)r * evp; },
                        [&](int r, int c, float v) { E[(size_t)r * Dp + c] = v; }, warp, lane);
        __syncthreads();
        // msg = relu(other + event); input of MLP = self + msg, both orientations in place (:94-95,182-184)
        for (int i = threadIdx.x; i < R3 * Dp; i += kEncThreads) {
            const int c = i % Dp;
            if (c < L.D) {
                const float s = S[i], g = G[i], e = E[i];
                S[i] = s + fmaxf(g + e, 0.f);
                G[i] = g + fmaxf(s + e, 0.f);
            }
        }
        __syncthreads();
        // ---------------- P2: event_conv.MLP on 6T rows ----------------
        warp_gemm<CN_H>(2 * R3, L.g0, blob, [&](int r) { return r < R3 ? S + (size_t)r * Dp : G + (size_t)(r - R3) * Dp; },
                        [&](int r, int c, float v) { Z[(size_t)r * H + c] = fmaxf(v, 0.f); }, warp, lane);
        __syncthreads();
        warp_gemm<CN_H>(2 * R3, L.g2, blob, [&](int r) { return Z + (size_t)r * H; },
                        [&](int r, int c, float v) { if (r < R3) F[(size_t)r * H2 + c] = v; else F[(size_t)(r - R3) * H2 + H + c] = v; }, warp, lane);
        __syncthreads();
        // ---------------- P3: attention (:799-846) ----------------
        warp_gemm<CN_2H>(T, L.w1, blob, [&](int r) { return F + (size_t)(3 * r + 2) * H2; },
                         [&](int r, int c, float v) { Q[(size_t)r * H2 + c] = v; }, warp, lane);
        warp_gemm<CN_2H>(2 * T, L.w2, blob, [&](int r) { return F + (size_t)(3 * (r >> 1) + (r & 1)) * H2; },
                         [&](int r, int c, float v) { Q[(size_t)(T + r) * H2 + c] = v; }, warp, lane);
        __syncthreads();
        for (int m = warp; m < T; m += kEncWarps) {
            const int64_t gm = m0 + m;
            const float *q = Q + (size_t)m * H2, *k0 = Q + (size_t)(T + 2 * m) * H2, *k1 = k0 + H2;
            float s0 = 0.f, s1 = 0.f;
            for (int c = lane; c < H2; c += 32) { s0 = fmaf(q[c], k0[c], s0); s1 = fmaf(q[c], k1[c], s1); }
            for (int o = 16; o; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
            if (L.use_temporal && gm < a.n_motifs) {
                const int64_t b = gm / a.W;
                const float cut = a.cut[b], sd = __fadd_rn(a.std_[b / a.group], 1e-6f);
                const float d0 = fabsf(__fsub_rn(cut, a.t[gm * 3 + 0])), d1 = fabsf(__fsub_rn(cut, a.t[gm * 3 + 1]));
                const float w0 = expf(__fdiv_rn(-d0, sd)), w1 = expf(__fdiv_rn(-d1, sd));   // :828
                s0 = __fmul_rn(s0, __fadd_rn(0.7f, __fmul_rn(0.3f, w0)));                   // :836
                s1 = __fmul_rn(s1, __fadd_rn(0.7f, __fmul_rn(0.3f, w1)));
            }
            const float mx = fmaxf(s0, s1), e0 = expf(s0 - mx), e1 = expf(s1 - mx);
            const float al0 = e0 / (e0 + e1), al1 = e1 / (e0 + e1);                         // softmax :839
            const float *f2 = F + (size_t)(3 * m + 2) * H2;
            for (int c = lane; c < H2; c += 32) O[(size_t)m * H2 + c] = f2[c] + fmaf(al0, k0[c], al1 * k1[c]);   // :841-842
        }
        __syncthreads();
        warp_gemm<CN_H>(T, L.a0, blob, [&](int r) { return O + (size_t)r * H2; },
                        [&](int r, int c, float v) { A1[(size_t)r * H + c] = fmaxf(v, 0.f); }, warp, lane);
        __syncthreads();
        warp_gemm<CN_H>(T, L.a3, blob, [&](int r) { return A1 + (size_t)r * H; },
                        [&](int r, int c, float v) { HC[(size_t)r * Mp + c] = v; }, warp, lane);
        for (int i = threadIdx.x; i < T * (Mp - H); i += kEncThreads) {                     // one-hot category (:308-315)
            const int m = i / (Mp - H), c = i - m * (Mp - H);
            const int64_t gm = m0 + m;
            const int cat = gm < a.n_motifs && a.cat ? a.cat[gm] : 255;
            HC[(size_t)m * Mp + H + c] = (L.if_cat && c == cat) ? 1.f : 0.f;
        }
        __syncthreads();
        // ---------------- P4: MLP + sigmoid (:123-125,200) ----------------
        warp_gemm<CN_M>(T, L.m0, blob, [&](int r) { return HC + (size_t)r * Mp; },
                        [&](int r, int c, float v) { M0[(size_t)r * Mp + c] = fmaxf(v, 0.f); }, warp, lane);
        for (int i = threadIdx.x; i < T * (Mp - L.M); i += kEncThreads) { const int m = i / (Mp - L.M); M0[(size_t)m * Mp + L.M + (i - m * (Mp - L.M))] = 0.f; }
        __syncthreads();
        warp_gemm<CN_H>(T, L.m3, blob, [&](int r) { return M0 + (size_t)r * Mp; },
                        [&](int r, int c, float v) { M1[(size_t)r * H + c] = fmaxf(v, 0.f); }, warp, lane);
        __syncthreads();
        for (int m = warp; m < T; m += kEncWarps) {
            const int64_t gm = m0 + m;
            float z = 0.f;
            for (int c = lane; c < H; c += 32) z = fmaf(M1[(size_t)m * H + c], __ldg(blob + L.m5.w + (size_t)c * L.m5.Np), z);
            for (int o = 16; o; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
            if (lane == 0 && gm < a.n_motifs) { z += __ldg(blob + L.m5.b); a.scores[gm] = 1.f / (1.f + expf(-z)); }
        }
    }
}

}  // namespace tm

using namespace tmb;

extern "C" int64_t tm_encoder_blob_floats(const tm_encoder_desc *desc) { return desc ? make_layout(*desc).total + tc_blob_floats(*desc) : -1; }

extern "C" int64_t tm_encoder_workspace_floats(const tm_encoder_desc *desc, int64_t B, int64_t W, int64_t group) {
    if (!desc || B < 0 || group <= 0) return -1;
    const int64_t n_std = (std::max<int64_t>(32, (B + group - 1) / group) + 31) & ~(int64_t)31;   // per-batch std, then the scorer's scratch (16-byte aligned)
    const int64_t slab = std::min<int64_t>(tc_slab_motifs(), (std::max<int64_t>(B * W, 1) + 127) / 128 * 128);   // whole tiles of 128 motifs
    const int64_t full = tc_slab_motifs();
    const int64_t n_g = ((desc->node_dim + 7) / 8 * 8 + 31) / 32;           // MLP.0 K chunks: the E scratch of the drain mode (node_dim > 32) holds n_g slabs per CTA
    // h scratch of the resident CTAs (192 KB each) + the tile counter, the Y slabs of the walk groups (32 KB each) and, at node_dim > 32, the E scratch
    // (n_g slabs each) behind it: the factor 2 on the h scratch covers the Y slabs; with an E scratch the whole is sized for a full grid
    return n_std + 2 * (n_g > 1 ? full : slab) * 3 * 2 * desc->hid_dim + (n_g > 1 ? 64 + full / 128 * n_g * 4096 : 0);
}

extern "C" int tm_encoder_project_edges(const tm_encoder_desc *desc, const float *d_blob, const float *d_edge_feat, int64_t n_edge_rows, float *d_out,
                                        int device, tm_stream stream) {
    if (!desc || !d_blob || n_edge_rows < 0 || (n_edge_rows > 0 && (!d_edge_feat || !d_out))) { set_error("tm_encoder_project_edges: bad argument"); return TM_ERR_ARG; }
    if (desc->node_dim < 1 || desc->node_dim > 256 || desc->edge_dim < 1 || desc->edge_dim > 1024) { set_error("tm_encoder_project_edges: node_dim must be in [1,256], edge_dim in [1,1024]"); return TM_ERR_UNSUPPORTED; }
    TM_DEVICE(device);
    return tc_project_edges(*desc, d_blob + make_layout(*desc).total, d_edge_feat, n_edge_rows, d_out, (cudaStream_t)stream);
}

extern "C" int tm_encoder_pack(const tm_encoder_desc *desc, const tm_encoder_params *p, float *h_blob) {
    if (!desc || !p || !h_blob) { set_error("tm_encoder_pack: bad argument"); return TM_ERR_ARG; }
    const EncLayout L = make_layout(*desc);
    memset(h_blob, 0, sizeof(float) * L.total);
    auto put = [&](const Lin &l, const float *w, const float *b) {   // nn.Linear weight [N][K] -> Wt [Kp][Np]
        if (!w || !b) return false;
        for (int n = 0; n < l.N; ++n) {
            for (int k = 0; k < l.K; ++k) h_blob[l.w + (int64_t)k * l.Np + n] = w[(int64_t)n * l.K + k];
            h_blob[l.b + n] = b[n];
        }
        return true;
    };
    bool ok = put(L.evt, p->lin_event_w, p->lin_event_b) && put(L.g0, p->gcn0_w, p->gcn0_b) && put(L.g2, p->gcn2_w, p->gcn2_b) &&
              put(L.w1, p->att_w1_w, p->att_w1_b) && put(L.w2, p->att_w2_w, p->att_w2_b) && put(L.a0, p->att_mlp0_w, p->att_mlp0_b) &&
              put(L.a3, p->att_mlp3_w, p->att_mlp3_b) && put(L.m0, p->mlp0_w, p->mlp0_b) && put(L.m3, p->mlp3_w, p->mlp3_b) &&
              put(L.m5, p->mlp5_w, p->mlp5_b) && p->basis_freq && p->phase;
    if (!ok) { set_error("tm_encoder_pack: a parameter pointer is null"); return TM_ERR_ARG; }
    for (int k = 0; k < L.D; ++k) { h_blob[L.freq + k] = p->basis_freq[k]; h_blob[L.phase + k] = p->phase[k]; }
    return tc_pack(*desc, *p, h_blob + L.total);     // second half of the blob: pre-split, pre-tiled tcgen05 operands
}

// shared-memory floats one tile of T motifs needs (layout of encode_kernel)
static int64_t tile_floats(const EncLayout &L, int T) {
    const int64_t R3 = 3 * T, H = L.H, Mp = r4(L.M);
    const int64_t p1 = R3 * L.evp + 3 * R3 * L.Dp + 2 * R3 * H;
    const int64_t p2 = R3 * 2 * H + (int64_t)T * 2 * H + (int64_t)T * H + 2 * (int64_t)T * Mp + (int64_t)T * H;
    return R3 * 2 * H + std::max(p1, p2);
}

template <int CN_D>
static int launch_encode(const EncLayout &L, const float *blob, const EncArgs &a, int grid, size_t smem, cudaStream_t st) {
    auto k = encode_kernel<CN_D, 2, 4, 3>;
    static bool attr_set[64] = {false};
    int dev = 0;
    TM_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !attr_set[dev]) {
        TM_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[dev] = true;
    }
    k<<<grid, kEncThreads, smem, st>>>(L, blob, a);
    TM_LAUNCH_CHECK();
    return TM_OK;
}

static int encode_score_impl(const tm_encoder_desc *desc, const float *d_blob, int64_t B, int64_t W, int64_t group,
                             const int32_t *d_nodes, const int32_t *d_eidx, const float *d_t, const uint8_t *d_cat,
                             const float *d_cut_time, const float *d_edge_identity,
                             const float *d_node_feat, int64_t n_node_rows, const float *d_edge_feat, int64_t n_edge_rows,
                             float *d_workspace, float *d_scores, float *d_y, float *const *peer_scores, int n_peers, int device, tm_stream stream) {
    if (!desc || !d_blob || B < 0 || W <= 0 || group <= 0 ||
        (B > 0 && (!d_nodes || !d_eidx || !d_t || !d_cut_time || !d_node_feat || !d_edge_feat || !d_workspace || !d_scores))) {
        set_error("tm_encode_score: bad argument");
        return TM_ERR_ARG;
    }
    if (desc->if_cat && !d_cat && B > 0) { set_error("tm_encode_score: if_cat needs d_cat"); return TM_ERR_ARG; }
    if (desc->hid_dim != 64 && desc->hid_dim != 32) { set_error("tm_encode_score: hid_dim %d unsupported (64 and 32, the reference's defaults)", desc->hid_dim); return TM_ERR_UNSUPPORTED; }
    if (desc->node_dim < 1 || desc->node_dim > 256 || desc->edge_dim < 1 || desc->edge_dim > 1024) { set_error("tm_encode_score: node_dim must be in [1,256], edge_dim in [1,1024]"); return TM_ERR_UNSUPPORTED; }
    if (B == 0) return TM_OK;
    TM_DEVICE(device);
    const EncLayout L = make_layout(*desc);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_groups = (B + group - 1) / group;
    if (desc->use_temporal) {
        time_std_kernel<<<(unsigned)n_groups, 256, 0, st>>>(B, W, group, d_t, d_cut_time, d_workspace);
        TM_LAUNCH_CHECK();
    }
    const char *which = getenv("TEMPME_ENCODER");     // "ffma" selects the fp32 CUDA-core kernel (A/B validation); default: tcgen05
    if (desc->edge_projected || desc->edge_identity_u8 || d_y || n_peers > 0 || desc->hid_dim != 64 || !which || strcmp(which, "ffma") != 0) {
        const int64_t n_std = (std::max<int64_t>(32, n_groups) + 31) & ~(int64_t)31;
        return tc_encode_score(*desc, d_blob + L.total, B, W, group, d_nodes, d_eidx, d_t, d_cat, d_cut_time, d_edge_identity, d_node_feat,
                               n_node_rows, d_edge_feat, n_edge_rows, d_workspace, d_workspace + n_std, d_scores, d_y, peer_scores, n_peers, device, st);
    }
    const int64_t cap = (227 * 1024 - 1024) / 4;
    int T = 64;
    while (T > 4 && tile_floats(L, T) > cap) T -= 4;
    if (tile_floats(L, T) > cap) { set_error("tm_encode_score: feature dims too large for one tile in shared memory"); return TM_ERR_UNSUPPORTED; }
    EncArgs a;
    a.n_motifs = B * W; a.W = W; a.group = group; a.nodes = d_nodes; a.eidx = d_eidx; a.t = d_t; a.cat = d_cat; a.cut = d_cut_time;
    a.eid = d_edge_identity; a.node_feat = d_node_feat; a.edge_feat = d_edge_feat; a.std_ = d_workspace;
    a.n_node_rows = n_node_rows; a.n_edge_rows = n_edge_rows; a.scores = d_scores; a.T = T;
    const int64_t n_tiles = (a.n_motifs + T - 1) / T;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int grid = (int)std::min<int64_t>(n_tiles, sms);
    const size_t smem = sizeof(float) * tile_floats(L, T);
    switch (L.evt.Np / 32) {
        case 1: return launch_encode<1>(L, d_blob, a, grid, smem, st);
        case 2: return launch_encode<2>(L, d_blob, a, grid, smem, st);
        case 3: return launch_encode<3>(L, d_blob, a, grid, smem, st);
        case 4: return launch_encode<4>(L, d_blob, a, grid, smem, st);
        case 5: return launch_encode<5>(L, d_blob, a, grid, smem, st);
        case 6: return launch_encode<6>(L, d_blob, a, grid, smem, st);
        case 7: return launch_encode<7>(L, d_blob, a, grid, smem, st);
        default: return launch_encode<8>(L, d_blob, a, grid, smem, st);
    }
}

extern "C" int tm_encode_score(const tm_encoder_desc *desc, const float *d_blob, int64_t B, int64_t W, int64_t group,
                               const int32_t *d_nodes, const int32_t *d_eidx, const float *d_t, const uint8_t *d_cat,
                               const float *d_cut_time, const float *d_edge_identity,
                               const float *d_node_feat, int64_t n_node_rows, const float *d_edge_feat, int64_t n_edge_rows,
                               float *d_workspace, float *d_scores, int device, tm_stream stream) {
    return encode_score_impl(desc, d_blob, B, W, group, d_nodes, d_eidx, d_t, d_cat, d_cut_time, d_edge_identity, d_node_feat, n_node_rows,
                             d_edge_feat, n_edge_rows, d_workspace, d_scores, nullptr, nullptr, 0, device, stream);
}

extern "C" int tm_encode_score_gather(const tm_encoder_desc *desc, const float *d_blob, int64_t B, int64_t W, int64_t group,
                                      const int32_t *d_nodes, const int32_t *d_eidx, const float *d_t, const uint8_t *d_cat,
                                      const float *d_cut_time, const float *d_edge_identity,
                                      const float *d_node_feat, int64_t n_node_rows, const float *d_edge_feat, int64_t n_edge_rows,
                                      float *d_workspace, float *d_scores, const uint64_t *h_peer_scores, int n_peers, int device, tm_stream stream) {
    float *peers[8] = {nullptr};
    if (n_peers < 0 || n_peers > 7 || (n_peers > 0 && !h_peer_scores)) { set_error("tm_encode_score_gather: 0 <= n_peers <= 7"); return TM_ERR_ARG; }
    for (int p = 0; p < n_peers; ++p) {
        peers[p] = reinterpret_cast<float *>((uintptr_t)h_peer_scores[p]);
        if (!peers[p]) { set_error("tm_encode_score_gather: null peer pointer"); return TM_ERR_ARG; }
    }
    return encode_score_impl(desc, d_blob, B, W, group, d_nodes, d_eidx, d_t, d_cat, d_cut_time, d_edge_identity, d_node_feat, n_node_rows,
                             d_edge_feat, n_edge_rows, d_workspace, d_scores, nullptr, peers, n_peers, device, stream);
}

extern "C" int tm_encode_attention(const tm_encoder_desc *desc, const float *d_blob, int64_t B, int64_t W, int64_t group,
                                   const int32_t *d_nodes, const int32_t *d_eidx, const float *d_t, const uint8_t *d_cat,
                                   const float *d_cut_time, const float *d_edge_identity,
                                   const float *d_node_feat, int64_t n_node_rows, const float *d_edge_feat, int64_t n_edge_rows,
                                   float *d_workspace, float *d_scores, float *d_y, int device, tm_stream stream) {
    if (!d_y && B > 0) { set_error("tm_encode_attention: d_y is required"); return TM_ERR_ARG; }
    return encode_score_impl(desc, d_blob, B, W, group, d_nodes, d_eidx, d_t, d_cat, d_cut_time, d_edge_identity, d_node_feat, n_node_rows,
                             d_edge_feat, n_edge_rows, d_workspace, d_scores, d_y, nullptr, 0, device, stream);
}
