// Enhance path: TempME.compute_walk_importance and the walk-weighted sum of enhance_predict_walks in eval mode
// (reference models/explainer.py:222-306).  The attention output of a walk is attention.MLP.3(y) + b with y = the hidden vector the
// scorer kernel writes (tm_encode_attention); the weighted sum over a root's walks commutes with that Linear:
//     sum_w weight_w (A3 y_w + b3) = A3 (sum_w weight_w y_w) + b3 sum_w weight_w
// so one 64 x 64 mat-vec per root replaces one per walk.
#include <stdint.h>

#include <algorithm>

#include "common.cuh"

namespace tmb {

__device__ __forceinline__ double block_sum(double v, double *sh) {          // all threads get the sum; blockDim.x <= 1024
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double s = 0;
    for (int i = 0; i < nw; ++i) s += sh[i];
    return s;
}

// One block per reference batch (`group` roots): the statistics of compute_walk_importance run over the whole [group, W] batch
// (:281 time_diff.std(), :295 avg_degree.mean() / .std(); torch's std is unbiased).
__global__ void walk_importance_kernel(int64_t B, int W, int64_t group, const float *__restrict__ t, const int32_t *__restrict__ nodes,
                                       const float *__restrict__ cut, const float *__restrict__ degree, int64_t n_nodes,
                                       float *__restrict__ weights) {
    __shared__ double sh[32];
    const int64_t b0 = (int64_t)blockIdx.x * group, nb = min(group, B - b0);
    const int64_t n = nb * W;
    auto walk = [&](int64_t i, float &diff, float &avg) {
        const int64_t w = b0 * W + i;
        const float mt = fmaxf(fmaxf(t[w * 3], t[w * 3 + 1]), t[w * 3 + 2]);           // :274
        diff = fabsf(__fsub_rn(cut[w / W], mt));                                        // :276-277
        float sum = 0.f; int cnt = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int32_t v = nodes[w * 6 + k];
            if (v > 0) { ++cnt; sum = __fadd_rn(sum, v < n_nodes ? degree[v] : 0.f); }  // :285-291
        }
        avg = __fdiv_rn(sum, __fadd_rn((float)cnt, 1e-6f));                             // :292
    };
    double sd = 0, sa = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { float d, a; walk(i, d, a); sd += d; sa += a; }
    const double md = block_sum(sd, sh) / (double)n, ma = block_sum(sa, sh) / (double)n;
    double vd = 0, va = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { float d, a; walk(i, d, a); vd += (d - md) * (d - md); va += (a - ma) * (a - ma); }
    vd = block_sum(vd, sh); va = block_sum(va, sh);
    const float std_d = n > 1 ? (float)sqrt(vd / (double)(n - 1)) : __int_as_float(0x7fc00000);
    const float std_a = n > 1 ? (float)sqrt(va / (double)(n - 1)) : __int_as_float(0x7fc00000), mean_a = (float)ma;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        float d, a; walk(i, d, a);
        const float rec = expf(__fdiv_rn(__fdiv_rn(-d, __fadd_rn(std_d, 1e-6f)), 1.0f));                            // :281
        const float dw = 1.f / (1.f + expf(-__fdiv_rn(__fsub_rn(a, mean_a), __fadd_rn(std_a, 1e-6f))));             // :295
        weights[b0 * W + i] = __fadd_rn(__fmul_rn(0.5f, rec), __fmul_rn(0.5f, dw));                                 // :298
    }
    __syncthreads();
    for (int64_t r = threadIdx.x; r < nb; r += blockDim.x) {       // normalise so that a root's weights sum to W (:301)
        float s = 0.f;
        for (int w = 0; w < W; ++w) s = __fadd_rn(s, weights[(b0 + r) * W + w]);
        const float den = __fadd_rn(__fdiv_rn(s, (float)W), 1e-6f);
        for (int w = 0; w < W; ++w) weights[(b0 + r) * W + w] = __fdiv_rn(weights[(b0 + r) * W + w], den);
    }
}

// One block of H threads per root: weighted sum of the walks' hidden vectors, attention.MLP.3 once, class counts (:245-253).
// The sum over the W walks and the 64-term dot products accumulate in double: the result is a signed sum whose fp32 evaluation order
// would otherwise decide its last digits (the reference's own order -- Linear per walk, then a pairwise torch.sum -- is a different one);
// against the float64 arbiter this keeps the embedding closer to the exact value than the reference's fp32 evaluation is.
__global__ void enhance_reduce_kernel(int64_t B, int W, int H, const float *__restrict__ y, const float *__restrict__ weights,
                                      const float *__restrict__ a3_w, const float *__restrict__ a3_b, const uint8_t *__restrict__ cat,
                                      int out_dim, float *__restrict__ out) {
    extern __shared__ double acc[];            // [H]
    const int64_t b = blockIdx.x;
    const int j = threadIdx.x;
    double s = 0.0, sw = 0.0;
    for (int w = 0; w < W; ++w) {
        const double wt = (double)weights[b * W + w];
        s += wt * (double)y[(b * W + w) * H + j];
        sw += wt;
    }
    acc[j] = s;
    __syncthreads();
    double o = (double)a3_b[j] * sw;
    for (int i = 0; i < H; ++i) o += (double)a3_w[j * H + i] * acc[i];
    out[b * out_dim + j] = (float)o;
    if (cat && j < 12) {
        int c = 0;
        for (int w = 0; w < W; ++w) c += cat[b * W + w] == j;
        out[b * out_dim + H + j] = (float)c;
    }
}

}  // namespace tmb

using namespace tmb;

extern "C" int tm_walk_importance(int64_t B, int64_t W, int64_t group, const float *d_t, const int32_t *d_nodes, const float *d_cut_time,
                                  const float *d_node_degree, int64_t n_nodes, float *d_weights, tm_stream stream) {
    if (B < 0 || W <= 0 || group <= 0 || (B > 0 && (!d_t || !d_nodes || !d_cut_time || !d_node_degree || !d_weights))) {
        set_error("tm_walk_importance: bad argument");
        return TM_ERR_ARG;
    }
    if (B == 0) return TM_OK;
    TM_DEVICE(device_of(d_weights));
    walk_importance_kernel<<<(unsigned)((B + group - 1) / group), 256, 0, (cudaStream_t)stream>>>(B, (int)W, group, d_t, d_nodes, d_cut_time,
                                                                                                   d_node_degree, n_nodes, d_weights);
    TM_LAUNCH_CHECK();
    return TM_OK;
}

extern "C" int tm_enhance_reduce(int64_t B, int64_t W, int hid_dim, const float *d_y, const float *d_weights, const float *d_att_mlp3_w,
                                 const float *d_att_mlp3_b, const uint8_t *d_cat_or_null, float *d_out, tm_stream stream) {
    if (B < 0 || W <= 0 || hid_dim < 12 || hid_dim > 1024 || (B > 0 && (!d_y || !d_weights || !d_att_mlp3_w || !d_att_mlp3_b || !d_out))) {
        set_error("tm_enhance_reduce: bad argument");
        return TM_ERR_ARG;
    }
    if (B == 0) return TM_OK;
    TM_DEVICE(device_of(d_out));
    const int out_dim = hid_dim + (d_cat_or_null ? 12 : 0);
    enhance_reduce_kernel<<<(unsigned)B, hid_dim, sizeof(double) * hid_dim, (cudaStream_t)stream>>>(B, (int)W, hid_dim, d_y, d_weights, d_att_mlp3_w,
                                                                                                   d_att_mlp3_b, d_cat_or_null, out_dim, d_out);
    TM_LAUNCH_CHECK();
    return TM_OK;
}
