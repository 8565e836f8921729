// Training-side kernels of the explainer:
//   * Beta sampling of TempME.beta_sample(training=True) (reference models/explainer.py:421-430,
//     torch.distributions.Beta(alpha, beta).rsample()): x = g1 / (g1 + g2) with g1 ~ Gamma(alpha), g2 ~ Gamma(beta) drawn by
//     Marsaglia-Tsang squeeze/rejection from counter-based Philox4x32-10 normals and uniforms.  alpha = max(10 p, 1) and
//     beta = max(10 (1 - p), 1) are never below 1, so the a < 1 boost step is not needed.  The two gammas are returned as well: the
//     reparameterised gradient is dx/dalpha = (dx/dg1) (dg1/dalpha) with dg/da the implicit gamma derivative.
//   * the gradient of TempME.kl_loss (models/explainer.py:432-453) with respect to the motif scores.
#include <math.h>
#include <stdint.h>

#include <algorithm>

#include "beta.cuh"
#include "common.cuh"

namespace tmb {

__global__ void beta_sample_kernel(int64_t n, const float *__restrict__ prob, const int32_t *__restrict__ node, uint64_t seed, uint64_t offset,
                                   float *__restrict__ out, float *__restrict__ g1, float *__restrict__ g2) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float a = 0.f, b = 0.f;
        const float x = beta_draw(prob[i], seed, offset + (uint64_t)i, &a, &b);
        out[i] = (node && node[i] == 0) ? 0.f : x;                  // padding mask (:400-404)
        if (g1) g1[i] = a;
        if (g2) g2[i] = b;
    }
}

__device__ __forceinline__ float warp_sum_f(float v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// d loss / d prob[b, w] of kl_loss, scaled by *grad_out.  One warp per root; same fp32 arithmetic for s, the class means and e as the
// forward kernel (kl.cu).  clamp(prob, 1e-6, 1 - 1e-6) passes the gradient inside the closed interval and blocks it outside (:435).
constexpr int kKlgWarps = 8;
__global__ void kl_grad_kernel(int64_t B, int W, const float *__restrict__ prob, const uint8_t *__restrict__ cat,
                               const float *__restrict__ null_vals, int n_cat, float target, int empirical, const float *__restrict__ grad_out,
                               float *__restrict__ grad_prob) {
    const int64_t b = (int64_t)blockIdx.x * kKlgWarps + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const float go = grad_out ? *grad_out : 1.f;
    const float *p = prob + b * W;
    float *gp = grad_prob + b * W;
    const float one_t = __fadd_rn(__fsub_rn(1.f, target), 1e-6f);
    auto clampp = [](float x) { return fminf(fmaxf(x, 1e-6f), 1.f - 1e-6f); };
    auto inside = [](float x) { return x >= 1e-6f && x <= 1.f - 1e-6f; };
    if (!empirical) {
        const float scale = go / ((float)B * (float)W);
        for (int w = lane; w < W; w += 32) {
            const float x = clampp(p[w]), y = 1.f - x;
            const float a = x / target, c = y / one_t;
            // d/dx [x log(x/t + eps) + (1-x) log((1-x)/ot + eps)]
            const float d = logf(a + 1e-6f) + a / (a + 1e-6f) - logf(c + 1e-6f) - c / (c + 1e-6f);
            gp[w] = inside(p[w]) ? scale * d : 0.f;
        }
        return;
    }
    const uint8_t *c = cat + b * W;
    float s = 0.f;
    for (int w = lane; w < W; w += 32) s += clampp(p[w]);
    s = __fdiv_rn(warp_sum_f(s), (float)W);
    const float r = 1.f - s, u = r / one_t;
    const float dA = -logf(u + 1e-6f) - u / (u + 1e-6f);              // d/ds of (1 - s) log((1 - s)/ot + eps), once per class (broadcast)
    // per class c: m_c = class mean, e_c = s m_c, dB_c = d/de [e log(e / n_c + eps)].  A walk of class k receives
    //   (1 / W) (n_cat dA + sum_c dB_c m_c)   through s   and   dB_k s / cnt_k   through its own class mean.
    __shared__ float sh_db[kKlgWarps][16], sh_cnt[kKlgWarps][16];
    float *dBc = sh_db[threadIdx.x >> 5], *cntc = sh_cnt[threadIdx.x >> 5];
    float common = (float)n_cat * dA;
    for (int k = 0; k < n_cat; ++k) {
        float sum = 0.f, cnt = 0.f;
        for (int w = lane; w < W; w += 32)
            if (c[w] == k) { sum += clampp(p[w]); cnt += 1.f; }
        sum = warp_sum_f(sum); cnt = warp_sum_f(cnt);
        const float m = __fdiv_rn(sum, fmaxf(cnt, 1.f)), e = __fmul_rn(s, m);
        const float n = __fadd_rn(__fmul_rn(target, null_vals[k]), 1e-6f), q = e / n;
        const float dB = logf(q + 1e-6f) + q / (q + 1e-6f);
        common += dB * m;
        if (lane == 0) { dBc[k] = dB; cntc[k] = fmaxf(cnt, 1.f); }
    }
    __syncwarp();
    const float scale = go / ((float)B * (float)n_cat);
    for (int w = lane; w < W; w += 32) {
        const int k = c[w];
        const float own = k < n_cat ? dBc[k] * s / cntc[k] : 0.f;
        gp[w] = inside(p[w]) ? scale * (common / (float)W + own) : 0.f;
    }
}

}  // namespace tmb

using namespace tmb;

extern "C" int tm_beta_sample(int64_t n, const float *d_prob, const int32_t *d_node_or_null, uint64_t seed, uint64_t offset, float *d_out,
                              float *d_g1_or_null, float *d_g2_or_null, tm_stream stream) {
    if (n < 0 || (n > 0 && (!d_prob || !d_out))) { set_error("tm_beta_sample: bad argument"); return TM_ERR_ARG; }
    if (n == 0) return TM_OK;
    TM_DEVICE(device_of(d_out));
    const int64_t blocks = std::min<int64_t>((n + 255) / 256, 148 * 16);
    beta_sample_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n, d_prob, d_node_or_null, seed, offset, d_out, d_g1_or_null, d_g2_or_null);
    TM_LAUNCH_CHECK();
    return TM_OK;
}

extern "C" int tm_kl_loss_backward(int64_t B, int64_t W, const float *d_prob, const uint8_t *d_cat, const float *d_null_values, int n_cat,
                                   float target, int empirical, const float *d_grad_out_or_null, float *d_grad_prob, tm_stream stream) {
    if (B <= 0 || W <= 0 || !d_prob || !d_grad_prob || (empirical && (!d_cat || !d_null_values || n_cat <= 0 || n_cat > 16))) {
        set_error("tm_kl_loss_backward: bad argument (empirical prior: at most 16 classes)");
        return TM_ERR_ARG;
    }
    TM_DEVICE(device_of(d_grad_prob));
    kl_grad_kernel<<<(unsigned)((B + kKlgWarps - 1) / kKlgWarps), kKlgWarps * 32, 0, (cudaStream_t)stream>>>(
        B, (int)W, d_prob, d_cat, d_null_values, n_cat, target, empirical, d_grad_out_or_null, d_grad_prob);
    TM_LAUNCH_CHECK();
    return TM_OK;
}
