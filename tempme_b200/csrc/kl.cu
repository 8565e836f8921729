// TempME.kl_loss, forward value (reference models/explainer.py:432-453): the regulariser the reference's eval loops log next to the
// prediction loss (temp_exp_main.py:326-328, :459-461).  prob [B, W] are the motif scores, cat [B, W] the walk classes.
//   prior == "empirical":  s_b = mean_w prob;  m_bc = mean of prob over the walks of class c (0 if none; torch_scatter "mean");
//                          e_bc = s_b m_bc;  n_c = target * null[c]  (null = list(null_model.values()), paired by position);
//                          loss = mean_{b,c} [ (1 - s_b) log((1 - s_b) / (1 - target + 1e-6) + 1e-6) + e_bc log(e_bc / (n_c + 1e-6) + 1e-6) ]
//   otherwise:             loss = mean_{b,w} [ p log(p / target + 1e-6) + (1 - p) log((1 - p) / (1 - target + 1e-6) + 1e-6) ]
// A warp per root writes the root's contribution in double; one block then adds the B contributions in a fixed order, so the value is
// deterministic.
#include <stdint.h>

#include "common.cuh"

namespace tmb {

__device__ __forceinline__ float warp_sum(float v) {
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr int kKlWarps = 8;

__global__ void kl_root_kernel(int64_t B, int W, const float *__restrict__ prob, const uint8_t *__restrict__ cat,
                               const float *__restrict__ null_vals, int n_cat, float target, int empirical, double *__restrict__ part) {
    const int64_t b = (int64_t)blockIdx.x * kKlWarps + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const float *p = prob + b * W;
    const float one_t = __fadd_rn(__fsub_rn(1.f, target), 1e-6f);
    auto clampp = [](float x) { return fminf(fmaxf(x, 1e-6f), 1.f - 1e-6f); };                       // :435
    double acc = 0;
    if (!empirical) {
        for (int w = lane; w < W; w += 32) {
            const float x = clampp(p[w]), y = __fsub_rn(1.f, x);
            acc += (double)__fadd_rn(__fmul_rn(x, logf(__fadd_rn(__fdiv_rn(x, target), 1e-6f))),
                                     __fmul_rn(y, logf(__fadd_rn(__fdiv_rn(y, one_t), 1e-6f))));      // :450-451
        }
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) part[b] = acc / (double)W;
        return;
    }
    const uint8_t *c = cat + b * W;
    float s = 0.f;
    for (int w = lane; w < W; w += 32) s += clampp(p[w]);
    s = __fdiv_rn(warp_sum(s), (float)W);                                                             // :438
    const float r = __fsub_rn(1.f, s);
    const float head = __fmul_rn(r, logf(__fadd_rn(__fdiv_rn(r, one_t), 1e-6f)));                     // :447, broadcast over the classes
    for (int k = 0; k < n_cat; ++k) {
        float sum = 0.f, cnt = 0.f;
        for (int w = lane; w < W; w += 32)
            if (c[w] == k) { sum += clampp(p[w]); cnt += 1.f; }
        sum = warp_sum(sum); cnt = warp_sum(cnt);
        const float e = __fmul_rn(s, __fdiv_rn(sum, fmaxf(cnt, 1.f)));                                // :443-444
        const float n = __fadd_rn(__fmul_rn(target, null_vals[k]), 1e-6f);                            // :445
        acc += (double)__fadd_rn(head, __fmul_rn(e, logf(__fadd_rn(__fdiv_rn(e, n), 1e-6f))));        // :447-448
    }
    if (lane == 0) part[b] = acc / (double)n_cat;
}

__global__ void kl_sum_kernel(int64_t B, const double *__restrict__ part, float *__restrict__ loss) {
    __shared__ double sh[1024];
    double v = 0;
    for (int64_t i = threadIdx.x; i < B; i += blockDim.x) v += part[i];
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = blockDim.x >> 1; o; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss = (float)(sh[0] / (double)B);
}

}  // namespace tmb

using namespace tmb;

extern "C" int tm_kl_loss(int64_t B, int64_t W, const float *d_prob, const uint8_t *d_cat, const float *d_null_values, int n_cat,
                          float target, int empirical, double *d_workspace, float *d_loss, tm_stream stream) {
    if (B <= 0 || W <= 0 || !d_prob || !d_workspace || !d_loss || (empirical && (!d_cat || !d_null_values || n_cat <= 0 || n_cat > 255))) {
        set_error("tm_kl_loss: bad argument");
        return TM_ERR_ARG;
    }
    TM_DEVICE(device_of(d_loss));
    kl_root_kernel<<<(unsigned)((B + kKlWarps - 1) / kKlWarps), kKlWarps * 32, 0, (cudaStream_t)stream>>>(
        B, (int)W, d_prob, d_cat, d_null_values, n_cat, target, empirical, d_workspace);
    TM_LAUNCH_CHECK();
    kl_sum_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(B, d_workspace, d_loss);
    TM_LAUNCH_CHECK();
    return TM_OK;
}
