// Beta(max(10 p, 1), max(10 (1 - p), 1)) draws for TempME.beta_sample(training=True) (reference models/explainer.py:421-430): shared by the
// stand-alone sampler (train.cu) and the fused motif -> edge aggregation (edge_imp.cu).
#pragma once
#include "common.cuh"

namespace tmb {

// uniform in (0, 1] from 32 random bits; standard normal by Box-Muller from two of them
__device__ __forceinline__ float u01(uint32_t r) { return ((float)(r >> 8) + 1.0f) * (1.0f / 16777216.0f); }

// Gamma(a, 1), a >= 1 (Marsaglia & Tsang 2000).  Counter = (attempt, index lo, index hi, stream), key = seed.
__device__ __noinline__ static float gamma_mt(float a, uint64_t seed, uint64_t idx, uint32_t stream) {
    const float d = a - (1.0f / 3.0f), c = rsqrtf(9.0f * d);
    float out = d;                                                  // value if the bounded loop is ever exhausted (probability < 1e-60)
    for (uint32_t it = 0; it < 64; ++it) {
        uint32_t o[4];
        philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), it, (uint32_t)idx, (uint32_t)(idx >> 32), stream, o);
        const float r = sqrtf(-2.0f * logf(u01(o[0]))), ang = 6.283185307179586f * u01(o[1]);
        float xs[2] = {r * cosf(ang), r * sinf(ang)};
        const float us[2] = {u01(o[2]), u01(o[3])};
        bool done = false;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (done) continue;
            const float x = xs[k], t = 1.0f + c * x;
            if (t <= 0.f) continue;
            const float v = t * t * t;
            if (logf(us[k]) < 0.5f * x * x + d - d * v + d * logf(v)) { out = d * v; done = true; }
        }
        if (done) break;
    }
    return out;
}

// One Beta draw for probability p (explainer.py:423-427); g1 / g2 optional outputs
__device__ __forceinline__ float beta_draw(float p, uint64_t seed, uint64_t idx, float *g1o, float *g2o) {
    const float alpha = fmaxf(__fmul_rn(p, 10.f), 1.f), beta = fmaxf(__fmul_rn(__fsub_rn(1.f, p), 10.f), 1.f);
    const float g1 = gamma_mt(alpha, seed, idx, 0x42u), g2 = gamma_mt(beta, seed, idx, 0x43u);
    if (g1o) *g1o = g1;
    if (g2o) *g2o = g2;
    return g1 / (g1 + g2);
}

}  // namespace tmb
