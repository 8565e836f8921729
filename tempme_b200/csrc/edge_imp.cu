// Motif -> edge explanation aggregation: TempME.retrieve_edge_imp_node in eval mode (reference models/explainer.py:354-406).
// The dependency gate runs on the tensor cores (gate_tc_kernel, encoder_tc.cu); this file holds the per-root segmented max /
// gather / Beta mean / padding mask and the C ABI entry points.
#include <stdint.h>

#include <algorithm>

#include "beta.cuh"
#include "common.cuh"

namespace tmb {

// One block per root.  The 3W (edge id, importance) pairs of the root's walks go into an open-addressing table in shared memory
// (atomicMax on the bit pattern: importances are non-negative floats), then every hop slot is one or two probes: the per-root
// scatter(max) + gather of the reference (:389-393) without the |slots| x 3W comparison matrix.  Ids that no walk carries give 0.
__global__ void edge_imp_kernel(int64_t B, int W3, int slots_mask, const int32_t *__restrict__ w_eidx, const float *__restrict__ walk_imp,
                                const float *__restrict__ scores, int K0, const int32_t *__restrict__ h0_node,
                                const int32_t *__restrict__ h0_eidx, int K1, const int32_t *__restrict__ h1_node,
                                const int32_t *__restrict__ h1_eidx, float *__restrict__ imp0, float *__restrict__ imp1, int sample,
                                uint64_t seed) {
    extern __shared__ int32_t sh[];             // keys [slots], then importance bit patterns [slots]
    int32_t *keys = sh, *vals = sh + slots_mask + 1;
    const int64_t b = blockIdx.x;
    for (int i = threadIdx.x; i <= slots_mask; i += blockDim.x) { keys[i] = INT32_MIN; vals[i] = 0; }
    __syncthreads();
    auto hash = [&](int32_t id) { return (int)(((uint32_t)id * 2654435761u) >> 8) & slots_mask; };
    for (int i = threadIdx.x; i < W3; i += blockDim.x) {
        const int32_t id = w_eidx[b * W3 + i];
        const float v = walk_imp ? walk_imp[b * W3 + i] : scores[b * (W3 / 3) + i / 3];      // graphlet_imp.repeat(1,1,3) (:363)
        int slot = hash(id);
        for (;;) {
            const int32_t prev = atomicCAS(&keys[slot], INT32_MIN, id);
            if (prev == INT32_MIN || prev == id) { atomicMax(&vals[slot], __float_as_int(fmaxf(v, 0.f))); break; }
            slot = (slot + 1) & slots_mask;
        }
    }
    __syncthreads();
    for (int s = threadIdx.x; s < K0 + K1; s += blockDim.x) {
        const bool l0 = s < K0;
        const int64_t o = l0 ? b * K0 + s : b * K1 + (s - K0);
        const int32_t id = l0 ? h0_eidx[o] : h1_eidx[o], node = l0 ? h0_node[o] : h1_node[o];
        float m = 0.f;
        for (int slot = hash(id);; slot = (slot + 1) & slots_mask) {
            const int32_t k = keys[slot];
            if (k == id) { m = __int_as_float(vals[slot]); break; }
            if (k == INT32_MIN) break;
        }
        const float alpha = fmaxf(__fmul_rn(m, 10.f), 1.f), beta = fmaxf(__fmul_rn(__fsub_rn(1.f, m), 10.f), 1.f);   // :423-424
        // training: one Beta draw per slot (:426-427; hop-1 slots use draw index o, hop-2 slots o + 2^40); else the mean (:429); padding :400-404
        const float val = sample ? beta_draw(m, seed, (uint64_t)o + (l0 ? 0ull : (1ull << 40)), nullptr, nullptr) : __fdiv_rn(alpha, __fadd_rn(alpha, beta));
        const float out = node == 0 ? 0.f : val;
        (l0 ? imp0 : imp1)[o] = out;
    }
}

}  // namespace tmb

using namespace tmb;

extern "C" int64_t tm_gate_blob_floats(const tm_gate_desc *desc) { return desc ? tc_gate_blob_floats(*desc) : -1; }

extern "C" int tm_gate_pack(const tm_gate_desc *desc, const tm_gate_params *p, float *h_blob) {
    if (!desc || !p || !h_blob || !p->w0 || !p->b0 || !p->w3 || !p->b3 || !p->w6 || !p->b6 || !p->basis_freq || !p->phase) {
        set_error("tm_gate_pack: bad argument");
        return TM_ERR_ARG;
    }
    if (desc->hid_dim != 64 && desc->hid_dim != 32) { set_error("tm_gate_pack: hid_dim %d unsupported (64 and 32, the defaults of temp_exp_main.py / enhance_main.py)", desc->hid_dim); return TM_ERR_UNSUPPORTED; }
    return tc_gate_pack(*desc, *p, h_blob);
}

extern "C" int tm_edge_importance(const tm_gate_desc *desc, const float *d_gate_blob, int64_t B, int64_t W, const float *d_scores,
                                  const int32_t *d_eidx, const float *d_t, const float *d_edge_feat, int64_t n_edge_rows,
                                  int64_t K0, const int32_t *d_h0_node, const int32_t *d_h0_eidx, int64_t K1, const int32_t *d_h1_node,
                                  const int32_t *d_h1_eidx, float *d_walk_imp, float *d_imp0, float *d_imp1, int beta_sample, uint64_t seed, int device,
                                  tm_stream stream) {
    if (B < 0 || W <= 0 || K0 < 0 || K1 < 0 ||
        (B > 0 && (!d_scores || !d_eidx || (K0 > 0 && (!d_h0_node || !d_h0_eidx || !d_imp0)) || (K1 > 0 && (!d_h1_node || !d_h1_eidx || !d_imp1))))) {
        set_error("tm_edge_importance: bad argument");
        return TM_ERR_ARG;
    }
    if (d_gate_blob && (!desc || !d_t || !d_edge_feat || !d_walk_imp)) { set_error("tm_edge_importance: the gate needs desc, d_t, d_edge_feat and d_walk_imp"); return TM_ERR_ARG; }
    if (B == 0 || K0 + K1 == 0) return TM_OK;
    int slots = 64;
    while (slots < 6 * W) slots <<= 1;                       // load factor <= 1/2
    const size_t smem = sizeof(int32_t) * 2 * (size_t)slots;
    if (smem > 48 * 1024) { set_error("tm_edge_importance: %lld walks per root exceed the shared-memory window", (long long)W); return TM_ERR_UNSUPPORTED; }
    TM_DEVICE(device);
    cudaStream_t st = (cudaStream_t)stream;
    if (d_gate_blob) {
        const int rc = tc_gate_launch(*desc, d_gate_blob, B * W * 3, d_eidx, d_t, d_scores, d_edge_feat, n_edge_rows, d_walk_imp, device, st);
        if (rc != TM_OK) return rc;
    }
    const int threads = (int)std::min<int64_t>(256, std::max<int64_t>(64, ((K0 + K1 + 31) / 32) * 32));
    edge_imp_kernel<<<(unsigned)B, threads, smem, st>>>(B, (int)(3 * W), slots - 1, d_eidx, d_gate_blob ? d_walk_imp : nullptr, d_scores, (int)K0, d_h0_node,
                                                       d_h0_eidx, (int)K1, d_h1_node, d_h1_eidx, d_imp0, d_imp1, beta_sample, seed);
    TM_LAUNCH_CHECK();
    return TM_OK;
}
