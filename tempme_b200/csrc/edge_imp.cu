// Motif -> edge explanation aggregation: TempME.retrieve_edge_imp_node in eval mode (reference models/explainer.py:354-406).
// The dependency gate runs on the tensor cores (gate_tc_kernel, encoder_tc.cu); this file holds the per-root segmented max /
// gather / Beta mean / padding mask and the C ABI entry points.
#include <algorithm>

#include "common.cuh"

namespace tmb {

// One block per root: the 3W (edge id, importance) pairs of the root's walks in shared memory, one thread per hop slot.
// scatter(reduce="max", dim_size=num_edges) of torch_scatter leaves ids that no walk carries at 0 (:389).
__global__ void edge_imp_kernel(int64_t B, int W3, const int32_t *__restrict__ w_eidx, const float *__restrict__ walk_imp,
                                const float *__restrict__ scores, int K0, const int32_t *__restrict__ h0_node,
                                const int32_t *__restrict__ h0_eidx, int K1, const int32_t *__restrict__ h1_node,
                                const int32_t *__restrict__ h1_eidx, float *__restrict__ imp0, float *__restrict__ imp1) {
    extern __shared__ int32_t sh[];             // ids [W3], then importances [W3]
    int32_t *ids = sh;
    float *val = reinterpret_cast<float *>(sh + W3);
    const int64_t b = blockIdx.x;
    for (int i = threadIdx.x; i < W3; i += blockDim.x) {
        ids[i] = w_eidx[b * W3 + i];
        val[i] = walk_imp ? walk_imp[b * W3 + i] : scores[b * (W3 / 3) + i / 3];      // graphlet_imp.repeat(1,1,3) (:363)
    }
    __syncthreads();
    for (int s = threadIdx.x; s < K0 + K1; s += blockDim.x) {
        const bool l0 = s < K0;
        const int64_t o = l0 ? b * K0 + s : b * K1 + (s - K0);
        const int32_t id = l0 ? h0_eidx[o] : h1_eidx[o], node = l0 ? h0_node[o] : h1_node[o];
        float m = 0.f;
        for (int i = 0; i < W3; ++i) m = ids[i] == id ? fmaxf(m, val[i]) : m;          // every thread reads the same word: broadcast
        const float alpha = fmaxf(__fmul_rn(m, 10.f), 1.f), beta = fmaxf(__fmul_rn(__fsub_rn(1.f, m), 10.f), 1.f);   // :423-424
        const float out = node == 0 ? 0.f : __fdiv_rn(alpha, __fadd_rn(alpha, beta));                                   // :429, :400-404
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
    if (desc->hid_dim != 64) { set_error("tm_gate_pack: hid_dim %d unsupported (only 64, the reference default)", desc->hid_dim); return TM_ERR_UNSUPPORTED; }
    return tc_gate_pack(*desc, *p, h_blob);
}

extern "C" int tm_edge_importance(const tm_gate_desc *desc, const float *d_gate_blob, int64_t B, int64_t W, const float *d_scores,
                                  const int32_t *d_eidx, const float *d_t, const float *d_edge_feat, int64_t n_edge_rows,
                                  int64_t K0, const int32_t *d_h0_node, const int32_t *d_h0_eidx, int64_t K1, const int32_t *d_h1_node,
                                  const int32_t *d_h1_eidx, float *d_walk_imp, float *d_imp0, float *d_imp1, int device, tm_stream stream) {
    if (B < 0 || W <= 0 || K0 < 0 || K1 < 0 ||
        (B > 0 && (!d_scores || !d_eidx || (K0 > 0 && (!d_h0_node || !d_h0_eidx || !d_imp0)) || (K1 > 0 && (!d_h1_node || !d_h1_eidx || !d_imp1))))) {
        set_error("tm_edge_importance: bad argument");
        return TM_ERR_ARG;
    }
    if (d_gate_blob && (!desc || !d_t || !d_edge_feat || !d_walk_imp)) { set_error("tm_edge_importance: the gate needs desc, d_t, d_edge_feat and d_walk_imp"); return TM_ERR_ARG; }
    if (B == 0 || K0 + K1 == 0) return TM_OK;
    const size_t smem = sizeof(int32_t) * 6 * (size_t)W;
    if (smem > 48 * 1024) { set_error("tm_edge_importance: %lld walks per root exceed the shared-memory window", (long long)W); return TM_ERR_UNSUPPORTED; }
    TM_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    if (d_gate_blob) {
        const int rc = tc_gate_launch(*desc, d_gate_blob, B * W * 3, d_eidx, d_t, d_scores, d_edge_feat, n_edge_rows, d_walk_imp, device, st);
        if (rc != TM_OK) return rc;
    }
    const int threads = (int)std::min<int64_t>(1024, std::max<int64_t>(64, ((K0 + K1 + 31) / 32) * 32));
    edge_imp_kernel<<<(unsigned)B, threads, smem, st>>>(B, (int)(3 * W), d_eidx, d_gate_blob ? d_walk_imp : nullptr, d_scores, (int)K0, d_h0_node,
                                                       d_h0_eidx, (int)K1, d_h1_node, d_h1_eidx, d_imp0, d_imp1);
    TM_LAUNCH_CHECK();
    return TM_OK;
}
