// Shared device/host helpers of libtempme_b200 (sm_100a).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/tempme_b200.h"

namespace tmb {

void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define TM_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            tmb::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return TM_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

#define TM_LAUNCH_CHECK()                                                                      \
    do {                                                                                       \
        tmb::g_launches.fetch_add(1, std::memory_order_relaxed);                                \
        cudaError_t e__ = cudaGetLastError();                                                  \
        if (e__ != cudaSuccess) {                                                              \
            tmb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
            return TM_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

// RAII current-device guard of the C entry points: switch to `dev` for the call, restore the caller's device on return
// (a process that drives several GPUs -- or torch's own current device -- is left as it was found).
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        int cur = -1;
        if (cudaGetDevice(&cur) != cudaSuccess) { ok = false; return; }
        if (dev >= 0 && dev != cur) { ok = cudaSetDevice(dev) == cudaSuccess; if (ok) prev = cur; }
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};
// device that owns a device pointer (-1: unknown / host memory): entry points without a device argument run where their data lives
inline int device_of(const void *p) {
    cudaPointerAttributes at;
    if (!p || cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return -1; }
    return (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) ? at.device : -1;
}
#define TM_DEVICE(dev)                                                                         \
    tmb::DeviceGuard guard__(dev);                                                             \
    if (!guard__.ok) { tmb::set_error("cannot select CUDA device %d (%s:%d)", (int)(dev), __FILE__, __LINE__); return TM_ERR_CUDA; }

// tensor-core scorer (encoder_tc.cu)
int64_t tc_blob_floats(const tm_encoder_desc &d);
int tc_pack(const tm_encoder_desc &d, const tm_encoder_params &p, float *blob);
int64_t tc_slab_motifs(int device = -1);
int tc_encode_score(const tm_encoder_desc &d, const float *d_blob_tc, int64_t B, int64_t W, int64_t group, const int32_t *nodes,
                    const int32_t *eidx, const float *t, const uint8_t *cat, const float *cut, const float *eid, const float *node_feat,
                    int64_t n_node_rows, const float *edge_feat, int64_t n_edge_rows, const float *std_, float *F, float *scores,
                    float *y_out, float *const *peer_scores, int n_peers, int device, cudaStream_t st);

int tc_project_edges(const tm_encoder_desc &d, const float *d_blob_tc, const float *edge_feat, int64_t n_edge_rows, float *P, cudaStream_t st);

// tensor map for tile::gather4 row gathers (encoder_tc.cu)
bool make_gather_map(CUtensorMap *map, const float *table, int64_t rows, int dim, int swizzle128);

// dependency gate of the motif -> edge aggregation (encoder_tc.cu)
int64_t tc_gate_blob_floats(const tm_gate_desc &d);
int tc_gate_pack(const tm_gate_desc &d, const tm_gate_params &p, float *blob);
int tc_gate_launch(const tm_gate_desc &d, const float *d_blob, int64_t n_events, const int32_t *eidx, const float *t, const float *scores,
                   const float *edge_feat, int64_t n_edge_rows, float *out, int device, cudaStream_t st);

// One CSR entry: 16 bytes so that a sampled neighbour costs one 128-bit load (one sector).
struct __align__(16) Entry {
    int32_t nbr;
    int32_t eidx;
    double ts;
};

// nodeedge2idx as a table: one 32-byte row (= one DRAM sector) per edge id with, for each of the (at most two) nodes whose list holds the
// edge, the cut (effective prefix length) AND the start of the node's window, so a lookup by (node, e_idx) needs no access to off[].
struct __align__(32) EdgeSlot {
    int32_t node_a, node_b;   // -1 = absent; node_a < node_b when both are present
    int32_t cut_a, cut_b;
    int64_t start_a, start_b;
};

// One row (= one DRAM sector) of the run directory: the run of the secondary index that belongs to (node, neighbour), with the run's first
// four positions inline -- most (node, neighbour) pairs of a temporal multigraph share one to four events, and for those the id filter of
// step 3 never touches skey.
struct __align__(32) RunSlot {
    uint64_t key;             // node << 32 | neighbour; ~0 = empty
    uint32_t start, len;      // run = skey[start .. start + len)
    uint32_t pos[4];          // positions (inside the node's window) of the run's first min(len, 4) entries, ascending
};

// Device view of the graph (all pointers device memory, immutable after build).
struct GraphView {
    int64_t n_nodes;
    int64_t n_entries;
    int64_t max_eidx;
    const int64_t *off;   // [n_nodes + 1]
    const Entry *entry;   // [n_entries]  time-sorted per node
    const uint64_t *skey; // [n_entries]  per node: (nbr << 32 | position) sorted -- secondary index for the id filter of step 3
    const EdgeSlot *etab; // [max_eidx + 1] nodeedge2idx as a table (see EdgeSlot)
    const RunSlot *htab;  // open-addressing directory of the runs of skey, key = node << 32 | nbr: skey[start .. start+len) = the positions of
                          //   neighbour nbr in node's list (the first four inline); nullptr: search skey instead
    uint64_t hmask;       // slots - 1 (power of two)
};

// 64-bit finaliser (murmur3) for the run directory
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}

}  // namespace tm

struct tm_graph {
    tmb::GraphView v;
    int device;
    int64_t device_bytes;
};

namespace tmb {

// ---------------- Philox4x32-10 (counter-based; DESIGN.md "RNG") ----------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                                        uint32_t c2, uint32_t c3, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ uint64_t draw_index(uint64_t seed, uint32_t stage, uint64_t row, uint32_t slot, uint64_t L) {
    uint32_t o[4];
    philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), slot >> 1, (uint32_t)row, (uint32_t)(row >> 32), stage, o);
    const uint64_t r = (slot & 1) ? ((uint64_t)o[3] << 32 | o[2]) : ((uint64_t)o[1] << 32 | o[0]);
    return __umul64hi(r, L);
}

// d_err protocol: 0 = fine, otherwise 1 + smallest failing row
__device__ __forceinline__ void report_row_error(int32_t *err, int64_t row) {
    if (!err) return;
    const int32_t val = (int32_t)min(row + 1, (int64_t)INT32_MAX);
    int32_t old = *((volatile int32_t *)err);
    while (old == 0 || val < old) {
        const int32_t prev = atomicCAS(err, old, val);
        if (prev == old) break;
        old = prev;
    }
}

__device__ __forceinline__ Entry load_entry(const Entry *p) {
    const int4 v = __ldg(reinterpret_cast<const int4 *>(p));
    Entry e;
    e.nbr = v.x; e.eidx = v.y;
    e.ts = __longlong_as_double(((long long)(uint32_t)v.w << 32) | (uint32_t)v.z);
    return e;
}

__device__ __forceinline__ EdgeSlot load_slot(const EdgeSlot *p) {       // two 128-bit loads of one sector
    const int4 a = __ldg(reinterpret_cast<const int4 *>(p)), b = __ldg(reinterpret_cast<const int4 *>(p) + 1);
    EdgeSlot t;
    t.node_a = a.x; t.node_b = a.y; t.cut_a = a.z; t.cut_b = a.w;
    t.start_a = ((long long)(uint32_t)b.y << 32) | (uint32_t)b.x; t.start_b = ((long long)(uint32_t)b.w << 32) | (uint32_t)b.z;
    return t;
}
__device__ __forceinline__ EdgeSlot empty_slot() { EdgeSlot t; t.node_a = t.node_b = t.cut_a = t.cut_b = -1; t.start_a = t.start_b = 0; return t; }
// the row of edge id e (absent ids give an empty row)
__device__ __forceinline__ EdgeSlot edge_slot(const GraphView &g, int32_t e) {
    return (e >= 0 && (int64_t)e <= g.max_eidx) ? load_slot(g.etab + e) : empty_slot();
}
// nodeedge2idx[node].get(e): -1 when absent (None); *start receives the node's window start when present
__device__ __forceinline__ int64_t dict_get(const GraphView &g, int64_t node, int32_t e, int64_t *start = nullptr) {
    const EdgeSlot t = edge_slot(g, e);
    if (node == t.node_a) { if (start) *start = t.start_a; return t.cut_a; }
    if (node == t.node_b) { if (start) *start = t.start_b; return t.cut_b; }
    return -1;
}

// bisect_left_adapt (utils/graph.py:511-530) as a warp-cooperative 33-ary search on the float64
// timestamps of one node: first index i in [0, len) with ts[i] >= x.  All 32 lanes must call.
__device__ __forceinline__ int64_t warp_lower_bound(const Entry *base, int64_t len, double x, int lane) {
    int64_t lo = 0, hi = len;
    while (hi - lo > 32) {
        const int64_t span = hi - lo;
        const int64_t pos = lo + ((int64_t)(lane + 1) * span) / 33;  // lo < pos < hi, increasing in lane
        const double v = __ldg(&base[pos].ts);
        const unsigned b = __ballot_sync(0xffffffffu, v < x);       // sorted -> lanes [0, cnt) are true
        const int cnt = __popc(b);
        const int64_t plo = lo + ((int64_t)cnt * span) / 33;        // pivot of lane cnt-1
        const int64_t phi = lo + ((int64_t)(cnt + 1) * span) / 33;  // pivot of lane cnt
        if (cnt > 0) lo = plo + 1;
        if (cnt < 32) hi = phi;
    }
    const int64_t pos = lo + lane;
    const bool lt = pos < hi && __ldg(&base[pos].ts) < x;
    return lo + __popc(__ballot_sync(0xffffffffu, lt));
}

}  // namespace tm
