// K1: the CSR build on the device.  Replaces NeighborFinder.__init__ / init_off_set / get_ts2idx (reference utils/graph.py:13-101)
// for the inputs every caller of the reference produces: the flattened adj_list (or the event list) is uploaded once and then
//   1. stably sorted by (node, timestamp): LSD radix sort of the float64 timestamps (order-preserving key transform), then of the node
//      ids -- ties keep insertion (CSV) order, as Python's sorted(key=ts) does (:48);
//   2. window offsets from the run lengths of the sorted node ids + an exclusive scan;
//   3. nodeedge2idx as the per-edge table: the cut of entry i is i, or the first slot of its timestamp run when a later, different
//      timestamp follows in the list (get_ts2idx's tie groups; the trailing group is never flushed, :77-101) -- two binary searches
//      per entry over the node's own window; the (edge, node) slots are claimed with atomicCAS and ordered by node id afterwards;
//   4. the secondary index skey = per node (neighbour << 32 | position) sorted: two more stable radix passes (neighbour, then node).
// Inputs the literal dict emulation treats specially -- the same (node, edge id) pair twice in one list (self-loops), negative
// timestamps (get_ts2idx starts with last_ts = -1) -- are detected on the device and sent to the host pass of graph.cu instead.
#include <cub/cub.cuh>

#include <algorithm>

#include "common.cuh"

namespace tmb {

namespace {

constexpr int kT = 256;
inline unsigned blocks_for(int64_t n) { return (unsigned)std::min<int64_t>((n + kT - 1) / kT, 148 * 64); }

enum { kFlagNodeRange = 1, kFlagNegEdge = 2, kFlagSpecialTs = 4, kFlagDupPair = 8, kFlagThreeNodes = 16 };

__global__ void expand_events_kernel(int64_t m, const int32_t *__restrict__ src, const int32_t *__restrict__ dst, const int32_t *__restrict__ eidx,
                                     const double *__restrict__ ts, int32_t *node, int32_t *nbr, int32_t *e, double *t) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (int64_t)gridDim.x * blockDim.x) {
        node[2 * k] = src[k]; nbr[2 * k] = dst[k]; node[2 * k + 1] = dst[k]; nbr[2 * k + 1] = src[k];   // adj[src].append((dst, e, t)); adj[dst].append((src, e, t))
        e[2 * k] = e[2 * k + 1] = eidx[k];
        t[2 * k] = t[2 * k + 1] = ts[k];
    }
}

// range checks + the sort keys of pass 1: timestamps as order-preserving unsigned keys (-0.0 sorts with +0.0, as `<` sees them)
__global__ void prepare_kernel(int64_t n, int64_t n_nodes, const int32_t *__restrict__ node, const int32_t *__restrict__ eidx, const double *__restrict__ ts,
                               unsigned long long *key, uint32_t *val, int *flags, int *max_e) {
    int f = 0, me = -1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t v = node[i], e = eidx[i];
        if (v < 0 || v >= n_nodes) f |= kFlagNodeRange;
        if (e < 0) f |= kFlagNegEdge;
        me = max(me, e);
        double t = ts[i];
        if (!(t >= 0.0)) f |= kFlagSpecialTs;                      // negative or NaN: the literal emulation on the host decides
        if (t == 0.0) t = 0.0;
        const unsigned long long b = (unsigned long long)__double_as_longlong(t);
        key[i] = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
        val[i] = (uint32_t)i;
    }
    if (f) atomicOr(flags, f);
    atomicMax(max_e, me);
}

__global__ void gather_node_kernel(int64_t n, const uint32_t *__restrict__ perm, const int32_t *__restrict__ node, uint32_t *out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (uint32_t)node[perm[i]];
}

// sorted node ids -> count per node (the thread at a run start finds the run's end by bisection)
__global__ void run_count_kernel(int64_t n, const uint32_t *__restrict__ snode, int64_t *cnt) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t v = snode[i];
        if (i > 0 && snode[i - 1] == v) continue;
        int64_t lo = i + 1, hi = n;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (snode[mid] <= v) lo = mid + 1; else hi = mid; }
        cnt[v] = lo - i;
    }
}

__global__ void build_entries_kernel(int64_t n, const uint32_t *__restrict__ perm, const int32_t *__restrict__ nbr, const int32_t *__restrict__ eidx,
                                     const double *__restrict__ ts, Entry *ent) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t p = perm[i];
        Entry e; e.nbr = nbr[p]; e.eidx = eidx[p]; e.ts = ts[p];
        ent[i] = e;
    }
}

// nodeedge2idx: claim the (edge, node) slot, store the cut (see the header comment, step 3)
__global__ void edge_table_kernel(int64_t n, const uint32_t *__restrict__ snode, const int64_t *__restrict__ off, const Entry *__restrict__ ent, EdgeSlot *etab, int *flags) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int32_t v = (int32_t)snode[p];
        const int64_t s = off[v], len = off[v + 1] - s, i = p - s;
        const double t = ent[p].ts;
        int64_t lo = 0, hi = i;                                   // first slot of the run of equal timestamps
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (ent[s + mid].ts < t) lo = mid + 1; else hi = mid; }
        const int64_t first = lo;
        lo = i + 1; hi = len;                                     // one past its last slot
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (ent[s + mid].ts <= t) lo = mid + 1; else hi = mid; }
        const int32_t cut = (int32_t)(lo < len ? first : i);      // flushed only when a later, different timestamp exists (:93-98)
        EdgeSlot *row = etab + ent[p].eidx;
        int *w = reinterpret_cast<int *>(row);                   // {node_a, node_b, cut_a, cut_b}, then the two window starts
        int prev = atomicCAS(w, -1, v);
        if (prev == -1) { w[2] = cut; row->start_a = s; continue; }
        if (prev == v) { atomicOr(flags, kFlagDupPair); continue; }
        prev = atomicCAS(w + 1, -1, v);
        if (prev == -1) { w[3] = cut; row->start_b = s; continue; }
        atomicOr(flags, prev == v ? kFlagDupPair : kFlagThreeNodes);
    }
}
// the host pass hands slot x to the smaller node id (it walks the nodes in ascending order): same convention here
__global__ void order_slots_kernel(int64_t n, EdgeSlot *etab) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        EdgeSlot t = etab[e];
        if (t.node_a == -1) { t.start_a = 0; t.start_b = 0; etab[e] = t; continue; }      // the 0xff fill also covered the starts
        if (t.node_b == -1) t.start_b = 0;
        else if (t.node_b < t.node_a) t = EdgeSlot{t.node_b, t.node_a, t.cut_b, t.cut_a, t.start_b, t.start_a};
        etab[e] = t;
    }
}

__global__ void nbr_keys_kernel(int64_t n, const Entry *__restrict__ ent, uint32_t *key, uint32_t *val) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) { key[i] = (uint32_t)ent[i].nbr; val[i] = (uint32_t)i; }
}
__global__ void gather_u32_kernel(int64_t n, const uint32_t *__restrict__ perm, const uint32_t *__restrict__ src, uint32_t *out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = src[perm[i]];
}
__global__ void skey_kernel(int64_t n, const uint32_t *__restrict__ perm, const uint32_t *__restrict__ snode, const int64_t *__restrict__ off,
                            const Entry *__restrict__ ent, unsigned long long *skey) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t p = perm[i];                                // slot i of the (node, neighbour, position) order holds entry p
        skey[i] = (unsigned long long)(uint32_t)ent[p].nbr << 32 | (unsigned long long)((int64_t)p - off[snode[p]]);
    }
}

struct Buf {                       // device allocations released on every exit path
    void *p = nullptr;
    ~Buf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, std::max<size_t>(bytes, 16)); }
    template <typename T> T *as() { return static_cast<T *>(p); }
    void *release() { void *q = p; p = nullptr; return q; }
    void free_now() { if (p) { cudaFree(p); p = nullptr; } }
};

#define TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { set_error("device graph build: %s failed: %s", #call, cudaGetErrorString(e__)); \
                                                                            return e__ == cudaErrorMemoryAllocation ? TM_ERR_NOMEM : TM_ERR_CUDA; } } while (0)

template <typename K>
int sort_pairs(K *&k_in, K *&k_out, uint32_t *&v_in, uint32_t *&v_out, int64_t n, int end_bit) {
    cub::DoubleBuffer<K> dk(k_in, k_out);
    cub::DoubleBuffer<uint32_t> dv(v_in, v_out);
    size_t tmp_bytes = 0;
    TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, n, 0, end_bit));
    Buf tmp;
    TRY(tmp.alloc(tmp_bytes));
    TRY(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, dk, dv, n, 0, end_bit));
    TRY(cudaDeviceSynchronize());
    k_in = dk.Current(); k_out = dk.Alternate(); v_in = dv.Current(); v_out = dv.Alternate();
    return TM_OK;
}

int bits_for(int64_t max_value) { int b = 1; while (b < 32 && (int64_t(1) << b) <= max_value) ++b; return b; }

}  // namespace

// Device build from entries already resident on the device (d_node / d_nbr / d_eidx / d_ts, n entries).  Returns TM_OK and fills *view /
// *device_bytes, or 1 when the input needs the host pass (see the header comment), or a negative tm_status.  Takes ownership of nothing.
int device_graph_build(int64_t n_nodes, int64_t n, const int32_t *d_node, const int32_t *d_nbr, const int32_t *d_eidx, const double *d_ts,
                       GraphView *view, int64_t *device_bytes) {
    Buf key64a, key64b, vala, valb, flags, cnt;
    TRY(key64a.alloc(sizeof(unsigned long long) * n)); TRY(key64b.alloc(sizeof(unsigned long long) * n));
    TRY(vala.alloc(sizeof(uint32_t) * n)); TRY(valb.alloc(sizeof(uint32_t) * n));
    TRY(flags.alloc(2 * sizeof(int)));
    const int init[2] = {0, -1};
    TRY(cudaMemcpy(flags.p, init, sizeof init, cudaMemcpyHostToDevice));
    prepare_kernel<<<blocks_for(n), kT>>>(n, n_nodes, d_node, d_eidx, d_ts, key64a.as<unsigned long long>(), vala.as<uint32_t>(), flags.as<int>(), flags.as<int>() + 1);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    int h[2];
    TRY(cudaMemcpy(h, flags.p, sizeof h, cudaMemcpyDeviceToHost));
    if (h[0] & kFlagNodeRange) { set_error("an entry's node id is outside [0, %lld)", (long long)n_nodes); return TM_ERR_NODE_RANGE; }
    if (h[0] & kFlagNegEdge) { set_error("an entry has a negative edge id"); return TM_ERR_EDGE_TABLE; }
    if (h[0] & kFlagSpecialTs) return 1;
    const int64_t max_e = h[1];
    // ---- 1. stable sort by (node, ts): timestamps first, then node ids
    unsigned long long *k64i = key64a.as<unsigned long long>(), *k64o = key64b.as<unsigned long long>();
    uint32_t *vi = vala.as<uint32_t>(), *vo = valb.as<uint32_t>();
    int rc = sort_pairs(k64i, k64o, vi, vo, n, 64);
    if (rc != TM_OK) return rc;
    uint32_t *k32i = reinterpret_cast<uint32_t *>(k64o), *k32o = k32i + n;            // the idle 64-bit key buffer holds both 32-bit key buffers
    gather_node_kernel<<<blocks_for(n), kT>>>(n, vi, d_node, k32i);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    rc = sort_pairs(k32i, k32o, vi, vo, n, bits_for(n_nodes));
    if (rc != TM_OK) return rc;
    const uint32_t *snode = k32i, *perm = vi;                                          // sorted node ids; perm[i] = input index of CSR slot i
    // ---- 2. offsets
    Buf off, ent, etab, skey;
    TRY(cnt.alloc(sizeof(int64_t) * (n_nodes + 1)));
    TRY(cudaMemset(cnt.p, 0, sizeof(int64_t) * (n_nodes + 1)));
    run_count_kernel<<<blocks_for(n), kT>>>(n, snode, cnt.as<int64_t>());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    TRY(off.alloc(sizeof(int64_t) * (n_nodes + 1)));
    {
        size_t tmp_bytes = 0;
        TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt.as<int64_t>(), off.as<int64_t>(), n_nodes + 1));
        Buf tmp;
        TRY(tmp.alloc(tmp_bytes));
        TRY(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, cnt.as<int64_t>(), off.as<int64_t>(), n_nodes + 1));
    }
    cnt.free_now();
    TRY(ent.alloc(sizeof(Entry) * n));
    build_entries_kernel<<<blocks_for(n), kT>>>(n, perm, d_nbr, d_eidx, d_ts, ent.as<Entry>());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    // ---- 3. nodeedge2idx
    TRY(etab.alloc(sizeof(EdgeSlot) * (max_e + 1)));
    TRY(cudaMemset(etab.p, 0xff, sizeof(EdgeSlot) * std::max<int64_t>(max_e + 1, 1)));
    edge_table_kernel<<<blocks_for(n), kT>>>(n, snode, off.as<int64_t>(), ent.as<Entry>(), etab.as<EdgeSlot>(), flags.as<int>());
    order_slots_kernel<<<blocks_for(max_e + 1), kT>>>(max_e + 1, etab.as<EdgeSlot>());
    g_launches.fetch_add(2, std::memory_order_relaxed);
    TRY(cudaMemcpy(h, flags.p, sizeof h, cudaMemcpyDeviceToHost));
    if (h[0] & kFlagThreeNodes) { set_error("an edge id occurs in the lists of more than two nodes"); return TM_ERR_EDGE_TABLE; }
    if (h[0] & kFlagDupPair) return 1;
    // ---- 4. secondary index: (node, neighbour, position) order by two more stable passes
    TRY(skey.alloc(sizeof(unsigned long long) * n));
    {
        Buf snode_keep, nk_a, nk_b, pv_a, pv_b;
        TRY(snode_keep.alloc(sizeof(uint32_t) * n));
        TRY(cudaMemcpy(snode_keep.p, snode, sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice));
        key64a.free_now(); key64b.free_now(); vala.free_now(); valb.free_now();         // (snode and perm lived in them)
        TRY(nk_a.alloc(sizeof(uint32_t) * n)); TRY(nk_b.alloc(sizeof(uint32_t) * n)); TRY(pv_a.alloc(sizeof(uint32_t) * n)); TRY(pv_b.alloc(sizeof(uint32_t) * n));
        uint32_t *ki = nk_a.as<uint32_t>(), *ko = nk_b.as<uint32_t>(), *pi = pv_a.as<uint32_t>(), *po = pv_b.as<uint32_t>();
        nbr_keys_kernel<<<blocks_for(n), kT>>>(n, ent.as<Entry>(), ki, pi);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        rc = sort_pairs(ki, ko, pi, po, n, bits_for(n_nodes));                          // neighbour ids are node ids
        if (rc != TM_OK) return rc;
        gather_u32_kernel<<<blocks_for(n), kT>>>(n, pi, snode_keep.as<uint32_t>(), ki);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        rc = sort_pairs(ki, ko, pi, po, n, bits_for(n_nodes));
        if (rc != TM_OK) return rc;
        skey_kernel<<<blocks_for(n), kT>>>(n, pi, snode_keep.as<uint32_t>(), off.as<int64_t>(), ent.as<Entry>(), skey.as<unsigned long long>());
        g_launches.fetch_add(1, std::memory_order_relaxed);
        TRY(cudaDeviceSynchronize());
    }
    TRY(cudaGetLastError());
    view->n_nodes = n_nodes; view->n_entries = n; view->max_eidx = max_e;
    *device_bytes = (int64_t)(sizeof(int64_t) * (n_nodes + 1) + sizeof(Entry) * n + sizeof(unsigned long long) * n + sizeof(EdgeSlot) * (max_e + 1));
    view->off = static_cast<const int64_t *>(off.release()); view->entry = static_cast<const Entry *>(ent.release());
    view->skey = static_cast<const uint64_t *>(skey.release()); view->etab = static_cast<const EdgeSlot *>(etab.release());
    view->htab = nullptr; view->hmask = 0;
    return TM_OK;
}

// Event list on the host -> entries on the device (the callers' loop, temp_exp_main.py:135-144) -> device build.
int device_graph_build_from_events(int64_t n_nodes, int64_t m, const int32_t *h_src, const int32_t *h_dst, const int32_t *h_eidx, const double *h_ts,
                                   GraphView *view, int64_t *device_bytes) {
    Buf src, dst, e, t, node, nbr, e2, t2;
    TRY(src.alloc(4 * m)); TRY(dst.alloc(4 * m)); TRY(e.alloc(4 * m)); TRY(t.alloc(8 * m));
    TRY(cudaMemcpy(src.p, h_src, 4 * m, cudaMemcpyHostToDevice)); TRY(cudaMemcpy(dst.p, h_dst, 4 * m, cudaMemcpyHostToDevice));
    TRY(cudaMemcpy(e.p, h_eidx, 4 * m, cudaMemcpyHostToDevice)); TRY(cudaMemcpy(t.p, h_ts, 8 * m, cudaMemcpyHostToDevice));
    TRY(node.alloc(8 * m)); TRY(nbr.alloc(8 * m)); TRY(e2.alloc(8 * m)); TRY(t2.alloc(16 * m));
    expand_events_kernel<<<blocks_for(m), kT>>>(m, src.as<int32_t>(), dst.as<int32_t>(), e.as<int32_t>(), t.as<double>(), node.as<int32_t>(), nbr.as<int32_t>(),
                                                e2.as<int32_t>(), t2.as<double>());
    g_launches.fetch_add(1, std::memory_order_relaxed);
    TRY(cudaDeviceSynchronize());
    src.free_now(); dst.free_now(); e.free_now(); t.free_now();
    return device_graph_build(n_nodes, 2 * m, node.as<int32_t>(), nbr.as<int32_t>(), e2.as<int32_t>(), t2.as<double>(), view, device_bytes);
}

}  // namespace tmb
