// Graph build + find_before.  Replaces NeighborFinder.__init__/init_off_set/get_ts2idx/find_before
// (reference utils/graph.py:13-146).  The build is a one-time host pass (counting sort by node,
// stable per-node sort by timestamp, literal emulation of get_ts2idx into a per-edge table),
// followed by one upload; every query after that runs on the device.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace tmb {

static thread_local std::string t_err;
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_err = buf;
}

// find_before, utils/graph.py:103-146.  One warp per row so that the time cut is the
// warp-cooperative search; e_idx rows cost one 16-byte table load.
__global__ void find_before_kernel(GraphView g, int64_t R, const int32_t *__restrict__ node,
                                   const double *__restrict__ cut_time, const int32_t *__restrict__ eidx,
                                   int64_t *__restrict__ o_start, int32_t *__restrict__ o_cut, int32_t *err) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= R) return;
    const int64_t v = node[row];
    if (v < 0 || v >= g.n_nodes) {
        if (lane == 0) { o_start[row] = 0; o_cut[row] = 0; report_row_error(err, row); }
        return;
    }
    const int32_t e = eidx ? eidx[row] : TM_EIDX_NONE;
    int64_t c, s = -1;
    if (e == TM_EIDX_NONE) {
        s = __ldg(g.off + v);
        c = warp_lower_bound(g.entry + s, __ldg(g.off + v + 1) - s, cut_time ? cut_time[row] : 0.0, lane);   // graph.py:129
    } else if (v > 0) {                                                                     // graph.py:133
        c = dict_get(g, v, e, &s);
        if (c < 0) { c = 0; if (lane == 0) report_row_error(err, row); }   // IndexError, graph.py:134-135
    } else c = 0;
    if (s < 0) s = __ldg(g.off + v);
    if (lane == 0) { o_start[row] = s; o_cut[row] = (int32_t)c; }
}

// Directory of the neighbour runs of skey.  One THREAD per CSR entry (hub windows of millions of entries would serialise a warp-per-node
// walk): the thread finds its node by bisection over off[] (L2 resident), and if its entry starts a run of one neighbour id it finds the
// run's end by bisection and inserts the run.  count != nullptr: only count the runs.
__global__ void run_directory_kernel(GraphView g, RunSlot *tab, uint64_t mask, unsigned long long *count) {
    unsigned long long local = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < g.n_entries; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t nb = (uint32_t)(g.skey[i] >> 32);
        int64_t lo = 0, hi = g.n_nodes;                      // node v with off[v] <= i < off[v + 1]
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (__ldg(g.off + mid + 1) <= i) lo = mid + 1; else hi = mid; }
        const int64_t v = lo, s = __ldg(g.off + v), e = __ldg(g.off + v + 1);
        if (i > s && (uint32_t)(g.skey[i - 1] >> 32) == nb) continue;
        if (count) { ++local; continue; }
        lo = i + 1; hi = e;                                  // first index whose neighbour id is larger
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if ((uint32_t)(g.skey[mid] >> 32) <= nb) lo = mid + 1; else hi = mid; }
        const uint64_t key = (uint64_t)v << 32 | nb;
        uint64_t slot = mix64(key) & mask;
        while (atomicCAS(reinterpret_cast<unsigned long long *>(&tab[slot].key), ~0ull, (unsigned long long)key) != ~0ull) slot = (slot + 1) & mask;
        RunSlot &r = tab[slot];
        r.start = (uint32_t)i; r.len = (uint32_t)(lo - i);
        for (int j = 0; j < 4; ++j) r.pos[j] = i + j < lo ? (uint32_t)g.skey[i + j] : 0xffffffffu;
    }
    if (count) {
        for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
    }
}

}  // namespace tm

using namespace tmb;

extern "C" int tm_version(void) { return 100; }
extern "C" const char *tm_last_error(void) { return tmb::t_err.c_str(); }
extern "C" uint64_t tm_launch_count(void) { return tmb::g_launches.load(); }


static void free_graph(tm_graph *g) {
    cudaFree((void *)g->v.off); cudaFree((void *)g->v.entry); cudaFree((void *)g->v.skey); cudaFree((void *)g->v.etab); cudaFree((void *)g->v.htab);
    delete g;
}

// run directory (node, neighbour) -> run of skey, built on the device; load factor <= 1/2
static int build_run_directory(tm_graph *g) {
    const int64_t n_entries = g->v.n_entries;
    cudaError_t ce = cudaSuccess;
    if (n_entries > 0 && !getenv("TEMPME_NO_RUN_DIRECTORY")) {
        unsigned long long *d_cnt = nullptr, runs = 0;
        RunSlot *d_h = nullptr;
        ce = cudaMalloc(&d_cnt, sizeof *d_cnt);
        if (ce == cudaSuccess) ce = cudaMemset(d_cnt, 0, sizeof *d_cnt);
        if (ce == cudaSuccess) {
            run_directory_kernel<<<148 * 16, 256>>>(g->v, nullptr, 0, d_cnt);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            ce = cudaMemcpy(&runs, d_cnt, sizeof runs, cudaMemcpyDeviceToHost);
        }
        uint64_t slots = 1024;
        while (slots < 2 * runs) slots <<= 1;
        if (ce == cudaSuccess) ce = cudaMalloc(&d_h, sizeof(RunSlot) * slots);
        if (ce == cudaSuccess) ce = cudaMemset(d_h, 0xff, sizeof(RunSlot) * slots);
        if (ce == cudaSuccess) {
            run_directory_kernel<<<148 * 16, 256>>>(g->v, d_h, slots - 1, nullptr);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            ce = cudaDeviceSynchronize();
        }
        cudaFree(d_cnt);
        if (ce != cudaSuccess) {
            set_error("run directory build failed: %s", cudaGetErrorString(ce));
            cudaFree(d_h);
            return ce == cudaErrorMemoryAllocation ? TM_ERR_NOMEM : TM_ERR_CUDA;
        }
        g->v.htab = d_h; g->v.hmask = slots - 1;
        g->device_bytes += (int64_t)(sizeof(RunSlot) * slots);
    }
    return TM_OK;
}

// device build (graph_build.cu): TM_OK, 1 = this input needs the literal host pass, < 0 = error
namespace tmb {
int device_graph_build(int64_t n_nodes, int64_t n, const int32_t *d_node, const int32_t *d_nbr, const int32_t *d_eidx, const double *d_ts,
                       GraphView *view, int64_t *device_bytes);
int device_graph_build_from_events(int64_t n_nodes, int64_t m, const int32_t *h_src, const int32_t *h_dst, const int32_t *h_eidx, const double *h_ts,
                                   GraphView *view, int64_t *device_bytes);
}

// L2 fetch granularity: the lookups of this library are random 16-32 byte reads (one sector); with the default 64-byte granularity every
// miss fetches a neighbouring sector it never uses.  TEMPME_L2_FETCH=32|64|128 sets cudaLimitMaxL2FetchGranularity when a graph is created.
static void apply_l2_fetch_hint() {
    const char *e = getenv("TEMPME_L2_FETCH");
    if (!e) return;
    const int v = atoi(e);
    if (v == 32 || v == 64 || v == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)v);
}

static bool use_device_build() { const char *e = getenv("TEMPME_GRAPH_BUILD"); return !(e && strcmp(e, "host") == 0); }

static int finish_device_graph(int rc, tm_graph *g, tm_graph **out) {
    if (rc != TM_OK) { delete g; return rc; }
    apply_l2_fetch_hint();
    const int rd = build_run_directory(g);
    if (rd != TM_OK) { free_graph(g); return rd; }
    *out = g;
    return TM_OK;
}

static inline int32_t slice_len(int64_t c, int64_t len) {  // python a[:c] on a list of length len
    if (c < 0) { c += len; if (c < 0) c = 0; }
    if (c > len) c = len;
    return (int32_t)c;
}

extern "C" int tm_graph_create(int64_t n_nodes, int64_t n_entries, const int32_t *h_node, const int32_t *h_nbr,
                               const int32_t *h_eidx, const double *h_ts, int device, tm_graph **out) {
    if (!out || n_nodes < 0 || n_entries < 0 || (n_entries > 0 && (!h_node || !h_nbr || !h_eidx || !h_ts))) {
        set_error("tm_graph_create: bad argument");
        return TM_ERR_ARG;
    }
    if (n_entries >= (int64_t)INT32_MAX) { set_error("tm_graph_create: more than 2^31-1 entries"); return TM_ERR_UNSUPPORTED; }
    {   // argument errors are reported from the host arrays, before any device work: first bad entry, same text on either build path
        int64_t bad_node = n_entries, bad_edge = n_entries;
#pragma omp parallel for schedule(static) reduction(min : bad_node, bad_edge)
        for (int64_t j = 0; j < n_entries; ++j) {
            if (h_node[j] < 0 || h_node[j] >= n_nodes) bad_node = std::min(bad_node, j);
            if (h_eidx[j] < 0) bad_edge = std::min(bad_edge, j);
        }
        if (bad_node < n_entries) { set_error("entry %lld: node %d outside [0, %lld)", (long long)bad_node, h_node[bad_node], (long long)n_nodes); return TM_ERR_NODE_RANGE; }
        if (bad_edge < n_entries) { set_error("entry %lld: negative edge id %d", (long long)bad_edge, h_eidx[bad_edge]); return TM_ERR_EDGE_TABLE; }
    }
    if (use_device_build() && n_entries > 0) {
        // K1 on the device: upload the entries, sort / scan / table build there (graph_build.cu).  Inputs that the literal get_ts2idx emulation
        // treats specially (a (node, edge id) pair twice in one list, negative timestamps) come back with rc 1 and take the host pass below.
        TM_DEVICE(device);
        void *dn = nullptr, *db = nullptr, *de = nullptr, *dt = nullptr;
        cudaError_t ce = cudaMalloc(&dn, 4 * n_entries);
        if (ce == cudaSuccess) ce = cudaMalloc(&db, 4 * n_entries);
        if (ce == cudaSuccess) ce = cudaMalloc(&de, 4 * n_entries);
        if (ce == cudaSuccess) ce = cudaMalloc(&dt, 8 * n_entries);
        if (ce == cudaSuccess) ce = cudaMemcpy(dn, h_node, 4 * n_entries, cudaMemcpyHostToDevice);
        if (ce == cudaSuccess) ce = cudaMemcpy(db, h_nbr, 4 * n_entries, cudaMemcpyHostToDevice);
        if (ce == cudaSuccess) ce = cudaMemcpy(de, h_eidx, 4 * n_entries, cudaMemcpyHostToDevice);
        if (ce == cudaSuccess) ce = cudaMemcpy(dt, h_ts, 8 * n_entries, cudaMemcpyHostToDevice);
        int rc = TM_ERR_CUDA;
        tm_graph *g = new tm_graph();
        g->device = device; g->device_bytes = 0;
        memset(&g->v, 0, sizeof g->v);
        if (ce == cudaSuccess) rc = device_graph_build(n_nodes, n_entries, (const int32_t *)dn, (const int32_t *)db, (const int32_t *)de, (const double *)dt, &g->v, &g->device_bytes);
        else set_error("graph upload failed: %s", cudaGetErrorString(ce));
        cudaFree(dn); cudaFree(db); cudaFree(de); cudaFree(dt);
        if (rc != 1) return finish_device_graph(rc, g, out);
        delete g;
    }
    std::vector<int64_t> off(n_nodes + 1, 0);
    int64_t max_e = -1;
    for (int64_t j = 0; j < n_entries; ++j) {
        const int32_t v = h_node[j];
        if (v < 0 || v >= n_nodes) { set_error("entry %lld: node %d outside [0, %lld)", (long long)j, v, (long long)n_nodes); return TM_ERR_NODE_RANGE; }
        if (h_eidx[j] < 0) { set_error("entry %lld: negative edge id %d", (long long)j, h_eidx[j]); return TM_ERR_EDGE_TABLE; }
        max_e = std::max<int64_t>(max_e, h_eidx[j]);
        off[v + 1]++;
    }
    for (int64_t v = 0; v < n_nodes; ++v) off[v + 1] += off[v];
    // stable counting sort by node (keeps insertion order inside each list)
    std::vector<Entry> ent(n_entries);
    {
        std::vector<int64_t> fill(off.begin(), off.end() - 1);
        for (int64_t j = 0; j < n_entries; ++j) {
            Entry &e = ent[fill[h_node[j]]++];
            e.nbr = h_nbr[j]; e.eidx = h_eidx[j]; e.ts = h_ts[j];
        }
    }
    // sorted(curr, key=lambda x: x[2]) -- stable, graph.py:48
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t v = 0; v < n_nodes; ++v) {
        Entry *b = ent.data() + off[v], *e = ent.data() + off[v + 1];
        auto lt = [](const Entry &x, const Entry &y) { return x.ts < y.ts; };
        if (!std::is_sorted(b, e, lt)) std::stable_sort(b, e, lt);
    }
    // nodeedge2idx as a table: claim the (edge, node) slots (sequential: two nodes share one row)
    std::vector<EdgeSlot> etab(max_e + 1, EdgeSlot{-1, -1, -1, -1, 0, 0});
    for (int64_t v = 0; v < n_nodes; ++v)
        for (int64_t p = off[v]; p < off[v + 1]; ++p) {
            EdgeSlot &t = etab[ent[p].eidx];
            if (t.node_a == (int32_t)v || t.node_b == (int32_t)v) continue;
            if (t.node_a == -1) { t.node_a = (int32_t)v; t.start_a = off[v]; }
            else if (t.node_b == -1) { t.node_b = (int32_t)v; t.start_b = off[v]; }
            else { set_error("edge id %d occurs in the lists of more than two nodes", ent[p].eidx); return TM_ERR_EDGE_TABLE; }
        }
    // get_ts2idx, graph.py:77-101, emulated literally; each node only touches its own table words
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t v = 0; v < n_nodes; ++v) {
        const int64_t s = off[v], len = off[v + 1] - s;
        auto slot = [&](int32_t e) -> int32_t & { EdgeSlot &t = etab[e]; return t.node_a == (int32_t)v ? t.cut_a : t.cut_b; };
        int64_t tie_lo = -1, tie_n = 0;
        double last_ts = -1.0;                                                    // :82
        for (int64_t i = 0; i < len; ++i) {
            const double t = ent[s + i].ts;
            slot(ent[s + i].eidx) = (int32_t)i;                                   // :85
            if (t == last_ts) { if (tie_n == 0) { tie_lo = i - 1; tie_n = 2; } else tie_n++; }   // :87-91
            if (!(t == last_ts) && tie_n > 0) {                                   // :93-98
                for (int64_t j = 0; j < tie_n; ++j) slot(ent[s + tie_lo + j].eidx) -= (int32_t)j;
                tie_n = 0;
            }
            last_ts = t;
        }
        for (int64_t i = 0; i < len; ++i) { int32_t &c = slot(ent[s + i].eidx); c = slice_len(c, len); }  // [:cut] semantics
    }
    // secondary index: per node the keys (nbr << 32 | position) sorted, i.e. positions grouped by neighbour id
    std::vector<uint64_t> skey(n_entries);
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t v = 0; v < n_nodes; ++v) {
        const int64_t s = off[v], len = off[v + 1] - s;
        for (int64_t i = 0; i < len; ++i) skey[s + i] = (uint64_t)(uint32_t)ent[s + i].nbr << 32 | (uint64_t)i;
        std::sort(skey.begin() + s, skey.begin() + s + len);
    }

    TM_DEVICE(device);
    tm_graph *g = new tm_graph();
    g->device = device;
    void *d_off = nullptr, *d_ent = nullptr, *d_nbr = nullptr, *d_tab = nullptr;
    const size_t b_off = sizeof(int64_t) * (n_nodes + 1), b_ent = sizeof(Entry) * std::max<int64_t>(n_entries, 1),
                 b_nbr = sizeof(uint64_t) * std::max<int64_t>(n_entries, 1), b_tab = sizeof(EdgeSlot) * std::max<int64_t>(max_e + 1, 1);
    cudaError_t ce = cudaMalloc(&d_off, b_off);
    if (ce == cudaSuccess) ce = cudaMalloc(&d_ent, b_ent);
    if (ce == cudaSuccess) ce = cudaMalloc(&d_nbr, b_nbr);
    if (ce == cudaSuccess) ce = cudaMalloc(&d_tab, b_tab);
    if (ce == cudaSuccess) ce = cudaMemcpy(d_off, off.data(), b_off, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess && n_entries) ce = cudaMemcpy(d_ent, ent.data(), sizeof(Entry) * n_entries, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess && n_entries) ce = cudaMemcpy(d_nbr, skey.data(), sizeof(uint64_t) * n_entries, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess && max_e >= 0) ce = cudaMemcpy(d_tab, etab.data(), sizeof(EdgeSlot) * (max_e + 1), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) {
        set_error("graph upload failed: %s", cudaGetErrorString(ce));
        cudaFree(d_off); cudaFree(d_ent); cudaFree(d_nbr); cudaFree(d_tab);
        delete g;
        return ce == cudaErrorMemoryAllocation ? TM_ERR_NOMEM : TM_ERR_CUDA;
    }
    g->v.n_nodes = n_nodes; g->v.n_entries = n_entries; g->v.max_eidx = max_e;
    g->v.off = (const int64_t *)d_off; g->v.entry = (const Entry *)d_ent; g->v.skey = (const uint64_t *)d_nbr; g->v.etab = (const EdgeSlot *)d_tab;
    g->v.htab = nullptr; g->v.hmask = 0;
    g->device_bytes = (int64_t)(b_off + b_ent + b_nbr + b_tab);
    const int rc_dir = build_run_directory(g);
    if (rc_dir != TM_OK) { free_graph(g); return rc_dir; }
    *out = g;
    return TM_OK;
}

extern "C" int tm_graph_create_from_events(int64_t n_nodes, int64_t n_events, const int32_t *h_src, const int32_t *h_dst,
                                           const int32_t *h_eidx, const double *h_ts, int device, tm_graph **out) {
    if (n_events < 0 || (n_events > 0 && (!h_src || !h_dst || !h_eidx || !h_ts))) { set_error("tm_graph_create_from_events: bad argument"); return TM_ERR_ARG; }
    if (!out || n_nodes < 0) { set_error("tm_graph_create_from_events: bad argument"); return TM_ERR_ARG; }
    if (2 * n_events >= (int64_t)INT32_MAX) { set_error("tm_graph_create_from_events: more than 2^31-1 entries"); return TM_ERR_UNSUPPORTED; }
    {
        int64_t bad_node = n_events, bad_edge = n_events;
#pragma omp parallel for schedule(static) reduction(min : bad_node, bad_edge)
        for (int64_t k = 0; k < n_events; ++k) {
            if (h_src[k] < 0 || h_src[k] >= n_nodes || h_dst[k] < 0 || h_dst[k] >= n_nodes) bad_node = std::min(bad_node, k);
            if (h_eidx[k] < 0) bad_edge = std::min(bad_edge, k);
        }
        if (bad_node < n_events) { set_error("event %lld: endpoint (%d, %d) outside [0, %lld)", (long long)bad_node, h_src[bad_node], h_dst[bad_node], (long long)n_nodes); return TM_ERR_NODE_RANGE; }
        if (bad_edge < n_events) { set_error("event %lld: negative edge id %d", (long long)bad_edge, h_eidx[bad_edge]); return TM_ERR_EDGE_TABLE; }
    }
    if (use_device_build() && n_events > 0) {
        TM_DEVICE(device);
        tm_graph *g = new tm_graph();
        g->device = device; g->device_bytes = 0;
        memset(&g->v, 0, sizeof g->v);
        const int rc = device_graph_build_from_events(n_nodes, n_events, h_src, h_dst, h_eidx, h_ts, &g->v, &g->device_bytes);
        if (rc != 1) return finish_device_graph(rc, g, out);
        delete g;
    }
    std::vector<int32_t> node(2 * n_events), nbr(2 * n_events), e(2 * n_events);
    std::vector<double> t(2 * n_events);
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n_events; ++k) {   // adj[src].append((dst,e,t)); adj[dst].append((src,e,t))
        node[2 * k] = h_src[k]; nbr[2 * k] = h_dst[k];
        node[2 * k + 1] = h_dst[k]; nbr[2 * k + 1] = h_src[k];
        e[2 * k] = e[2 * k + 1] = h_eidx[k];
        t[2 * k] = t[2 * k + 1] = h_ts[k];
    }
    return tm_graph_create(n_nodes, 2 * n_events, node.data(), nbr.data(), e.data(), t.data(), device, out);
}

extern "C" void tm_graph_destroy(tm_graph *g) {
    if (!g) return;
    DeviceGuard guard(g->device);       // also runs from NeighborFinder.__del__ at GC time: the caller's device is restored
    cudaFree((void *)g->v.off); cudaFree((void *)g->v.entry); cudaFree((void *)g->v.skey); cudaFree((void *)g->v.etab); cudaFree((void *)g->v.htab);
    delete g;
}

extern "C" int tm_graph_sizes(const tm_graph *g, int64_t *n_nodes, int64_t *n_entries, int64_t *max_eidx, int64_t *device_bytes) {
    if (!g) { set_error("tm_graph_sizes: null graph"); return TM_ERR_ARG; }
    if (n_nodes) *n_nodes = g->v.n_nodes;
    if (n_entries) *n_entries = g->v.n_entries;
    if (max_eidx) *max_eidx = g->v.max_eidx;
    if (device_bytes) *device_bytes = g->device_bytes;
    return TM_OK;
}

extern "C" int tm_graph_export(const tm_graph *g, int64_t *h_off, int32_t *h_nbr, int32_t *h_eidx, double *h_ts) {
    if (!g) { set_error("tm_graph_export: null graph"); return TM_ERR_ARG; }
    TM_DEVICE(g->device);
    if (h_off) TM_CUDA(cudaMemcpy(h_off, g->v.off, sizeof(int64_t) * (g->v.n_nodes + 1), cudaMemcpyDeviceToHost));
    if ((h_nbr || h_eidx || h_ts) && g->v.n_entries) {
        std::vector<Entry> ent(g->v.n_entries);
        TM_CUDA(cudaMemcpy(ent.data(), g->v.entry, sizeof(Entry) * g->v.n_entries, cudaMemcpyDeviceToHost));
        for (int64_t p = 0; p < g->v.n_entries; ++p) {
            if (h_nbr) h_nbr[p] = ent[p].nbr;
            if (h_eidx) h_eidx[p] = ent[p].eidx;
            if (h_ts) h_ts[p] = ent[p].ts;
        }
    }
    return TM_OK;
}

extern "C" int tm_graph_export_edge_table(const tm_graph *g, int32_t *h_tab) {
    if (!g || !h_tab) { set_error("tm_graph_export_edge_table: bad argument"); return TM_ERR_ARG; }
    TM_DEVICE(g->device);
    if (g->v.max_eidx >= 0) {
        std::vector<EdgeSlot> tab(g->v.max_eidx + 1);
        TM_CUDA(cudaMemcpy(tab.data(), g->v.etab, sizeof(EdgeSlot) * tab.size(), cudaMemcpyDeviceToHost));
        for (size_t e = 0; e < tab.size(); ++e) { h_tab[4 * e] = tab[e].node_a; h_tab[4 * e + 1] = tab[e].node_b; h_tab[4 * e + 2] = tab[e].cut_a; h_tab[4 * e + 3] = tab[e].cut_b; }
    }
    return TM_OK;
}

extern "C" int tm_graph_export_skey(const tm_graph *g, uint64_t *h_skey) {
    if (!g || !h_skey) { set_error("tm_graph_export_skey: bad argument"); return TM_ERR_ARG; }
    TM_DEVICE(g->device);
    if (g->v.n_entries) TM_CUDA(cudaMemcpy(h_skey, g->v.skey, sizeof(uint64_t) * g->v.n_entries, cudaMemcpyDeviceToHost));
    return TM_OK;
}

extern "C" int tm_find_before_batch(const tm_graph *g, int64_t R, const int32_t *d_node, const double *d_cut_time,
                                    const int32_t *d_eidx, int64_t *d_start, int32_t *d_cut, int32_t *d_err, tm_stream stream) {
    if (!g || R < 0 || (R > 0 && (!d_node || !d_start || !d_cut || (!d_cut_time && !d_eidx)))) { set_error("tm_find_before_batch: bad argument"); return TM_ERR_ARG; }
    if (R == 0) return TM_OK;
    TM_DEVICE(g->device);
    const int threads = 256;
    const int64_t blocks = (R * 32 + threads - 1) / threads;
    find_before_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(g->v, R, d_node, d_cut_time, d_eidx, d_start, d_cut, d_err);
    TM_LAUNCH_CHECK();
    return TM_OK;
}
