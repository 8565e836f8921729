// Sampling kernels: temporal neighbour sampling (get_temporal_neighbor), the fused 3-event walk
// sampler (get_next_step + get_final_step + anonymisation class), motif-class histograms and
// edge-identity counts.  Reference: utils/graph.py:197-476, utils/null_model.py:75-82,
// processed/data_preprocess.py:148-208,327-343.  One warp per hop row, one thread per walk slot: the row's
// windows are found with one or two 16-byte table loads, draws are counter-based (no state), the
// gathers are 128-bit loads of whole CSR entries, and the neighbour-id filter of step 3 is a key-range
// lookup in a per-node (neighbour, position)-sorted secondary index instead of an O(prefix) scan.
#include "common.cuh"

namespace tmb {

constexpr int kWarpsPerBlock = 8;

// prep-order category id (processed/data_preprocess.py:171) of anonymised row [1, c, t]
__device__ __forceinline__ int class_prep(int c, int t) {
    // c: 2 -> "1,2,t", 3 -> "1,3,t", 1 -> "1,1,t"
    const unsigned m2 = 0x2103u, m3 = 0x5647u, m1 = 0x89abu;  // nibble t of each = id
    const unsigned m = c == 2 ? m2 : (c == 3 ? m3 : m1);
    return (int)((m >> (4 * t)) & 0xfu);
}
// 0-based position of a prep-order id in the null-model key order (utils/null_model.py:90)
__constant__ int kPrepToNull[12] = {1, 3, 2, 0, 5, 6, 7, 4, 11, 10, 9, 8};

// ---------------------------------------------------------------------------------------------
// get_temporal_neighbor, utils/graph.py:197-231 (uniform / bias == 0 branch)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
sample_hop_kernel(GraphView g, int64_t R, const int32_t *__restrict__ node, const double *__restrict__ cut_time,
                  const int32_t *__restrict__ eidx, int n, uint64_t seed, uint32_t stage, uint64_t row_offset,
                  const uint32_t *__restrict__ inject, int32_t *__restrict__ o_node, int32_t *__restrict__ o_eidx,
                  float *__restrict__ o_ts, int32_t *err) {
    extern __shared__ uint32_t sh_draw[];  // [kWarpsPerBlock][n]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + warp;
    if (row >= R) return;
    uint32_t *d = sh_draw + (size_t)warp * n;
    const int64_t v = node[row];
    int64_t s = 0, c = 0;
    if (v < 0 || v >= g.n_nodes) {
        if (lane == 0) report_row_error(err, row);
    } else {
        const int32_t e = eidx ? eidx[row] : TM_EIDX_NONE;
        if (e == TM_EIDX_NONE) {
            s = __ldg(g.off + v);
            c = warp_lower_bound(g.entry + s, __ldg(g.off + v + 1) - s, cut_time ? cut_time[row] : 0.0, lane);
        } else if (v > 0) {                     // window start and cut from the edge's row: no access to off[]
            c = dict_get(g, v, e, &s);
            if (c < 0) { c = 0; if (lane == 0) report_row_error(err, row); }
        }
    }
    const int64_t ob = row * n;
    if (c == 0) {  // no previous neighbours: padding row, no draw consumed (graph.py:214-215)
        for (int k = lane; k < n; k += 32) { o_node[ob + k] = 0; o_eidx[ob + k] = 0; o_ts[ob + k] = 0.f; }
        return;
    }
    for (int k = lane; k < n; k += 32) {
        uint64_t x;
        if (inject) {
            x = inject[ob + k];
            if (x >= (uint64_t)c) { x = c - 1; report_row_error(err, row); }
        } else x = draw_index(seed, stage, row_offset + row, k, (uint64_t)c);
        d[k] = (uint32_t)x;
    }
    __syncwarp();
    for (int k = lane; k < n; k += 32) {  // np.sort of the sampled indices (graph.py:218) by rank
        const uint32_t mine = d[k];
        int rank = 0;
        for (int j = 0; j < n; ++j) { const uint32_t x = d[j]; rank += (x < mine) || (x == mine && j < k); }
        const Entry en = load_entry(g.entry + s + mine);
        o_node[ob + rank] = en.nbr; o_eidx[ob + rank] = en.eidx; o_ts[ob + rank] = (float)en.ts;
    }
}

// ---------------------------------------------------------------------------------------------
// Neighbour-id filter of get_final_step cases 1/2 (graph.py:358-371,398-411) without the O(prefix)
// scan: skey[] holds, per node, (nbr << 32 | position) sorted, so "entries of prefix(v, cut) whose
// neighbour is x" is the key range [(x,0), (x,cut)) and the k-th of them (in position order) is one load.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t key_lower_bound(const uint64_t *__restrict__ k, int64_t lo, int64_t hi, uint64_t key) {
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(k + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}
struct IdRange {              // positions (< cut) of neighbour x in the node's window, ascending: skey[base .. base+cnt), or the first cnt of pos[] (inl)
    int64_t base, cnt;
    uint32_t pos[4];
    bool inl;
};
__device__ __forceinline__ uint32_t range_pos(const uint64_t *__restrict__ k, const IdRange &r, int64_t i) {
    if (r.inl) return i == 0 ? r.pos[0] : i == 1 ? r.pos[1] : i == 2 ? r.pos[2] : r.pos[3];
    return (uint32_t)__ldg(k + r.base + i);
}
// k = skey of the node's window (length len).  With the run directory the run of x costs one or two 32-byte probes instead of a
// lower bound over the whole window; runs of up to four entries are answered from the probe itself, longer ones by a lower bound over
// the run only.
__device__ __forceinline__ IdRange id_prefix(const GraphView &g, int64_t node, const uint64_t *__restrict__ k, int64_t win, int64_t len, int64_t cut, int32_t x) {
    IdRange r;
    r.inl = false; r.base = 0; r.cnt = 0;
    r.pos[0] = r.pos[1] = r.pos[2] = r.pos[3] = 0xffffffffu;
    const uint64_t hi_key = (uint64_t)(uint32_t)x << 32;
    if (g.htab) {
        const uint64_t key = (uint64_t)node << 32 | (uint32_t)x;
        uint64_t slot = mix64(key) & g.hmask;
        for (;;) {
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(g.htab + slot));
            const uint64_t kk = (uint64_t)q.y << 32 | q.x;
            if (kk == key) {
                if (cut <= 0) break;
                if (q.w <= 4u) {                             // the whole run is inline; positions are ascending, absent ones are 0xffffffff
                    const uint4 pp = __ldg(reinterpret_cast<const uint4 *>(g.htab + slot) + 1);
                    r.inl = true;
                    r.pos[0] = pp.x; r.pos[1] = pp.y; r.pos[2] = pp.z; r.pos[3] = pp.w;
                    const uint64_t c = (uint64_t)cut;
                    r.cnt = (int64_t)(pp.x < c) + (pp.y < c) + (pp.z < c) + (pp.w < c);
                } else {
                    r.base = (int64_t)q.z - win;
                    r.cnt = key_lower_bound(k, r.base, r.base + (int64_t)q.w, hi_key | (uint64_t)cut) - r.base;
                }
                break;
            }
            if (kk == ~0ull) break;
            slot = (slot + 1) & g.hmask;
        }
        return r;
    }
    r.base = key_lower_bound(k, 0, len, hi_key);
    r.cnt = cut > 0 ? key_lower_bound(k, r.base, len, hi_key | (uint64_t)cut) - r.base : 0;
    return r;
}
// position of the element of merged rank kk in the union of two ascending position lists (all positions distinct)
__device__ __forceinline__ int64_t select_union(const uint64_t *__restrict__ k, const IdRange &a, const IdRange &b, int64_t kk) {
    int64_t lo = max((int64_t)0, kk - b.cnt), hi = min(kk, a.cnt);
    while (lo < hi) {   // lo = how many of the kk smaller elements come from a
        const int64_t mid = (lo + hi) >> 1;
        if (range_pos(k, a, mid) < range_pos(k, b, kk - mid - 1)) lo = mid + 1; else hi = mid;
    }
    const uint32_t pa = lo < a.cnt ? range_pos(k, a, lo) : 0xffffffffu;
    const uint32_t pb = kk - lo < b.cnt ? range_pos(k, b, kk - lo) : 0xffffffffu;
    return (int64_t)min(pa, pb);
}

// ---------------------------------------------------------------------------------------------
// find_k_walks = get_next_step (graph.py:308-333) + get_final_step (:335-476).  One THREAD per
// first-hop slot (it owns the N2 walks of that slot): every step is a short chain of dependent
// 16-byte loads, so the latency is hidden by thread-level parallelism (>100k slots in flight).
// ---------------------------------------------------------------------------------------------
template <int CAP>
#ifndef TM_WALKS_MINBLOCKS
#define TM_WALKS_MINBLOCKS 4
#endif
__global__ void __launch_bounds__(256, TM_WALKS_MINBLOCKS)
sample_walks_kernel(GraphView g, int64_t B, int n, int N2, const int32_t *__restrict__ root,
                    const int32_t *__restrict__ h1_node, const int32_t *__restrict__ h1_eidx, const float *__restrict__ h1_ts,
                    uint64_t seed, uint64_t row_offset, const uint32_t *__restrict__ inj2, const uint32_t *__restrict__ inj3,
                    const int32_t *__restrict__ pre2,   // optional [B*n*N2][3] = given (src2, tgt2, e2): get_final_step on its own
                    const float *__restrict__ pre2_t,
                    int32_t *__restrict__ o_nodes, int32_t *__restrict__ o_eidx, float *__restrict__ o_t,
                    int32_t *__restrict__ o_anony, uint8_t *__restrict__ o_cat,
                    unsigned long long *hist_null, unsigned long long *hist_prep, unsigned long long *scanned, int staged) {
    __shared__ unsigned int sh_hist[12];
    __shared__ unsigned long long sh_scan;
    extern __shared__ __align__(16) int32_t stage[];   // staged: nodes [256 N2][6] | eidx [256 N2][3] | t [256 N2][3] of the block's walks
    if (threadIdx.x < 12) sh_hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) sh_scan = 0;
    __syncthreads();
    const int64_t r2 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // loop variable i of get_next_step
    if (r2 < B * n) {
        const int64_t W = (int64_t)n * N2;
        const int64_t b = r2 / n;
        const int64_t s1 = root[b], t1n = h1_node[r2];
        const int32_t e1 = h1_eidx[r2];
        const float t1 = h1_ts[r2];
        auto in_range = [&](int64_t v) { return v >= 0 && v < g.n_nodes; };
        // ---- step 2: find_before_walk([root, nbr], e_idx=e1): node 0 -> 0, missing key -> 0 (graph.py:174-176)
        int64_t c_a = 0, c_b = 0, s_a = 0, s_b = 0;
        {
            const EdgeSlot t = edge_slot(g, e1);            // one sector: both nodes' cuts and window starts
            if (s1 > 0) { if (s1 == t.node_a) { c_a = t.cut_a; s_a = t.start_a; } else if (s1 == t.node_b) { c_a = t.cut_b; s_a = t.start_b; } }
            if (t1n > 0) { if (t1n == t.node_a) { c_b = t.cut_a; s_b = t.start_a; } else if (t1n == t.node_b) { c_b = t.cut_b; s_b = t.start_b; } }
        }
        const int64_t L = pre2 ? 0 : c_a + c_b;
        uint64_t d[CAP];
        if (L > 0) {
#pragma unroll
            for (int j = 0; j < CAP; ++j)
                if (j < N2) d[j] = inj2 ? min((uint64_t)inj2[r2 * N2 + j], (uint64_t)L - 1) : draw_index(seed, TM_STAGE_STEP2, row_offset * n + r2, j, (uint64_t)L);
#pragma unroll
            for (int i = 1; i < CAP; ++i)      // np.sort (graph.py:328): insertion sort of <= CAP draws
                if (i < N2) {
                    const uint64_t v = d[i];
                    int j = i - 1;
                    while (j >= 0 && d[j] > v) { d[j + 1] = d[j]; --j; }
                    d[j + 1] = v;
                }
        }
        unsigned long long scan_acc = 0;
        for (int jj = 0; jj < N2; ++jj) {
            int64_t s2 = 0, t2n = 0; int32_t e2 = 0; float t2 = 0.f;
            if (pre2) {
                const int64_t w2 = r2 * N2 + jj;
                s2 = pre2[w2 * 3]; t2n = pre2[w2 * 3 + 1]; e2 = pre2[w2 * 3 + 2]; t2 = pre2_t ? pre2_t[w2] : 0.f;
            } else if (L > 0) {
                const int64_t sd = (int64_t)d[jj];
                const bool from_a = sd < c_a;
                const Entry en = load_entry(g.entry + (from_a ? s_a + sd : s_b + (sd - c_a)));
                s2 = from_a ? s1 : t1n; t2n = en.nbr; e2 = en.eidx; t2 = (float)en.ts;   // graph.py:329-332
            }
            // ---- step 3: get_final_step for walk w
            int64_t A, Bn; int32_t fa1, fa2, fb; int code;
            if (s1 == s2 && t1n != t2n) { A = s1; Bn = t2n; fa1 = (int32_t)t1n; fa2 = (int32_t)t2n; fb = (int32_t)t1n; code = 2; }       // :355
            else if (t1n == s2 && s1 != t2n) { A = t1n; Bn = t2n; fa1 = (int32_t)s1; fa2 = (int32_t)t2n; fb = (int32_t)s1; code = 3; }  // :395
            else { A = t1n; Bn = t2n; fa1 = fa2 = fb = -1; code = 1; }                                                                     // :436
            // cut = nodeedge2idx[x].get(e2) if x > 0 else 0; None -> whole list (graph.py:357-358)
            int64_t cA = 0, cB = 0, sA = 0, sB = 0, lenA = -1, lenB = -1;       // len < 0: not loaded (only the directory-less filter and the None case need it)
            {
                const EdgeSlot t = edge_slot(g, e2);
                auto window = [&](int64_t v, int64_t &c, int64_t &st, int64_t &len) {
                    if (v <= 0 || !in_range(v)) return;
                    if (v == t.node_a) { c = t.cut_a; st = t.start_a; }
                    else if (v == t.node_b) { c = t.cut_b; st = t.start_b; }
                    else { st = __ldg(g.off + v); len = __ldg(g.off + v + 1) - st; c = len; }    // dict.get -> None -> [:None] = the whole list
                    if (!g.htab && len < 0) len = __ldg(g.off + v + 1) - st;
                };
                window(A, cA, sA, lenA);
                window(Bn, cB, sB, lenB);
            }
            int64_t nA, nB;
            IdRange ra1, ra2, rb;
            ra1.cnt = ra2.cnt = rb.cnt = 0; ra1.base = ra2.base = rb.base = 0; ra1.inl = ra2.inl = rb.inl = false;
            if (code == 1) { nA = cA; nB = cB; }
            else {
                ra1 = id_prefix(g, A, g.skey + sA, sA, lenA, cA, fa1);
                if (fa2 != fa1) ra2 = id_prefix(g, A, g.skey + sA, sA, lenA, cA, fa2);
                rb = id_prefix(g, Bn, g.skey + sB, sB, lenB, cB, fb);
                nA = ra1.cnt + ra2.cnt; nB = rb.cnt;
                scan_acc += (unsigned long long)(cA + cB);
            }
            int64_t src3 = 0, tgt3 = 0; int32_t e3 = 0; float t3 = 0.f; int tc = 0;
            const int64_t w = r2 * N2 + jj;
            if (nA + nB > 0) {
                int64_t k = inj3 ? (int64_t)min((uint64_t)inj3[w], (uint64_t)(nA + nB) - 1)
                                 : (int64_t)draw_index(seed, TM_STAGE_STEP3, row_offset * W + w, 0, (uint64_t)(nA + nB));
                int64_t p;
                if (k < nA) { src3 = A; p = sA + (code == 1 ? k : select_union(g.skey + sA, ra1, ra2, k)); }
                else { k -= nA; src3 = Bn; p = sB + (code == 1 ? k : (int64_t)range_pos(g.skey + sB, rb, k)); }
                const Entry en = load_entry(g.entry + p);
                tgt3 = en.nbr; e3 = en.eidx; t3 = (float)en.ts;
                if (code == 2) tc = (src3 == s1 && tgt3 == t1n) ? 1 : (src3 == s1 && tgt3 == t2n) ? 2 : (src3 == t1n && tgt3 == t2n) ? 3 : 0;      // :386-393
                else if (code == 3) tc = (src3 == t1n && tgt3 == s1) ? 1 : (src3 == t1n && tgt3 == t2n) ? 3 : (src3 == t2n && tgt3 == s1) ? 2 : 0; // :427-434
                else tc = (src3 == s1 && tgt3 != t1n) ? 3 : (src3 == t1n && tgt3 != s1) ? 2 : (src3 == s1 && tgt3 == t1n) ? 1 : (src3 == t1n && tgt3 == s1) ? 1 : 0;  // :464-473
            }
            if (staged) {       // the block's walks are contiguous in every output array: park them in shared memory, write them out coalesced below
                const int lw = (int)threadIdx.x * N2 + jj;
                int32_t *sn = stage + lw * 6, *se = stage + (int)blockDim.x * N2 * 6 + lw * 3;
                float *st_ = reinterpret_cast<float *>(stage + (int)blockDim.x * N2 * 9) + lw * 3;
                sn[0] = (int32_t)src3; sn[1] = (int32_t)tgt3; sn[2] = (int32_t)s2; sn[3] = (int32_t)t2n; sn[4] = (int32_t)s1; sn[5] = (int32_t)t1n;
                se[0] = e3; se[1] = e2; se[2] = e1;
                st_[0] = t3; st_[1] = t2; st_[2] = t1;
            } else {
                int2 *on = reinterpret_cast<int2 *>(o_nodes + w * 6);   // [src3,tgt3,src2,tgt2,src1,tgt1], graph.py:303
                on[0] = make_int2((int32_t)src3, (int32_t)tgt3);
                on[1] = make_int2((int32_t)s2, (int32_t)t2n);
                on[2] = make_int2((int32_t)s1, (int32_t)t1n);
                o_eidx[w * 3 + 0] = e3; o_eidx[w * 3 + 1] = e2; o_eidx[w * 3 + 2] = e1;     // :304
                o_t[w * 3 + 0] = t3; o_t[w * 3 + 1] = t2; o_t[w * 3 + 2] = t1;               // :305
            }
            if (o_anony) { o_anony[w * 3 + 0] = 1; o_anony[w * 3 + 1] = code; o_anony[w * 3 + 2] = tc; }
            const int cls = class_prep(code, tc);
            if (o_cat) o_cat[w] = (uint8_t)cls;
            if (hist_null || hist_prep) atomicAdd(&sh_hist[cls], 1u);
        }
        if (scanned && scan_acc) atomicAdd(&sh_scan, scan_acc);
    }
    __syncthreads();
    if (staged) {           // 16-byte stores of the block's contiguous output ranges (a thread's own stores would touch one sector per value)
        const int64_t row0 = (int64_t)blockIdx.x * blockDim.x;
        const int64_t live = min((int64_t)blockDim.x, B * n - row0) * N2;                  // walks of this block
        auto flush = [&](const int32_t *src, int32_t *dst, int per_walk) {
            const int64_t words = live * per_walk;                                         // multiple of 4 except in the last block
            dst += row0 * N2 * per_walk;
            for (int64_t i = 4 * (int64_t)threadIdx.x; i + 3 < words; i += 4 * blockDim.x) *reinterpret_cast<int4 *>(dst + i) = *reinterpret_cast<const int4 *>(src + i);
            for (int64_t i = (words & ~(int64_t)3) + threadIdx.x; i < words; i += blockDim.x) dst[i] = src[i];
        };
        flush(stage, o_nodes, 6);
        flush(stage + (int)blockDim.x * N2 * 6, o_eidx, 3);
        flush(stage + (int)blockDim.x * N2 * 9, reinterpret_cast<int32_t *>(o_t), 3);
    }
    if (threadIdx.x < 12 && sh_hist[threadIdx.x]) {
        if (hist_prep) atomicAdd(hist_prep + threadIdx.x, (unsigned long long)sh_hist[threadIdx.x]);
        if (hist_null) atomicAdd(hist_null + kPrepToNull[threadIdx.x], (unsigned long long)sh_hist[threadIdx.x]);
    }
    if (threadIdx.x == 0 && scanned && sh_scan) atomicAdd(scanned, sh_scan);
}

// ---------------------------------------------------------------------------------------------
// statistic (utils/null_model.py:75-82) / marginal ids (processed/data_preprocess.py:171-208)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
class_hist_kernel(int64_t count, const int32_t *__restrict__ anony, unsigned long long *hist_null,
                  unsigned long long *hist_prep, uint8_t *__restrict__ o_cat, int32_t *err) {
    __shared__ unsigned int sh_hist[12];
    if (threadIdx.x < 12) sh_hist[threadIdx.x] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
        const int a0 = anony[3 * i], c = anony[3 * i + 1], t = anony[3 * i + 2];
        const bool ok = a0 == 1 && c >= 1 && c <= 3 && t >= 0 && t <= 3;
        if (!ok) { report_row_error(err, i); if (o_cat) o_cat[i] = 255; continue; }   // KeyError in the reference
        const int cls = class_prep(c, t);
        if (o_cat) o_cat[i] = (uint8_t)cls;
        // warp-aggregated: one shared-memory atomic per distinct class per warp
        const unsigned peers = __match_any_sync(__activemask(), cls);
        if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&sh_hist[cls], (unsigned)__popc(peers));
    }
    __syncthreads();
    if (threadIdx.x < 12 && sh_hist[threadIdx.x]) {
        if (hist_prep) atomicAdd(hist_prep + threadIdx.x, (unsigned long long)sh_hist[threadIdx.x]);
        if (hist_null) atomicAdd(hist_null + kPrepToNull[threadIdx.x], (unsigned long long)sh_hist[threadIdx.x]);
    }
}

// ---------------------------------------------------------------------------------------------
// new_edge_info (processed/data_preprocess.py:327-343): for every walk event, how many walks of the same root carry its edge id
// at position 0, 1, 2 (edge id 0 = padding, counted like any id).  One WARP per root: the root's 3W ids are counted in a small
// open-addressing table in shared memory (key, three 10-bit position counters packed in one word), then every id looks its slot up
// again: O(3W) table operations per root instead of the (3W)^2 comparisons of the all-pairs form.
// ---------------------------------------------------------------------------------------------
constexpr int kEidWarps = 8;
template <typename OutT>      // float (the reference's dtype) or uint8_t (the pipeline's compact form: counts <= W <= 255)
__global__ void __launch_bounds__(kEidWarps * 32)
edge_identity_kernel(int64_t B, int W, int slots_mask, const int32_t *__restrict__ eidx, OutT *__restrict__ out) {
    extern __shared__ int32_t sh_tab[];         // per warp: keys [slots], counters [slots]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, slots = slots_mask + 1;
    int32_t *keys = sh_tab + (size_t)warp * 2 * slots;
    unsigned *cnt = reinterpret_cast<unsigned *>(keys + slots);
    const int64_t b = (int64_t)blockIdx.x * kEidWarps + warp;
    if (b >= B) return;
    const int n3 = W * 3;
    const int32_t *src = eidx + b * n3;
    for (int i = lane; i < slots; i += 32) { keys[i] = INT32_MIN; cnt[i] = 0u; }
    __syncwarp();
    auto hash = [&](int32_t id) { return (int)(((uint32_t)id * 2654435761u) >> 10) & slots_mask; };
    for (int i = lane; i < n3; i += 32) {
        const int32_t id = src[i];
        const unsigned inc = 1u << (10 * (i % 3));              // position of the event inside its walk
        for (int slot = hash(id);; slot = (slot + 1) & slots_mask) {
            const int32_t prev = atomicCAS(&keys[slot], INT32_MIN, id);
            if (prev == INT32_MIN || prev == id) { atomicAdd(&cnt[slot], inc); break; }
        }
    }
    __syncwarp();
    for (int i = lane; i < n3; i += 32) {
        const int32_t id = src[i];
        int slot = hash(id);
        while (keys[slot] != id) slot = (slot + 1) & slots_mask;
        const unsigned c = cnt[slot];
        if (sizeof(OutT) == 1) {            // bytes: [walk][position][4] = three counts and a pad byte, one 4-byte store (and one 4-byte load in the scorer)
            reinterpret_cast<uchar4 *>(out)[b * n3 + i] = make_uchar4((unsigned char)(c & 1023u), (unsigned char)((c >> 10) & 1023u), (unsigned char)((c >> 20) & 1023u), 0);
        } else {
            OutT *o = out + (b * n3 + i) * 3;
            o[0] = (OutT)(c & 1023u); o[1] = (OutT)((c >> 10) & 1023u); o[2] = (OutT)((c >> 20) & 1023u);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// get_next_step with e_idx_l = None (utils/graph.py:308-333 through find_before_walk's bisect branch, :170-171): the two prefixes of
// row i, [source_i, nbr_i], are cut by TIME (strict lower bound on the float64 timestamps; no node-0 special case on this branch).
// find_k_walks never takes this branch (it always passes e_idx); it exists for callers of get_next_step itself.  One thread per row.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t thread_lower_bound(const Entry *__restrict__ base, int64_t len, double x) {
    int64_t lo = 0, hi = len;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (__ldg(&base[mid].ts) < x) lo = mid + 1; else hi = mid; }
    return lo;
}
__global__ void __launch_bounds__(256)
next_step_time_kernel(GraphView g, int64_t R, int N2, const int32_t *__restrict__ source, const int32_t *__restrict__ nbr, const double *__restrict__ cut_time,
                      uint64_t seed, uint64_t row_offset, const uint32_t *__restrict__ inj, int32_t *__restrict__ o_src, int32_t *__restrict__ o_tgt,
                      int32_t *__restrict__ o_e, float *__restrict__ o_t, int32_t *err) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    const int64_t a = source[i], b = nbr[i];
    for (int j = 0; j < N2; ++j) { o_src[i * N2 + j] = 0; o_tgt[i * N2 + j] = 0; o_e[i * N2 + j] = 0; o_t[i * N2 + j] = 0.f; }
    if (a < 0 || a >= g.n_nodes || b < 0 || b >= g.n_nodes) { report_row_error(err, i); return; }
    const int64_t s_a = __ldg(g.off + a), s_b = __ldg(g.off + b);
    const int64_t c_a = thread_lower_bound(g.entry + s_a, __ldg(g.off + a + 1) - s_a, cut_time[i]);
    const int64_t c_b = thread_lower_bound(g.entry + s_b, __ldg(g.off + b + 1) - s_b, cut_time[i]);
    const int64_t L = c_a + c_b;
    if (L == 0) return;
    uint64_t d[TM_MAX_STEP2_FANOUT];
    for (int j = 0; j < N2; ++j) d[j] = inj ? min((uint64_t)inj[i * N2 + j], (uint64_t)L - 1) : draw_index(seed, TM_STAGE_STEP2, row_offset + i, j, (uint64_t)L);
    for (int x = 1; x < N2; ++x) {            // np.sort (graph.py:328)
        const uint64_t v = d[x];
        int y = x - 1;
        while (y >= 0 && d[y] > v) { d[y + 1] = d[y]; --y; }
        d[y + 1] = v;
    }
    for (int j = 0; j < N2; ++j) {
        const int64_t sd = (int64_t)d[j];
        const bool from_a = sd < c_a;
        const Entry en = load_entry(g.entry + (from_a ? s_a + sd : s_b + (sd - c_a)));
        o_src[i * N2 + j] = (int32_t)(from_a ? a : b); o_tgt[i * N2 + j] = en.nbr; o_e[i * N2 + j] = en.eidx; o_t[i * N2 + j] = (float)en.ts;
    }
}

}  // namespace tm

using namespace tmb;

extern "C" int tm_sample_hop(const tm_graph *g, int64_t R, const int32_t *d_node, const double *d_cut_time,
                             const int32_t *d_eidx, int n, uint64_t seed, uint32_t stage, uint64_t row_offset,
                             const uint32_t *d_inject, int32_t *d_o_node, int32_t *d_o_eidx, float *d_o_ts,
                             int32_t *d_err, tm_stream stream) {
    if (!g || R < 0 || n <= 0 || (R > 0 && (!d_node || !d_o_node || !d_o_eidx || !d_o_ts || (!d_cut_time && !d_eidx)))) {
        set_error("tm_sample_hop: bad argument");
        return TM_ERR_ARG;
    }
    const size_t smem = sizeof(uint32_t) * kWarpsPerBlock * (size_t)n;
    if (smem > 48 * 1024) { set_error("tm_sample_hop: fan-out %d too large (max %d)", n, 48 * 1024 / 4 / kWarpsPerBlock); return TM_ERR_UNSUPPORTED; }
    if (R == 0) return TM_OK;
    TM_DEVICE(g->device);
    const int64_t blocks = (R + kWarpsPerBlock - 1) / kWarpsPerBlock;
    sample_hop_kernel<<<(unsigned)blocks, kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
        g->v, R, d_node, d_cut_time, d_eidx, n, seed, stage, row_offset, d_inject, d_o_node, d_o_eidx, d_o_ts, d_err);
    TM_LAUNCH_CHECK();
    return TM_OK;
}

// find_k_hop (utils/graph.py:233-262) as one call: hop 0 on the roots (window by e_idx where d_eidx has one, else by time), hop l >= 1 on
// the flattened records of hop l - 1, looked up by e_idx (:247-250).  h_o_node / h_o_eidx / h_o_ts: HOST arrays of k device pointers, hop l
// = [B, n^(l+1)].  Draw contract: stage l, row = row_offset * n^l + i (so shards and chunks reproduce the single call).
extern "C" int tm_sample_khop(const tm_graph *g, int64_t B, int k, int n, const int32_t *d_root, const double *d_cut_time, const int32_t *d_eidx,
                              uint64_t seed, uint64_t row_offset, int32_t *const *h_o_node, int32_t *const *h_o_eidx, float *const *h_o_ts,
                              int32_t *d_err, tm_stream stream) {
    if (!g || B < 0 || k < 0 || n <= 0 || (k > 0 && (!h_o_node || !h_o_eidx || !h_o_ts))) { set_error("tm_sample_khop: bad argument"); return TM_ERR_ARG; }
    if (B == 0) return TM_OK;                        // empty batch: nothing to write (the outputs may be null)
    int64_t rows = B;
    uint64_t off = row_offset;
    for (int l = 0; l < k; ++l) {
        if (!h_o_node[l] || !h_o_eidx[l] || !h_o_ts[l]) { set_error("tm_sample_khop: null output for hop %d", l); return TM_ERR_ARG; }
        const int rc = l == 0 ? tm_sample_hop(g, rows, d_root, d_cut_time, d_eidx, n, seed, 0, off, nullptr, h_o_node[0], h_o_eidx[0], h_o_ts[0], d_err, stream)
                              : tm_sample_hop(g, rows, h_o_node[l - 1], nullptr, h_o_eidx[l - 1], n, seed, (uint32_t)l, off, nullptr, h_o_node[l], h_o_eidx[l],
                                              h_o_ts[l], d_err, stream);
        if (rc != TM_OK) return rc;
        rows *= n; off *= (uint64_t)n;
    }
    return TM_OK;
}

static int walks_impl(const tm_graph *g, int64_t B, int n, int N2, const int32_t *d_root,
                      const int32_t *d_h1_node, const int32_t *d_h1_eidx, const float *d_h1_ts,
                      uint64_t seed, uint64_t row_offset, const uint32_t *d_inject2, const uint32_t *d_inject3,
                      const int32_t *d_pre2, const float *d_pre2_t,
                      int32_t *d_o_nodes, int32_t *d_o_eidx, float *d_o_t, int32_t *d_o_anony, uint8_t *d_o_cat,
                      unsigned long long *d_hist_null, unsigned long long *d_hist_prep,
                      unsigned long long *d_scanned, tm_stream stream) {
    if (!g || B < 0 || n <= 0 || N2 <= 0 || (B > 0 && (!d_root || !d_h1_node || !d_h1_eidx || !d_h1_ts || !d_o_nodes || !d_o_eidx || !d_o_t))) {
        set_error("tm_sample_walks: bad argument");
        return TM_ERR_ARG;
    }
    if (N2 > TM_MAX_STEP2_FANOUT) { set_error("tm_sample_walks: step-2 fan-out %d > %d", N2, TM_MAX_STEP2_FANOUT); return TM_ERR_UNSUPPORTED; }
    if (B == 0) return TM_OK;
    TM_DEVICE(g->device);
    // threads per block: the block's slots finish together (staged outputs), so smaller blocks wait less for their slowest walk -- 64 at N2 = 1
    // (cfg2: 4.41 -> 3.47 ms per step), 128 at N2 <= 4 (cfg5: 8.95 -> 8.70 ms); wider walks are written directly from 256-thread blocks (staging
    // 30 KB per 128 slots at N2 = 5 costs more occupancy than the coalesced stores return: cfg4 1.33 vs 1.80 ms).  TEMPME_WALKS_BLOCK overrides.
    static const int wb_env = []() { const char *e = getenv("TEMPME_WALKS_BLOCK"); const int v = e ? atoi(e) : 0; return v == 64 || v == 128 || v == 256 ? v : 0; }();
    const int wb = wb_env ? wb_env : N2 == 1 ? 64 : N2 <= 4 ? 128 : 256;
    const int64_t rows = B * n, blocks = (rows + wb - 1) / wb;
    // outputs staged in shared memory and written coalesced when a block's walks fit 48 KB (N2 <= 4 at 48 bytes per walk); the base
    // addresses of the block ranges are 16-byte aligned when the arrays are (256 N2 walks x 12 / 24 bytes)
    const size_t stage_bytes = (size_t)wb * N2 * 48;
    const bool staged = N2 <= 4 && stage_bytes <= 48 * 1024 && !getenv("TEMPME_WALKS_NO_STAGING") && ((uintptr_t)d_o_nodes % 16 == 0) && ((uintptr_t)d_o_eidx % 16 == 0) &&
                        ((uintptr_t)d_o_t % 16 == 0);
    // 48 KB of staging (N2 = 4) plus the kernel's static shared memory is above the default 48 KB limit: opt in once per device
    static bool attr_set[64] = {false};
    if (g->device >= 0 && g->device < 64 && !attr_set[g->device]) {
        TM_CUDA(cudaFuncSetAttribute(sample_walks_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        TM_CUDA(cudaFuncSetAttribute(sample_walks_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        TM_CUDA(cudaFuncSetAttribute(sample_walks_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        TM_CUDA(cudaFuncSetAttribute(sample_walks_kernel<TM_MAX_STEP2_FANOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr_set[g->device] = true;
    }
#define TM_WALKS(CAP) sample_walks_kernel<CAP><<<(unsigned)blocks, wb, staged ? stage_bytes : 0, (cudaStream_t)stream>>>(    \
        g->v, B, n, N2, d_root, d_h1_node, d_h1_eidx, d_h1_ts, seed, row_offset, d_inject2, d_inject3, d_pre2, d_pre2_t,     \
        d_o_nodes, d_o_eidx, d_o_t, d_o_anony, d_o_cat, d_hist_null, d_hist_prep, d_scanned, staged ? 1 : 0)
    if (N2 == 1) TM_WALKS(1); else if (N2 <= 4) TM_WALKS(4); else if (N2 <= 8) TM_WALKS(8); else TM_WALKS(TM_MAX_STEP2_FANOUT);
#undef TM_WALKS
    TM_LAUNCH_CHECK();
    return TM_OK;
}

extern "C" int tm_sample_walks(const tm_graph *g, int64_t B, int n, int N2, const int32_t *d_root,
                               const int32_t *d_h1_node, const int32_t *d_h1_eidx, const float *d_h1_ts,
                               uint64_t seed, uint64_t row_offset, const uint32_t *d_inject2, const uint32_t *d_inject3,
                               int32_t *d_o_nodes, int32_t *d_o_eidx, float *d_o_t, int32_t *d_o_anony, uint8_t *d_o_cat,
                               unsigned long long *d_hist_null, unsigned long long *d_hist_prep,
                               unsigned long long *d_scanned, tm_stream stream) {
    return walks_impl(g, B, n, N2, d_root, d_h1_node, d_h1_eidx, d_h1_ts, seed, row_offset, d_inject2, d_inject3, nullptr, nullptr,
                      d_o_nodes, d_o_eidx, d_o_t, d_o_anony, d_o_cat, d_hist_null, d_hist_prep, d_scanned, stream);
}

extern "C" int tm_walk_final_step(const tm_graph *g, int64_t R, const int32_t *d_src1, const int32_t *d_tgt1, const int32_t *d_e1,
                                  const float *d_t1, const int32_t *d_step2, const float *d_t2, uint64_t seed, uint64_t row_offset,
                                  const uint32_t *d_inject3, int32_t *d_o_nodes, int32_t *d_o_eidx, float *d_o_t, int32_t *d_o_anony,
                                  tm_stream stream) {
    if (!d_step2) { set_error("tm_walk_final_step: d_step2 is required"); return TM_ERR_ARG; }
    // one root per walk (n = N2 = 1): row i of get_final_step is walk i
    return walks_impl(g, R, 1, 1, d_src1, d_tgt1, d_e1, d_t1, seed, row_offset, nullptr, d_inject3, d_step2, d_t2,
                      d_o_nodes, d_o_eidx, d_o_t, d_o_anony, nullptr, nullptr, nullptr, nullptr, stream);
}

extern "C" int tm_walk_next_step_time(const tm_graph *g, int64_t R, int N2, const int32_t *d_source, const int32_t *d_nbr, const double *d_cut_time,
                                      uint64_t seed, uint64_t row_offset, const uint32_t *d_inject, int32_t *d_o_src, int32_t *d_o_tgt, int32_t *d_o_eidx,
                                      float *d_o_ts, int32_t *d_err, tm_stream stream) {
    if (!g || R < 0 || N2 <= 0 || (R > 0 && (!d_source || !d_nbr || !d_cut_time || !d_o_src || !d_o_tgt || !d_o_eidx || !d_o_ts))) {
        set_error("tm_walk_next_step_time: bad argument");
        return TM_ERR_ARG;
    }
    if (N2 > TM_MAX_STEP2_FANOUT) { set_error("tm_walk_next_step_time: fan-out %d > %d", N2, TM_MAX_STEP2_FANOUT); return TM_ERR_UNSUPPORTED; }
    if (R == 0) return TM_OK;
    TM_DEVICE(g->device);
    next_step_time_kernel<<<(unsigned)((R + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g->v, R, N2, d_source, d_nbr, d_cut_time, seed, row_offset, d_inject,
                                                                                          d_o_src, d_o_tgt, d_o_eidx, d_o_ts, d_err);
    TM_LAUNCH_CHECK();
    return TM_OK;
}

extern "C" int tm_class_hist(int64_t count, const int32_t *d_anony, unsigned long long *d_hist_null,
                             unsigned long long *d_hist_prep, uint8_t *d_o_cat, int32_t *d_err, tm_stream stream) {
    if (count < 0 || (count > 0 && !d_anony)) { set_error("tm_class_hist: bad argument"); return TM_ERR_ARG; }
    if (count == 0) return TM_OK;
    TM_DEVICE(device_of(d_anony));
    const int64_t blocks = std::min<int64_t>((count + 255) / 256, 148 * 8);
    class_hist_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(count, d_anony, d_hist_null, d_hist_prep, d_o_cat, d_err);
    TM_LAUNCH_CHECK();
    return TM_OK;
}

template <typename OutT>
static int edge_identity_impl(int64_t B, int64_t W, const int32_t *d_eidx, OutT *d_out, tm_stream stream) {
    if (B < 0 || W <= 0 || (B > 0 && (!d_eidx || !d_out))) { set_error("tm_edge_identity: bad argument"); return TM_ERR_ARG; }
    if (W > 1023) { set_error("tm_edge_identity: %lld walks per root exceed the 10-bit position counters", (long long)W); return TM_ERR_UNSUPPORTED; }
    if (sizeof(OutT) == 1 && W > 255) { set_error("tm_edge_identity_u8: %lld walks per root do not fit a byte count", (long long)W); return TM_ERR_UNSUPPORTED; }
    int slots = 64;
    while (slots < 6 * W) slots <<= 1;                       // load factor <= 1/2
    const size_t smem = sizeof(int32_t) * 2 * (size_t)slots * kEidWarps;
    if (B == 0) return TM_OK;
    TM_DEVICE(device_of(d_eidx));
    static bool attr_set[64] = {false};
    int dev = 0;
    TM_CUDA(cudaGetDevice(&dev));
    if (smem > 48 * 1024 && dev < 64 && !attr_set[dev]) {
        TM_CUDA(cudaFuncSetAttribute(edge_identity_kernel<OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set[dev] = true;
    }
    if (smem > 200 * 1024) { set_error("tm_edge_identity: %lld walks per root exceed the shared-memory window", (long long)W); return TM_ERR_UNSUPPORTED; }
    edge_identity_kernel<OutT><<<(unsigned)((B + kEidWarps - 1) / kEidWarps), kEidWarps * 32, smem, (cudaStream_t)stream>>>(B, (int)W, slots - 1, d_eidx, d_out);
    TM_LAUNCH_CHECK();
    return TM_OK;
}

extern "C" int tm_edge_identity(int64_t B, int64_t W, const int32_t *d_eidx, float *d_out, tm_stream stream) {
    return edge_identity_impl<float>(B, W, d_eidx, d_out, stream);
}

extern "C" int tm_edge_identity_u8(int64_t B, int64_t W, const int32_t *d_eidx, uint8_t *d_out, tm_stream stream) {
    return edge_identity_impl<uint8_t>(B, W, d_eidx, d_out, stream);
}
