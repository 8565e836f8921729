/*
 * tempme_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, CPU restatement of the reference's (dharunm236/TempME) temporal
 * graph hot path.  It exists to CHECK the CUDA product path; nothing under
 * tempme_b200/ may import, link or call it.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Parity status: PINNED.  tests/golden/make_golden.py runs the unmodified Python
 * reference (imported from /root/reference) with numpy.random.randint replaced
 * by the counter-based draws below, and tests/test_oracle_golden.py checks every
 * function here bit-for-bit against those committed vectors.
 *
 * Each function cites the reference lines (relative to /root/reference) it
 * restates.  The code is written for obviousness, not speed; OpenMP is used
 * only over independent rows so the same file can serve as the multi-core
 * CPU baseline ("port") in bench.py.
 *
 * Random draws (the deterministic-draw contract, DESIGN.md "RNG"):
 *   c      = Philox4x32-10(key = (seed_lo, seed_hi), ctr = (slot >> 1, row_lo, row_hi, stage))
 *   r64    = (slot & 1) ? (c[3] << 32 | c[2]) : (c[1] << 32 | c[0])
 *   index  = (r64 * L) >> 64                       -- uniform on [0, L)
 * stage: 0,1,2.. = hop level of get_temporal_neighbor, 16 = get_next_step,
 * 17 = get_final_step; row = the reference loop variable `i` + row_offset.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define ORC_OK 0
#define ORC_ERR_ARG -1
#define ORC_ERR_EIDX_NOT_FOUND -2 /* IndexError at utils/graph.py:134-135 */
#define ORC_ERR_NODE_RANGE -3
#define ORC_ERR_NOMEM -4

typedef struct {
    int64_t n_nodes;   /* len(adj_list) */
    int64_t n_entries; /* sum of list lengths */
    int64_t *off;      /* off_set_l, n_nodes+1         graph.py:44,54 */
    int32_t *nbr;      /* node_idx_l                  graph.py:49 */
    int32_t *eidx;     /* edge_idx_l                  graph.py:50 */
    double *ts;        /* node_ts_l                   graph.py:52 */
    /* nodeedge2idx[node]: sorted unique keys + dict values, graph.py:56,77-101 */
    int64_t *koff;     /* n_nodes+1 */
    int32_t *kkey;
    int64_t *kval;
} orc_graph;

/* ---------------- Philox4x32-10 (Salmon et al., SC'11) ---------------- */
static void philox4x32_10(uint32_t k0, uint32_t k1, const uint32_t c[4], uint32_t out[4]) {
    uint32_t c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void orc_philox4x32_10(uint32_t k0, uint32_t k1, const uint32_t *ctr, uint32_t *out) {
    philox4x32_10(k0, k1, ctr, out);
}

static uint64_t draw_index(uint64_t seed, uint32_t stage, uint64_t row, uint32_t slot, uint64_t L) {
    uint32_t c[4] = {slot >> 1, (uint32_t)row, (uint32_t)(row >> 32), stage}, o[4];
    philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), c, o);
    uint64_t r = (slot & 1) ? ((uint64_t)o[3] << 32 | o[2]) : ((uint64_t)o[1] << 32 | o[0]);
    return (uint64_t)(((unsigned __int128)r * L) >> 64);
}

uint64_t orc_draw_index(uint64_t seed, uint32_t stage, uint64_t row, uint32_t slot, uint64_t L) {
    return draw_index(seed, stage, row, slot, L);
}

/* ---------------- graph build: utils/graph.py:33-66 ---------------- */
typedef struct { double ts; int64_t ord; int32_t nbr, eidx; } orc_triple;

static int cmp_triple(const void *a, const void *b) {
    const orc_triple *x = a, *y = b;
    if (x->ts < y->ts) return -1;
    if (x->ts > y->ts) return 1;
    return (x->ord > y->ord) - (x->ord < y->ord); /* stable: sorted(key=ts), graph.py:48 */
}
static int cmp_i32(const void *a, const void *b) {
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
    return (x > y) - (x < y);
}
static int64_t key_find(const int32_t *keys, int64_t n, int32_t e) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t m = (lo + hi) >> 1; if (keys[m] < e) lo = m + 1; else hi = m; }
    return (lo < n && keys[lo] == e) ? lo : -1;
}

void orc_graph_free(orc_graph *g) {
    if (!g) return;
    free(g->off); free(g->nbr); free(g->eidx); free(g->ts);
    free(g->koff); free(g->kkey); free(g->kval); free(g);
}

/*
 * Entries are the flattened adj_list: entry j belongs to list entry_node[j] and
 * entries of one node appear in their insertion order (the callers append every
 * event to both endpoints in CSV order, temp_exp_main.py:135-144).
 */
int orc_graph_build(int64_t n_nodes, int64_t n_entries, const int32_t *entry_node,
                    const int32_t *entry_nbr, const int32_t *entry_eidx, const double *entry_ts,
                    orc_graph **out) {
    if (n_nodes < 0 || n_entries < 0 || !out) return ORC_ERR_ARG;
    orc_graph *g = calloc(1, sizeof(*g));
    if (!g) return ORC_ERR_NOMEM;
    g->n_nodes = n_nodes; g->n_entries = n_entries;
    g->off = calloc(n_nodes + 1, sizeof(int64_t));
    g->koff = calloc(n_nodes + 1, sizeof(int64_t));
    g->nbr = malloc((n_entries + 1) * sizeof(int32_t));
    g->eidx = malloc((n_entries + 1) * sizeof(int32_t));
    g->ts = malloc((n_entries + 1) * sizeof(double));
    g->kkey = malloc((n_entries + 1) * sizeof(int32_t));
    g->kval = malloc((n_entries + 1) * sizeof(int64_t));
    orc_triple *tr = malloc((n_entries + 1) * sizeof(orc_triple));
    int64_t *fill = calloc(n_nodes + 1, sizeof(int64_t));
    if (!g->off || !g->koff || !g->nbr || !g->eidx || !g->ts || !g->kkey || !g->kval || !tr || !fill) {
        free(tr); free(fill); orc_graph_free(g); return ORC_ERR_NOMEM;
    }
    for (int64_t j = 0; j < n_entries; ++j) {
        if (entry_node[j] < 0 || entry_node[j] >= n_nodes) { free(tr); free(fill); orc_graph_free(g); return ORC_ERR_NODE_RANGE; }
        g->off[entry_node[j] + 1]++;
    }
    for (int64_t v = 0; v < n_nodes; ++v) g->off[v + 1] += g->off[v];
    for (int64_t j = 0; j < n_entries; ++j) {
        int64_t p = g->off[entry_node[j]] + fill[entry_node[j]]++;
        tr[p].ts = entry_ts[j]; tr[p].ord = j; tr[p].nbr = entry_nbr[j]; tr[p].eidx = entry_eidx[j];
    }
    int64_t ktot = 0;
    for (int64_t v = 0; v < n_nodes; ++v) {
        int64_t s = g->off[v], len = g->off[v + 1] - s;
        qsort(tr + s, len, sizeof(orc_triple), cmp_triple);
        for (int64_t i = 0; i < len; ++i) { g->nbr[s + i] = tr[s + i].nbr; g->eidx[s + i] = tr[s + i].eidx; g->ts[s + i] = tr[s + i].ts; }
        /* get_ts2idx, graph.py:77-101, emulated literally on a sorted-key map */
        int32_t *keys = g->kkey + ktot; int64_t *vals = g->kval + ktot;
        for (int64_t i = 0; i < len; ++i) keys[i] = g->eidx[s + i];
        qsort(keys, len, sizeof(int32_t), cmp_i32);
        int64_t nk = 0;
        for (int64_t i = 0; i < len; ++i) if (i == 0 || keys[i] != keys[nk - 1]) keys[nk++] = keys[i];
        int64_t tie_lo = -1, tie_n = 0; /* tie_ts_e_indices as a slot range [tie_lo, tie_lo+tie_n) */
        double last_ts = -1.0;           /* graph.py:82 */
        for (int64_t i = 0; i < len; ++i) {
            double t = g->ts[s + i];
            vals[key_find(keys, nk, g->eidx[s + i])] = i;            /* ts2idx[e_idx] = i, :85 */
            if (t == last_ts) {                                      /* :87-91 */
                if (tie_n == 0) { tie_lo = i - 1; tie_n = 2; } else tie_n++;
            }
            if (!(t == last_ts) && tie_n > 0) {                      /* :93-98 */
                for (int64_t j = 0; j < tie_n; ++j) vals[key_find(keys, nk, g->eidx[s + tie_lo + j])] -= j;
                tie_n = 0;
            }
            last_ts = t;                                             /* :99 */
        }
        g->koff[v + 1] = ktot + nk; ktot += nk;
    }
    free(tr); free(fill);
    *out = g;
    return ORC_OK;
}

void orc_graph_sizes(const orc_graph *g, int64_t *n_nodes, int64_t *n_entries, int64_t *n_keys) {
    *n_nodes = g->n_nodes; *n_entries = g->n_entries; *n_keys = g->koff[g->n_nodes];
}
void orc_graph_export(const orc_graph *g, int64_t *off, int32_t *nbr, int32_t *eidx, double *ts) {
    memcpy(off, g->off, (g->n_nodes + 1) * sizeof(int64_t));
    memcpy(nbr, g->nbr, g->n_entries * sizeof(int32_t));
    memcpy(eidx, g->eidx, g->n_entries * sizeof(int32_t));
    memcpy(ts, g->ts, g->n_entries * sizeof(double));
}

/* nodeedge2idx[node].get(e): returns 1 and *val if present, 0 if None */
int orc_dict_get(const orc_graph *g, int64_t node, int32_t e, int64_t *val) {
    int64_t s = g->koff[node], n = g->koff[node + 1] - s;
    int64_t k = key_find(g->kkey + s, n, e);
    if (k < 0) return 0;
    *val = g->kval[s + k];
    return 1;
}

/* Python slice a[:c] on a list of length len -> effective prefix length */
static int64_t slice_len(int64_t c, int64_t len) {
    if (c < 0) { c += len; if (c < 0) c = 0; }
    if (c > len) c = len;
    return c;
}

/* bisect_left_adapt, graph.py:511-530 */
static int64_t bisect_left_adapt(const double *a, int64_t n, double x) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) / 2; if (a[mid] < x) lo = mid + 1; else hi = mid; }
    return lo;
}

/*
 * find_before, graph.py:103-146 -> window [start, start+cut).
 * use_e == 0: bisect on float64 (:129).  use_e != 0: dict lookup, node 0 -> 0 (:133), None -> IndexError.
 */
int orc_find_before(const orc_graph *g, int64_t node, double cut_time, int use_e, int32_t e,
                    int64_t *start, int64_t *cut) {
    if (node < 0 || node >= g->n_nodes) return ORC_ERR_NODE_RANGE;
    int64_t s = g->off[node], len = g->off[node + 1] - s, c;
    if (!use_e) c = bisect_left_adapt(g->ts + s, len, cut_time);
    else if (node > 0) { if (!orc_dict_get(g, node, e, &c)) return ORC_ERR_EIDX_NOT_FOUND; c = slice_len(c, len); }
    else c = 0;
    *start = s; *cut = c;
    return ORC_OK;
}

int orc_find_before_batch(const orc_graph *g, int64_t R, const int32_t *node, const double *cut_time,
                          const int32_t *e, int64_t *start, int64_t *cut) {
    for (int64_t i = 0; i < R; ++i) {
        int rc = orc_find_before(g, node[i], cut_time ? cut_time[i] : 0.0, e != NULL, e ? e[i] : 0, start + i, cut + i);
        if (rc) return rc;
    }
    return ORC_OK;
}

static void sort_u64(uint64_t *a, int n) { /* np.sort of the sampled indices, graph.py:218 */
    for (int i = 1; i < n; ++i) { uint64_t v = a[i]; int j = i - 1; while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; --j; } a[j + 1] = v; }
}

/*
 * get_temporal_neighbor, graph.py:197-231 (bias == 0, 'multinomial' path).
 * Outputs are zero-initialised [R, n]; empty window -> row stays zero, no draw (:214-215).
 */
int orc_sample_hop(const orc_graph *g, int64_t R, const int32_t *node, const double *cut_time,
                   const int32_t *e, int n, uint64_t seed, uint32_t stage, uint64_t row_offset,
                   int32_t *o_node, int32_t *o_eidx, float *o_ts) {
    int err = ORC_OK;
    memset(o_node, 0, sizeof(int32_t) * R * n); memset(o_eidx, 0, sizeof(int32_t) * R * n); memset(o_ts, 0, sizeof(float) * R * n);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < R; ++i) {
        int64_t s, c;
        int rc = orc_find_before(g, node[i], cut_time ? cut_time[i] : 0.0, e != NULL, e ? e[i] : 0, &s, &c);
        if (rc) {
#pragma omp critical
            err = rc;
            continue;
        }
        if (c == 0) continue;
        uint64_t *d = malloc(sizeof(uint64_t) * n);
        for (int k = 0; k < n; ++k) d[k] = draw_index(seed, stage, row_offset + i, k, (uint64_t)c);
        sort_u64(d, n);
        for (int k = 0; k < n; ++k) {
            o_node[i * n + k] = g->nbr[s + d[k]];
            o_ts[i * n + k] = (float)g->ts[s + d[k]];   /* .astype(np.float32), :208,229 */
            o_eidx[i * n + k] = g->eidx[s + d[k]];
        }
        free(d);
    }
    return err;
}

/* prefix length used by find_before_walk (graph.py:174-176): node 0 -> 0, None -> 0 */
static int64_t walk_cut_step2(const orc_graph *g, int64_t node, int32_t e) {
    int64_t c;
    if (node <= 0) return 0;
    if (!orc_dict_get(g, node, e, &c)) return 0;
    return slice_len(c, g->off[node + 1] - g->off[node]);
}
/* prefix length used by get_final_step (graph.py:357,366,...): node 0 -> 0, None -> whole list ([:None]) */
static int64_t walk_cut_step3(const orc_graph *g, int64_t node, int32_t e) {
    int64_t c, len = g->off[node + 1] - g->off[node];
    if (node <= 0) return 0;
    if (!orc_dict_get(g, node, e, &c)) return len;
    return slice_len(c, len);
}

/*
 * find_k_walks, graph.py:265-306 = get_next_step (:308-333) + get_final_step (:335-476).
 * root[B]; h1_* [B, n] = first-hop sample (find_k_hop record 0).
 * Outputs: nodes [B, W, 6] = [src3,tgt3,src2,tgt2,src1,tgt1], eidx [B, W, 3] = [e3,e2,e1],
 *          t [B, W, 3] = [t3,t2,t1], anony [B, W, 3]; W = n*N2, walk w = i1*N2 + j (:283-305).
 * scanned (optional, [B*W]): number of prefix entries inspected for the neighbour-id filter
 *          in cases 1/2 (the "S" of SURVEY 8(d)); 0 for case 3.
 */
int orc_sample_walks(const orc_graph *g, int64_t B, int n, int N2, const int32_t *root,
                     const int32_t *h1_node, const int32_t *h1_eidx, const float *h1_ts,
                     uint64_t seed, uint64_t row_offset,
                     int32_t *o_nodes, int32_t *o_eidx, float *o_t, int32_t *o_anony, int64_t *scanned) {
    const int64_t W = (int64_t)n * N2;
    for (int64_t b = 0; b < B; ++b)
        if (root[b] < 0 || root[b] >= g->n_nodes) return ORC_ERR_NODE_RANGE;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t r2 = 0; r2 < B * n; ++r2) { /* loop variable i of get_next_step, :323 */
        const int64_t b = r2 / n;
        const int64_t source = root[b], src_id = h1_node[r2];
        const int32_t e1 = h1_eidx[r2];
        /* find_before_walk([source, src_id], e_idx=e1), :149-194 */
        const int64_t c_a = walk_cut_step2(g, source, e1), c_b = walk_cut_step2(g, src_id, e1);
        const int64_t s_a = g->off[source], s_b = g->off[src_id];
        const int64_t L = c_a + c_b;
        uint64_t d2[64]; uint64_t *d = N2 <= 64 ? d2 : malloc(sizeof(uint64_t) * N2);
        if (L > 0) {
            for (int k = 0; k < N2; ++k) d[k] = draw_index(seed, 16, row_offset * n + r2, k, (uint64_t)L);
            sort_u64(d, N2); /* :328 */
        }
        for (int j = 0; j < N2; ++j) {
            int64_t src2 = 0, tgt2 = 0; int32_t e2 = 0; float t2 = 0.f;
            if (L > 0) {
                int64_t p = (int64_t)d[j] < c_a ? s_a + (int64_t)d[j] : s_b + ((int64_t)d[j] - c_a);
                src2 = (int64_t)d[j] < c_a ? source : src_id;     /* :330 */
                tgt2 = g->nbr[p]; e2 = g->eidx[p]; t2 = (float)g->ts[p];
            }
            /* ---- get_final_step, loop variable i = walk row, :353 ---- */
            const int64_t w = r2 * N2 + j;
            const int64_t s1 = source, t1n = src_id, s2 = src2, t2n = tgt2;
            int64_t A, Bn, fa1, fa2, fb; /* lists A, Bn and the ids to keep (-1 = keep all) */
            int code;
            if (s1 == s2 && t1n != t2n) { A = s1; Bn = t2n; fa1 = t1n; fa2 = t2n; fb = t1n; code = 2; }        /* :355-372 */
            else if (t1n == s2 && s1 != t2n) { A = t1n; Bn = t2n; fa1 = s1; fa2 = t2n; fb = s1; code = 3; }    /* :395-412 */
            else { A = t1n; Bn = t2n; fa1 = fa2 = fb = -1; code = 1; }                                           /* :436-449 */
            const int64_t cA = walk_cut_step3(g, A, e2), cB = walk_cut_step3(g, Bn, e2);
            const int64_t sA = g->off[A], sB = g->off[Bn];
            int64_t nA = 0, nB = 0;
            if (fa1 < 0) { nA = cA; nB = cB; }
            else {
                for (int64_t q = 0; q < cA; ++q) nA += (g->nbr[sA + q] == fa1 || g->nbr[sA + q] == fa2);
                for (int64_t q = 0; q < cB; ++q) nB += (g->nbr[sB + q] == fb);
            }
            if (scanned) scanned[w] = fa1 < 0 ? 0 : cA + cB;
            int64_t src3 = 0, tgt3 = 0; int32_t e3 = 0; float t3 = 0.f; int t = 0;
            if (nA + nB > 0) {
                int64_t k = (int64_t)draw_index(seed, 17, row_offset * W + w, 0, (uint64_t)(nA + nB)), p = -1;
                if (k < nA) {
                    src3 = A;
                    if (fa1 < 0) p = sA + k;
                    else for (int64_t q = 0, m = 0; q < cA; ++q)
                        if (g->nbr[sA + q] == fa1 || g->nbr[sA + q] == fa2) { if (m == k) { p = sA + q; break; } ++m; }
                } else {
                    k -= nA; src3 = Bn;
                    if (fa1 < 0) p = sB + k;
                    else for (int64_t q = 0, m = 0; q < cB; ++q)
                        if (g->nbr[sB + q] == fb) { if (m == k) { p = sB + q; break; } ++m; }
                }
                tgt3 = g->nbr[p]; e3 = g->eidx[p]; t3 = (float)g->ts[p];
                if (code == 2) {        /* :386-393 */
                    if (src3 == s1 && tgt3 == t1n) t = 1; else if (src3 == s1 && tgt3 == t2n) t = 2;
                    else if (src3 == t1n && tgt3 == t2n) t = 3; else t = 0;
                } else if (code == 3) { /* :427-434 */
                    if (src3 == t1n && tgt3 == s1) t = 1; else if (src3 == t1n && tgt3 == t2n) t = 3;
                    else if (src3 == t2n && tgt3 == s1) t = 2; else t = 0;
                } else {                /* :464-473 */
                    if (src3 == s1 && tgt3 != t1n) t = 3; else if (src3 == t1n && tgt3 != s1) t = 2;
                    else if (src3 == s1 && tgt3 == t1n) t = 1; else if (src3 == t1n && tgt3 == s1) t = 1; else t = 0;
                }
            }
            int32_t *on = o_nodes + w * 6;
            on[0] = (int32_t)src3; on[1] = (int32_t)tgt3; on[2] = (int32_t)s2; on[3] = (int32_t)t2n; on[4] = (int32_t)s1; on[5] = (int32_t)t1n; /* :303 */
            o_eidx[w * 3 + 0] = e3; o_eidx[w * 3 + 1] = e2; o_eidx[w * 3 + 2] = e1;  /* :304 */
            o_t[w * 3 + 0] = t3; o_t[w * 3 + 1] = t2; o_t[w * 3 + 2] = h1_ts[r2];       /* :305 */
            o_anony[w * 3 + 0] = 1; o_anony[w * 3 + 1] = code; o_anony[w * 3 + 2] = t;  /* :394,435,474 */
        }
        if (d != d2) free(d);
    }
    return ORC_OK;
}

/*
 * Motif classes.  anony = [1, c, t]; two label orders coexist in the reference:
 *  null model (utils/null_model.py:90)            keys 1..12 = 120,121,123,122,130,131,133,132,110,111,112,113
 *  preprocessing (processed/data_preprocess.py:171) ids 0..11 = 121,122,123,120,131,133,132,130,113,112,111,110
 * Returns -1 for a row that is not one of the 12 strings (the reference would raise KeyError).
 */
static int class_null(int c, int t) {
    static const int m2[4] = {1, 2, 4, 3}, m3[4] = {5, 6, 8, 7}, m1[4] = {9, 10, 11, 12};
    if (t < 0 || t > 3) return -1;
    return c == 2 ? m2[t] : c == 3 ? m3[t] : c == 1 ? m1[t] : -1;
}
static int class_prep(int c, int t) {
    static const int m2[4] = {3, 0, 1, 2}, m3[4] = {7, 4, 6, 5}, m1[4] = {11, 10, 9, 8};
    if (t < 0 || t > 3) return -1;
    return c == 2 ? m2[t] : c == 3 ? m3[t] : c == 1 ? m1[t] : -1;
}

/* statistic, utils/null_model.py:75-82: hist[k-1] += 1 for key k in 1..12 */
int orc_class_hist_null(int64_t count, const int32_t *anony, int64_t *hist12) {
    for (int64_t i = 0; i < count; ++i) {
        int k = anony[3 * i] == 1 ? class_null(anony[3 * i + 1], anony[3 * i + 2]) : -1;
        if (k < 0) return ORC_ERR_ARG;
        hist12[k - 1]++;
    }
    return ORC_OK;
}
/* marginal, processed/data_preprocess.py:171-208: category id 0..11 per motif (+ counts per id) */
int orc_class_ids_prep(int64_t count, const int32_t *anony, int32_t *cat, int64_t *hist12) {
    for (int64_t i = 0; i < count; ++i) {
        int k = anony[3 * i] == 1 ? class_prep(anony[3 * i + 1], anony[3 * i + 2]) : -1;
        if (k < 0) return ORC_ERR_ARG;
        cat[i] = k;
        if (hist12) hist12[k]++;
    }
    return ORC_OK;
}

/*
 * new_edge_info, processed/data_preprocess.py:327-343.
 * eidx [B, W, 3] -> out [B, W, 3, 3]: out[b,m,c,p] = #{walks m' : eidx[b,m',p] == eidx[b,m,c]}.
 */
void orc_edge_identity(int64_t B, int64_t W, const int32_t *eidx, double *out) {
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; ++b) {
        const int32_t *cc = eidx + b * W * 3;
        for (int64_t m = 0; m < W; ++m)
            for (int c = 0; c < 3; ++c) {
                const int32_t id = cc[m * 3 + c];
                for (int p = 0; p < 3; ++p) {
                    int64_t cnt = 0;
                    for (int64_t m2 = 0; m2 < W; ++m2) cnt += (cc[m2 * 3 + p] == id);
                    out[((b * W + m) * 3 + c) * 3 + p] = (double)cnt;
                }
            }
    }
}
