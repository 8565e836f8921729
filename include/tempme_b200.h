/*
 * tempme_b200.h -- C ABI of libtempme_b200.so, the B200 (sm_100a) implementation of TempME's
 * data-parallel hot path: time-ordered neighbour lookup, retrospective 3-event motif walk
 * sampling, event anonymisation / motif-class histogram, edge-identity counts and the
 * TimeEncode + motif-encoder scorer.
 *
 * The reference (dharunm236/TempME) is pure Python and has no FFI layer; its "plugin API" is the
 * duck-typed NeighborFinder / TempME surface.  Each entry point below names the reference
 * function it replaces (file:line relative to the reference root); tempme_b200/ (Python, ctypes)
 * mirrors the reference classes on top of these calls, and INTEGRATION.md shows the binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *  - every function returns TM_OK (0) or a negative tm_status; tm_last_error() gives the text
 *    (thread-local); no C++ exception crosses the boundary;
 *  - pointers named d_* are DEVICE pointers on the graph's device, h_* are HOST pointers;
 *  - every launch goes to the cudaStream_t passed as `stream` (0 = legacy default stream) and is
 *    asynchronous; the library never synchronises the device on the hot path;
 *  - graph handles are immutable after creation: concurrent readers are safe;
 *  - a data-dependent failure that the reference reports as an exception (IndexError for an
 *    e_idx that is not in a node's list, utils/graph.py:134-135) is reported through the
 *    optional d_err word: 0 = fine, otherwise 1 + index of the smallest failing row.
 */
#ifndef TEMPME_B200_H
#define TEMPME_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    TM_OK = 0,
    TM_ERR_ARG = -1,         /* bad argument (null pointer, negative size, unsupported fan-out) */
    TM_ERR_CUDA = -2,        /* a CUDA runtime call failed; see tm_last_error() */
    TM_ERR_NOMEM = -3,
    TM_ERR_NODE_RANGE = -4,  /* node id outside [0, n_nodes) in the adjacency entries */
    TM_ERR_EDGE_TABLE = -5,  /* an edge id is negative or occurs in the lists of more than two nodes */
    TM_ERR_UNSUPPORTED = -6
} tm_status;

typedef struct tm_graph tm_graph;
typedef void *tm_stream; /* cudaStream_t */

/* e_idx value meaning "this row has no e_idx: cut by time" (the e_idx_l=None call of the reference) */
#define TM_EIDX_NONE INT32_MIN
/* stage ids of the draw contract (DESIGN.md "RNG") */
#define TM_STAGE_STEP2 16u
#define TM_STAGE_STEP3 17u
#define TM_MAX_STEP2_FANOUT 32

int tm_version(void);
const char *tm_last_error(void);

/* ------------------------------------------------------------------------------------------
 * Graph: replaces NeighborFinder.__init__ / init_off_set / get_ts2idx (utils/graph.py:13-101).
 * The build runs on the device (upload, stable radix sorts by timestamp and node, scan, per-edge table, secondary index); inputs with
 * a (node, edge id) pair twice in one list or negative timestamps take the literal host pass instead (TEMPME_GRAPH_BUILD=host forces it).
 * Input is the flattened adj_list: entry j belongs to the list of node h_entry_node[j]; entries
 * of one node appear in insertion order.  Per node the entries are stably sorted by timestamp
 * (graph.py:48) into a device-resident CSR; nodeedge2idx is materialised as a per-edge table.
 * ------------------------------------------------------------------------------------------ */
int tm_graph_create(int64_t n_nodes, int64_t n_entries, const int32_t *h_entry_node,
                    const int32_t *h_entry_nbr, const int32_t *h_entry_eidx, const double *h_entry_ts,
                    int device, tm_graph **out);
/* Same, from an event list: event k is appended to src's list and then to dst's list
 * (the callers' loop, temp_exp_main.py:135-144 / utils/null_model.py:55-64). */
int tm_graph_create_from_events(int64_t n_nodes, int64_t n_events, const int32_t *h_src, const int32_t *h_dst,
                                const int32_t *h_eidx, const double *h_ts, int device, tm_graph **out);
void tm_graph_destroy(tm_graph *g);
int tm_graph_sizes(const tm_graph *g, int64_t *n_nodes, int64_t *n_entries, int64_t *max_eidx, int64_t *device_bytes);
/* Public attribute surface of NeighborFinder (off_set_l, node_idx_l, edge_idx_l, node_ts_l; graph.py:22-27) */
int tm_graph_export(const tm_graph *g, int64_t *h_off, int32_t *h_nbr, int32_t *h_eidx, double *h_ts);
/* nodeedge2idx as a table: h_tab[e] = {node_a, node_b, cut_a, cut_b} (-1 = absent), e in [0, max_eidx] */
int tm_graph_export_edge_table(const tm_graph *g, int32_t *h_tab);
/* The secondary index of the neighbour-id filter (get_final_step, utils/graph.py:358-371): per node the keys (neighbour << 32 | position)
 * in ascending order, h_skey [n_entries].  For tests of the build (device sort passes == host pass). */
int tm_graph_export_skey(const tm_graph *g, uint64_t *h_skey);

/* find_before (utils/graph.py:103-146) for R rows: window = [d_start[i], d_start[i] + d_cut[i]).
 * d_cut_time may be NULL when every row carries an e_idx; d_eidx may be NULL (all rows cut by time). */
int tm_find_before_batch(const tm_graph *g, int64_t R, const int32_t *d_node, const double *d_cut_time,
                         const int32_t *d_eidx, int64_t *d_start, int32_t *d_cut, int32_t *d_err, tm_stream stream);

/* ------------------------------------------------------------------------------------------
 * Draws.  index = (r64 * L) >> 64 with r64 from Philox4x32-10(key = seed, ctr = (slot >> 1,
 * row_lo, row_hi, stage)), low or high 64-bit half by slot & 1; row = row_offset(...) + local row.
 * If d_inject is non-NULL the draws are taken from it instead (d_inject[row_local * fanout + slot],
 * already reduced to [0, L)): the mode used to replay indices recorded from the reference's own
 * MT19937 stream.
 * ------------------------------------------------------------------------------------------ */

/* get_temporal_neighbor (utils/graph.py:197-231), uniform branch: n sorted draws per non-empty
 * window.  Outputs [R, n]: node (i32), eidx (i32), ts (f32).  stage = hop level. */
int tm_sample_hop(const tm_graph *g, int64_t R, const int32_t *d_node, const double *d_cut_time,
                  const int32_t *d_eidx, int n, uint64_t seed, uint32_t stage, uint64_t row_offset,
                  const uint32_t *d_inject, int32_t *d_o_node, int32_t *d_o_eidx, float *d_o_ts,
                  int32_t *d_err, tm_stream stream);
/* find_k_hop (utils/graph.py:233-262) as one call: hop 0 on the roots, hop l >= 1 on the flattened records of hop l - 1 looked up by e_idx
 * (:247-250).  h_o_node / h_o_eidx / h_o_ts are HOST arrays of k device pointers; hop l has shape [B, n^(l+1)].  Stage l draws use
 * row = row_offset * n^l + i, as k separate tm_sample_hop calls would. */
int tm_sample_khop(const tm_graph *g, int64_t B, int k, int n, const int32_t *d_root, const double *d_cut_time, const int32_t *d_eidx,
                   uint64_t seed, uint64_t row_offset, int32_t *const *h_o_node, int32_t *const *h_o_eidx, float *const *h_o_ts,
                   int32_t *d_err, tm_stream stream);

/* find_k_walks = get_next_step + get_final_step (utils/graph.py:265-476).
 * d_root [B]; d_h1_* [B, n] (first-hop record).  W = n * N2 walks per root, walk w = i1 * N2 + j.
 * Outputs: d_o_nodes [B, W, 6] = [src3,tgt3,src2,tgt2,src1,tgt1]; d_o_eidx [B, W, 3] = [e3,e2,e1];
 * d_o_t [B, W, 3] = [t3,t2,t1]; optional d_o_anony [B, W, 3] = [1, c, t]; optional d_o_cat [B*W]
 * = category id 0..11 in the order of processed/data_preprocess.py:171; optional histograms
 * (accumulated, not zeroed): d_hist_null[12] in the key order of utils/null_model.py:90 and
 * d_hist_prep[12] in the data_preprocess order; optional d_scanned[1] accumulates the number of
 * prefix entries the neighbour-id filter of cases 1/2 had to inspect.
 * row_offset = global index of the first root (RNG rows are row_offset*n + i and row_offset*W + i). */
int tm_sample_walks(const tm_graph *g, int64_t B, int n, int N2, const int32_t *d_root,
                    const int32_t *d_h1_node, const int32_t *d_h1_eidx, const float *d_h1_ts,
                    uint64_t seed, uint64_t row_offset, const uint32_t *d_inject2, const uint32_t *d_inject3,
                    int32_t *d_o_nodes, int32_t *d_o_eidx, float *d_o_t, int32_t *d_o_anony, uint8_t *d_o_cat,
                    unsigned long long *d_hist_null, unsigned long long *d_hist_prep,
                    unsigned long long *d_scanned, tm_stream stream);

/* get_final_step (utils/graph.py:335-476) on its own: R walks whose first two events are given.
 * d_src1/d_tgt1/d_e1/d_t1 [R] = first event (root, first-hop neighbour, e_idx, time); d_step2 [R, 3] = (src2, tgt2, e2),
 * d_t2 [R] (may be NULL).  Draw rows are row_offset + i, stage 17.  Outputs as tm_sample_walks with W = 1. */
int tm_walk_final_step(const tm_graph *g, int64_t R, const int32_t *d_src1, const int32_t *d_tgt1, const int32_t *d_e1,
                       const float *d_t1, const int32_t *d_step2, const float *d_t2, uint64_t seed, uint64_t row_offset,
                       const uint32_t *d_inject3, int32_t *d_o_nodes, int32_t *d_o_eidx, float *d_o_t, int32_t *d_o_anony,
                       tm_stream stream);
/* get_next_step with e_idx_l = None (utils/graph.py:308-333; find_before_walk's bisect branch :170-171): row i draws N2 second events from
 * the prefixes of [d_source[i], d_nbr[i]] cut by TIME at d_cut_time[i] (strict lower bound, float64).  Outputs [R, N2].  Draw contract:
 * stage TM_STAGE_STEP2, row = row_offset + i (the e_idx form inside tm_sample_walks uses the same rows). */
int tm_walk_next_step_time(const tm_graph *g, int64_t R, int N2, const int32_t *d_source, const int32_t *d_nbr, const double *d_cut_time,
                           uint64_t seed, uint64_t row_offset, const uint32_t *d_inject, int32_t *d_o_src, int32_t *d_o_tgt, int32_t *d_o_eidx,
                           float *d_o_ts, int32_t *d_err, tm_stream stream);

/* statistic (utils/null_model.py:75-82) / marginal (processed/data_preprocess.py:148-208) on
 * anonymised rows [count, 3]: accumulates both 12-bin histograms, optionally writes category ids.
 * A row that is none of the 12 classes (KeyError in the reference) sets d_err. */
int tm_class_hist(int64_t count, const int32_t *d_anony, unsigned long long *d_hist_null,
                  unsigned long long *d_hist_prep, uint8_t *d_o_cat, int32_t *d_err, tm_stream stream);

/* new_edge_info (processed/data_preprocess.py:327-343): d_eidx [B, W, 3] -> d_out [B, W, 3, 3] (f32) */
int tm_edge_identity(int64_t B, int64_t W, const int32_t *d_eidx, float *d_out, tm_stream stream);
/* The same counts as bytes (W <= 255), d_out [B, W, 3, 4] u8 = three counts and a pad byte per walk event: the compact form the device
 * pipeline hands to the scorer (tm_encoder_desc.edge_identity_u8) -- a third of the bytes, one 4-byte store / load per event. */
int tm_edge_identity_u8(int64_t B, int64_t W, const int32_t *d_eidx, uint8_t *d_out, tm_stream stream);

/* ------------------------------------------------------------------------------------------
 * Encoder / scorer: TempME.forward (models/explainer.py:174-201) in eval mode.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int32_t node_dim;     /* D  = base.n_feat_th.shape[1] = time_dim   explainer.py:107-109 */
    int32_t edge_dim;     /* Ed = base.e_feat_th.shape[1]              explainer.py:108 */
    int32_t hid_dim;      /* H: 64 or 32 (temp_exp_main.py:40, enhance_main.py:66)   explainer.py:111 */
    int32_t use_temporal; /* TemporalAwareAttention (1) or Attention (0) explainer.py:121 */
    int32_t if_cat;       /* one-hot category features                 explainer.py:116,195-197 */
    int32_t edge_projected; /* 0: d_edge_feat is the base model's edge feature table [rows, Ed].  1: d_edge_feat is the table made by
                             * tm_encoder_project_edges, [rows, D]: lin_event's edge columns already applied per edge id */
    int32_t walk_fanout;  /* layout hint, 0 / 1 = none: find_k_walks numbers the walks of a root w = i1 * N2 + j (utils/graph.py:290-300), so
                           * walk_fanout = N2 consecutive walks share the event next to the root.  The scorer then evaluates that event's
                           * layers (and the products that depend on it alone) once per group; every tile of groups checks the premise
                           * on its own operands and repeats the work per walk where it does not hold, so any value gives the same scores */
    int32_t edge_identity_u8; /* 1: d_edge_identity points to the byte counts of tm_edge_identity_u8 ([B, W, 3, 4] u8) instead of floats */
} tm_encoder_desc;

/* Host pointers to the reference's parameters (nn.Linear layout: weight [out, in] row-major). */
typedef struct {
    const float *lin_event_w, *lin_event_b; /* event_conv.lin_event  [D, Ed+3+D]   explainer.py:82 */
    const float *gcn0_w, *gcn0_b;           /* event_conv.MLP.0      [H, D]        :84 */
    const float *gcn2_w, *gcn2_b;           /* event_conv.MLP.2      [H, H] */
    const float *att_w1_w, *att_w1_b;       /* attention.W1          [2H, 2H]      :772 */
    const float *att_w2_w, *att_w2_b;       /* attention.W2          [2H, 2H]      :773 */
    const float *att_mlp0_w, *att_mlp0_b;   /* attention.MLP.0       [H, 2H]       :776-781 */
    const float *att_mlp3_w, *att_mlp3_b;   /* attention.MLP.3 (.2)  [H, H] */
    const float *mlp0_w, *mlp0_b;           /* MLP.0                 [M, M], M = H + 12 (or H)   :123-125 */
    const float *mlp3_w, *mlp3_b;           /* MLP.3                 [H, M] */
    const float *mlp5_w, *mlp5_b;           /* MLP.5                 [1, H] */
    const float *basis_freq, *phase;        /* time_encoder          [D], [D]      :49-50 */
} tm_encoder_params;

/* Size in floats of the packed device weight blob, and the packer (host -> host blob; the caller
 * uploads it once and keeps it resident). */
int64_t tm_encoder_blob_floats(const tm_encoder_desc *desc);
int tm_encoder_pack(const tm_encoder_desc *desc, const tm_encoder_params *params, float *h_blob);
/* Edge projection: d_out[e][n] = sum_j lin_event.weight[n][j] * d_edge_feat[e][j] (j < Ed), i.e. the part of event_conv.lin_event
 * (explainer.py:93) that depends on the edge id alone, as a table [n_edge_rows, D] in place of one Ed x D product per walk event.  Rebuild it
 * whenever the weights (d_blob) or the feature table change; pass it as d_edge_feat with desc->edge_projected = 1. */
int tm_encoder_project_edges(const tm_encoder_desc *desc, const float *d_blob, const float *d_edge_feat, int64_t n_edge_rows, float *d_out,
                             int device, tm_stream stream);
/* Workspace (floats) tm_encode_score needs for B roots in groups of `group` roots. */
int64_t tm_encoder_workspace_floats(const tm_encoder_desc *desc, int64_t B, int64_t W, int64_t group);

/* Scores for B roots x W walks.  `group` = roots per reference batch: the temporal attention
 * normalises by the unbiased std of |cut_time - t| over a whole batch [group, W, 2]
 * (explainer.py:826-828), so scores depend on the batch a root is scored with.
 * d_nodes [B,W,6] i32, d_eidx [B,W,3] i32, d_t [B,W,3] f32, d_cat [B*W] u8, d_cut_time [B] f32,
 * d_edge_identity [B,W,3,3] f32, feature tables row-major f32.  d_scores [B*W] f32. */
int tm_encode_score(const tm_encoder_desc *desc, const float *d_blob, int64_t B, int64_t W, int64_t group,
                    const int32_t *d_nodes, const int32_t *d_eidx, const float *d_t, const uint8_t *d_cat,
                    const float *d_cut_time, const float *d_edge_identity,
                    const float *d_node_feat, int64_t n_node_rows, const float *d_edge_feat, int64_t n_edge_rows,
                    float *d_workspace, float *d_scores, int device, tm_stream stream);

/* As tm_encode_score, with the score gather of the sharded path (tempme_b200/dist.py; the reference has no multi-GPU path) fused into
 * the kernel: every score is also stored to h_peer_scores[p] + (its index), p < n_peers <= 7 -- device addresses, mapped on this GPU, of this
 * rank's [B*W] segment inside each peer GPU's gathered buffer (NVLink peer stores, 128 bytes per warp).  The stores are visible to the peers
 * when the kernel has completed; order them with the collective that follows on the stream (the histogram all-reduce).  h_peer_scores is a
 * HOST array read during the call. */
int tm_encode_score_gather(const tm_encoder_desc *desc, const float *d_blob, int64_t B, int64_t W, int64_t group,
                           const int32_t *d_nodes, const int32_t *d_eidx, const float *d_t, const uint8_t *d_cat,
                           const float *d_cut_time, const float *d_edge_identity,
                           const float *d_node_feat, int64_t n_node_rows, const float *d_edge_feat, int64_t n_edge_rows,
                           float *d_workspace, float *d_scores, const uint64_t *h_peer_scores, int n_peers, int device, tm_stream stream);

/* As tm_encode_score, and additionally d_y [B*W, hid_dim] = relu(attention.MLP.0(.)) of every walk: the input of attention.MLP.3, i.e. the
 * attention output of TempME.enhance_predict_walks (models/explainer.py:240-243) before its last Linear. */
int tm_encode_attention(const tm_encoder_desc *desc, const float *d_blob, int64_t B, int64_t W, int64_t group,
                        const int32_t *d_nodes, const int32_t *d_eidx, const float *d_t, const uint8_t *d_cat,
                        const float *d_cut_time, const float *d_edge_identity,
                        const float *d_node_feat, int64_t n_node_rows, const float *d_edge_feat, int64_t n_edge_rows,
                        float *d_workspace, float *d_scores, float *d_y, int device, tm_stream stream);

/* ---- enhance path (models/explainer.py:222-306, eval mode).
 * tm_walk_importance: TempME.compute_walk_importance (:257-306) -> d_weights [B,W]; the recency std and the degree mean / std run over
 * each reference batch of `group` roots; d_node_degree [n_nodes] f32 (TempME.node_degree).
 * tm_enhance_reduce: sum over the walks of weight * attention output (:245-249) = attention.MLP.3 applied once per root to the weighted
 * sum of d_y, plus (d_cat != NULL) the 12 per-root class counts (:251-253, :307-313) -> d_out [B, hid_dim (+ 12)]. */
int tm_walk_importance(int64_t B, int64_t W, int64_t group, const float *d_t, const int32_t *d_nodes, const float *d_cut_time,
                       const float *d_node_degree, int64_t n_nodes, float *d_weights, tm_stream stream);
int tm_enhance_reduce(int64_t B, int64_t W, int hid_dim, const float *d_y, const float *d_weights, const float *d_att_mlp3_w,
                      const float *d_att_mlp3_b, const uint8_t *d_cat_or_null, float *d_out, tm_stream stream);

/* ---- TempME.kl_loss, forward value (models/explainer.py:432-453; logged by the reference's eval loops, temp_exp_main.py:326-328 / :459-461).
 * d_prob [B,W] motif scores, d_cat [B,W] walk classes, d_null_values [n_cat] = list(null_model.values()) (paired with the class index by
 * position, as the reference does), empirical != 0 for prior == "empirical".  d_workspace: B doubles.  d_loss: 1 float. */
int tm_kl_loss(int64_t B, int64_t W, const float *d_prob, const uint8_t *d_cat, const float *d_null_values, int n_cat, float target,
               int empirical, double *d_workspace, float *d_loss, tm_stream stream);

/* Kernel timing of the tensor-core scorer (CUDA events around its launch on the caller's
 * stream).  tm_encoder_profile(1) starts collecting; tm_encoder_profile_read returns in h_event_ms the accumulated
 * milliseconds of the scorer kernel since the last read (synchronises on the recorded events); h_motif_ms is 0 since the
 * event-level and the motif-level phases run as one kernel. */
int tm_encoder_profile(int enable);
int tm_encoder_profile_read(float *h_event_ms, float *h_motif_ms);

/* ---- motif -> edge explanation aggregation: TempME.retrieve_edge_imp_node, eval mode (models/explainer.py:354-406).
 * walk_imp[b, 3w + p] = score[b, w] * (0.5 + 0.5 sigmoid(edge_dependency_gcn([edge_feat[e] | TimeEncode(t)])))  (:363-386; the gate is
 * skipped when d_gate_blob is NULL = use_dependency_aware_sampling False); edge_imp = per-root scatter-max of walk_imp over the edge ids
 * of the root's walks (0 for ids no walk carries, :389); gathered to the hop-1 / hop-2 slots (:392-393); E[Beta(max(10 p, 1),
 * max(10 (1 - p), 1))] (:396-397, :421-430 with training = False); slots whose node id is 0 give 0 (:400-404). */
typedef struct {
    int32_t edge_dim, time_dim, hid_dim;     /* time_dim = node_dim (explainer.py:109); hid_dim 64 or 32 */
} tm_gate_desc;
typedef struct {                              /* edge_dependency_gcn (explainer.py:143-151), nn.Linear layout, host pointers */
    const float *w0, *b0;                     /* .0  [H, Ed + D] */
    const float *w3, *b3;                     /* .3  [H/2, H] */
    const float *w6, *b6;                     /* .6  [1, H/2] */
    const float *basis_freq, *phase;          /* time_encoder [D] */
} tm_gate_params;
int64_t tm_gate_blob_floats(const tm_gate_desc *desc);
int tm_gate_pack(const tm_gate_desc *desc, const tm_gate_params *params, float *h_blob);
/* d_scores [B,W] f32, d_eidx [B,W,3] i32, d_t [B,W,3] f32 (walks[1], walks[2]); hop slots d_h{0,1}_node / d_h{0,1}_eidx [B,K0] / [B,K1] i32
 * (subgraph node_records / eidx_records); d_walk_imp [B*W*3] f32 workspace; outputs d_imp0 [B,K0], d_imp1 [B,K1] f32. */
int tm_edge_importance(const tm_gate_desc *desc, const float *d_gate_blob, int64_t B, int64_t W, const float *d_scores,
                       const int32_t *d_eidx, const float *d_t, const float *d_edge_feat, int64_t n_edge_rows,
                       int64_t K0, const int32_t *d_h0_node, const int32_t *d_h0_eidx, int64_t K1, const int32_t *d_h1_node,
                       const int32_t *d_h1_eidx, float *d_walk_imp, float *d_imp0, float *d_imp1, int beta_sample, uint64_t seed, int device,
                       tm_stream stream);

/* ---- Training side of the explainer.
 * tm_beta_sample: TempME.beta_sample(prob, training=True) (models/explainer.py:421-427): d_out[i] ~ Beta(max(10 p_i, 1), max(10 (1 - p_i), 1)),
 * drawn as g1 / (g1 + g2) from two Marsaglia-Tsang gammas on counter-based Philox draws (key = seed, counter = offset + i); slots whose
 * d_node_or_null id is 0 give 0 (:400-404).  d_g1 / d_g2 (optional) receive the gammas for the reparameterised gradient.
 * tm_edge_importance with beta_sample != 0 draws the same way inside the aggregation kernel (the reference's eval loops pass
 * training=args.if_bern, temp_exp_main.py:312-318,446-453).
 * tm_kl_loss_backward: d loss / d prob of tm_kl_loss scaled by *d_grad_out (NULL: 1) -> d_grad_prob [B,W]. */
int tm_beta_sample(int64_t n, const float *d_prob, const int32_t *d_node_or_null, uint64_t seed, uint64_t offset, float *d_out,
                   float *d_g1_or_null, float *d_g2_or_null, tm_stream stream);
int tm_kl_loss_backward(int64_t B, int64_t W, const float *d_prob, const uint8_t *d_cat, const float *d_null_values, int n_cat, float target,
                        int empirical, const float *d_grad_out_or_null, float *d_grad_prob, tm_stream stream);

/* fp32-accurate GEMM on the tensor cores (tcgen05, 3xTF32 split, TMEM accumulators): C[M,N] (+)= A[M,K] . B[N,K]^T (+ bias[N]), row-major
 * with leading dimensions.  The forward / dgrad / wgrad products of the explainer's nn.Linear layers when gradients are requested
 * (models/explainer.py:174-201 under temp_exp_main.py:605-632's loss.backward()); tempme_b200/training.py: TcLinear. */
int tm_gemm_tf32x3(int64_t M, int64_t N, int64_t K, const float *d_A, int64_t lda, const float *d_B, int64_t ldb, float *d_C, int64_t ldc,
                   const float *d_bias_or_null, int accumulate, tm_stream stream);

/* Hardware self-test of the tcgen05/TMEM conventions the scorer relies on: C[128,N] = A[128,K] * B[N,K]^T on the
 * tensor cores (mode 0: one TF32 pass, mode 1: 3xTF32 split accumulation).  K % 8 == 0, N % 16 == 0, N <= 256. */
int tm_selftest_gemm(const float *d_A, const float *d_B, float *d_C, int K, int N, int mode, tm_stream stream);

/* Issue-rate probe of the tensor pipe: groups x 8 back-to-back tcgen05.mma (M = 128, N, K = 8, tf32, A operand in TMEM) with loop-invariant
 * operands on one otherwise idle SM.  d_out[0] = cycles to issue them, d_out[1] = cycles until they have completed (tools/mma_rate.py). */
int tm_selftest_mma_rate(int N, int groups, long long *d_out, tm_stream stream);

/* Self-test of the TimeEncode cosine (reference models/explainer.py:55-58): d_out[i] = cos(d_x[i]) evaluated by the
 * scorer's device routine (exact integer argument reduction; arguments reach 1e8 and beyond). */
int tm_selftest_cos(const float *d_x, float *d_out, int64_t n, tm_stream stream);

/* Self-test of the tensor-map row gather (cp.async.bulk.tensor tile::gather4) the scorer stages feature rows with: 128 row indices,
 * columns [col, col + 32) of a row-major fp32 table [rows, dim] -> d_out = the raw 128 x 32 floats left in shared memory (row r of the
 * staging at floats [32 r, 32 r + 32); with swizzle128 the 16-byte chunk c of row r sits at chunk c ^ (r & 7)). */
int tm_selftest_gather4(const float *d_table, int64_t rows, int dim, const int32_t *d_idx128, int col, int swizzle128, float *d_out, tm_stream stream);

/* Launch counter: number of kernels this library has launched in this process (bench gpu_launches). */
uint64_t tm_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* TEMPME_B200_H */
