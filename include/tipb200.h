/* libtipb200 -- C ABI of the B200-native TIP tri-graph encoder/decoder hot path.
 *
 * This is the drop-in boundary.  The reference (NYXFLOWER/TIP) is pure Python:
 * its operators are torch.nn.Module.forward methods in src/layers.py that bottom out in
 * torch_geometric / torch_scatter / ATen kernels.  Each entry point below replaces the device work
 * of one of those call sites; the reference line it stands in for is cited on every declaration
 * (paths relative to the reference root).  The Python mirror of the reference's module API lives in
 * tip_b200/layers.py and binds these symbols through ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every function returns TIPB_OK (0) or a negative TIPB_ERR_* code; tipb_last_error() gives the
 *     thread-local message.  Nothing throws, exits or allocates: all outputs and all scratch memory
 *     are caller-owned device buffers (sizes from the *_bytes queries, which make no CUDA calls).
 *   - pointers are device pointers on the current device, 16-byte aligned, row-major, dense.
 *     Indices arrive as int64 exactly as the reference's LongTensors; floats are fp32.
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*) with no host synchronisation,
 *     so every call is CUDA-graph capturable.  No floating-point atomics: results are
 *     run-to-run deterministic.
 *   - there is no CPU fallback.
 */
#ifndef TIPB200_H
#define TIPB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TIPB_VERSION 100

#define TIPB_OK 0
#define TIPB_ERR_INVALID_ARGUMENT (-1)
#define TIPB_ERR_CUDA (-2)
#define TIPB_ERR_UNSUPPORTED (-3)

int tipb_version(void);
const char* tipb_last_error(void);

/* ---------------------------------------------------------------- primitives (exported for tests) */
size_t tipb_scan_workspace_bytes(int64_t n);
int tipb_exclusive_scan_i32(const int32_t* in, int32_t* out /* n+1: out[n] = total */, int64_t n, void* ws,
                            size_t ws_bytes, void* stream);
size_t tipb_sort_workspace_bytes(int64_t n);
/* stable LSD radix sort on the low key_bits bits; result in (keys_out, vals_out), inputs clobbered */
int tipb_sort_pairs_u32(uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out, int64_t n,
                        int key_bits, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- typed CSR ("plan")
 * Replaces: the relation range_list layout of src/utils.py:26-32 + the per-call
 * x.index_select / torch_scatter index handling inside MessagePassing.propagate
 * (src/layers.py:78-79,159-160,230).  Built once per graph and cached by the caller.
 * Field order of the plan buffer (int32 arrays unless noted): */
enum {
    TIPB_CSR_COUNTS = 0,   /* [16]  see TIPB_CSR_COUNT_* */
    TIPB_CSR_EID,          /* [entries]   input position of each sorted entry (>= n_edges: reversed copy) */
    TIPB_CSR_OTHER,        /* [entries]   opposite endpoint */
    TIPB_CSR_SEG_PTR,      /* [seg_cap+1] entry offsets of the non-empty (node, relation) segments */
    TIPB_CSR_SEG_NODE,     /* [seg_cap] */
    TIPB_CSR_SEG_REL,      /* [seg_cap] */
    TIPB_CSR_NODE_PTR,     /* [n_nodes+1] segment offsets per node */
    TIPB_CSR_DEG,          /* [n_nodes]   entries per node */
    TIPB_CSR_INV_DEG,      /* [n_nodes]   float: 1 / max(deg, 1)  (torch_scatter mean's clamp) */
    TIPB_CSR_REL_SEG_PTR,  /* [n_rel+1] */
    TIPB_CSR_REL_SEG,      /* [seg_cap]   segment ids listed relation-major */
    TIPB_CSR_NFIELDS
};
enum {
    TIPB_CSR_COUNT_SEGMENTS = 0, /* S: number of non-empty segments */
    TIPB_CSR_COUNT_VALID = 1,    /* entries kept (self loops / out-of-range entries are dropped) */
    TIPB_CSR_COUNT_STATUS = 2,   /* bit0: an index was out of range; bit1: range_list is not cumulative */
    TIPB_CSR_COUNT_REL_MAJOR = 3 /* 1: segments are ordered (relation, node) instead of (node, relation) */
};
size_t tipb_typed_csr_bytes(int64_t n_entries, int64_t n_nodes, int64_t n_rel);
size_t tipb_typed_csr_workspace_bytes(int64_t n_entries, int64_t n_nodes, int64_t n_rel);
int tipb_typed_csr_layout(int64_t n_entries, int64_t n_nodes, int64_t n_rel,
                          int64_t* offsets_bytes /* [TIPB_CSR_NFIELDS] */, int64_t* seg_capacity);
/* edge_type may be NULL when range_list ([n_rel,2], cumulative) is given, and both may be NULL when
 * n_rel == 1.  by_src=0 groups by edge_index[1] (message target), 1 by edge_index[0].
 * doubled=1 lists each edge in both directions (n_entries = 2*n_edges).
 * rel_major=0: segments ordered (node, relation): NODE_PTR gives each node's segment range directly and
 *              REL_SEG_PTR/REL_SEG list the segment ids relation by relation (what the R-GCN kernels need);
 * rel_major=1: segments ordered (relation, node): REL_SEG_PTR gives each relation's segment range directly and
 *              NODE_PTR indexes REL_SEG, which then lists the segment ids node by node (decoder plans: every
 *              relation's entries stay together, so building the plan of a fresh negative sample is local). */
int tipb_typed_csr_build(const int64_t* edge_index /* [2,n_edges] */, const int64_t* edge_type,
                         const int64_t* range_list, int64_t n_edges, int64_t n_nodes, int64_t n_other,
                         int64_t n_rel, int by_src, int doubled, int drop_self_loops, int rel_major, void* plan,
                         size_t plan_bytes, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- R-GCN with basis decomposition
 * Replaces MyRGCNConv.forward / MyRGCNConv2.forward (src/layers.py:76-94, 157-188) and their
 * autograd backward:   out_i = 1/max(deg_i,1) * sum_{e=(j->i), r=type(e)} x_j W_r + x_i root (+ bias),
 *                      W_r = sum_b att[r,b] basis[b].
 * Evaluated as  H[s] = sum_{e in segment s} x_j   (segmented gather-reduce over the typed CSR),
 *               G[i,b,:] = sum_{s in node i} att[rel_s,b] H[s]  and  out_i = inv_deg_i G[i] . basis + x_i root.
 * `g_saved` ([n_nodes, n_bases, f_in], written by fwd) is what bwd needs for d_basis. */
/* the edge pass on its own: out[s,:] = sum_{e in segment s} feat[other_e,:]  (out: [seg_capacity, f]); this is the
 * kernel the roofline in bench.py is quoted on */
int tipb_seg_aggregate(const void* plan, int64_t n_entries, int64_t n_nodes, int64_t n_rel, const float* feat,
                       int64_t n_feat_rows, int f, float* out, void* stream);
size_t tipb_rgcn_workspace_bytes(int64_t n_entries, int64_t n_nodes, int64_t n_rel, int f_in, int f_out, int n_bases);
int tipb_rgcn_fwd(const void* plan_by_dst, int64_t n_entries, int64_t n_nodes, int64_t n_rel, const float* x,
                  const float* basis, const float* att, const float* root, const float* bias /* or NULL */,
                  int f_in, int f_out, int n_bases, int relu_out, float* out, float* g_saved, void* ws,
                  size_t ws_bytes, void* stream);
/* The node contraction G_i = att_i^T H_i of the forward pass runs on tcgen05 (csrc/rgcn_tc.cuh) for f_in in {32, 64} and
 * n_bases in {16, 32}; its barrier waits are bounded: 0 = no protocol fault so far (one blocking device read). */
int tipb_rgcn_tc_status(void);
int tipb_rgcn_bwd(const void* plan_by_src, int64_t n_entries, int64_t n_nodes, int64_t n_rel,
                  const float* inv_deg_dst /* TIPB_CSR_INV_DEG of the by-dst plan */, const float* x,
                  const float* basis, const float* att, const float* root, const float* g_saved,
                  const float* grad_out, const float* out_for_relu /* fwd output when relu_out, else NULL */,
                  int f_in, int f_out, int n_bases, float* d_x, float* d_basis, float* d_att, float* d_root,
                  float* d_bias /* or NULL */, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- P-P GCN (symmetric-normalised SpMM)
 * Replaces torch_geometric GCNConv(cached=True).forward as called from PPEncoder
 * (src/layers.py:386-387, 391-395): out = D^-1/2 (A + I) D^-1/2 x + bias, D = in-degree + 1,
 * existing self loops dropped (add_remaining_self_loops).  The plan is a typed CSR with n_rel = 1
 * built with drop_self_loops = 1; the by-src plan of the same graph drives the backward pass. */
int tipb_gcn_norm(const void* plan, int64_t n_entries, int64_t n_nodes, float* dis /* [n_nodes] deg^-1/2 */,
                  void* stream);
/* out[i] = dis_out[i] * ( sum_{j in N(i)} dis_in[j] x[j] + dis_in[i] x[i] ) + bias ; optional ReLU */
int tipb_gcn_spmm(const void* plan, int64_t n_entries, int64_t n_nodes, const float* dis_out, const float* dis_in,
                  const float* x, const float* bias /* or NULL */, int f, int relu, float* out, void* stream);

/* ---------------------------------------------------------------- P->D hierarchy conv
 * Replaces MyHierarchyConv.forward (src/layers.py:229-242): mean over incoming edges on the rows
 * [n_source, n_source+n_target), times weight[f_in, f_out].  `mean` ([n_target, f_in]) is saved for bwd. */
size_t tipb_hier_workspace_bytes(int64_t n_source, int64_t n_target, int f_in, int f_out);
int tipb_hier_fwd(const void* plan_by_dst, int64_t n_entries, int64_t n_source, int64_t n_target, const float* x,
                  const float* weight, int f_in, int f_out, float* mean, float* out, void* stream);
int tipb_hier_bwd(const void* plan_by_src, int64_t n_entries, int64_t n_source, int64_t n_target,
                  const float* inv_deg_dst, const float* mean, const float* weight, const float* grad_out,
                  int f_in, int f_out, float* d_x /* [n_source+n_target,f_in] */, float* d_weight, void* ws,
                  size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- DistMult decoder
 * Replaces MultiInnerProductDecoder.forward (src/layers.py:590-592):
 *   value_e = sum_k z[i_e,k] z[j_e,k] weight[r_e,k];  score = sigmoid(value) (or value). */
int tipb_decoder_fwd(const float* z, const float* weight, const int64_t* edge_index, const int64_t* edge_type,
                     int64_t n_edges, int64_t n_nodes, int64_t n_rel, int dim, int apply_sigmoid, float* out,
                     void* stream);
/* d_score -> d_z, d_weight for an arbitrary edge list, through a doubled typed CSR of that list */
size_t tipb_decoder_workspace_bytes(int64_t n_edges, int64_t n_nodes, int64_t n_rel, int dim);
int tipb_decoder_bwd(const void* plan_doubled, int64_t n_edges, int64_t n_nodes, int64_t n_rel, const float* z,
                     const float* weight, const float* grad_out /* [n_edges] */, int dim, int apply_sigmoid,
                     float* d_z, float* d_weight, void* ws, size_t ws_bytes, void* stream);
/* Fused TIP loss (src/layers.py:335-340) over one edge set: scores, BCE terms and the analytic
 * gradient in one pass.  sign=+1: positives, -mean(log(s + 1e-13)); sign=-1: negatives,
 * -mean(log(1 - s + 1e-13)).  Outputs: loss_out[0] (scalar), d_z, d_weight = gradients of that term
 * for an upstream gradient of 1 (accumulate=1 adds into d_z/d_weight/loss_out instead of overwriting). */
int tipb_decoder_bce_fused(const void* plan_doubled, int64_t n_edges, int64_t n_nodes, int64_t n_rel, const float* z,
                           const float* weight, int dim, int sign, int accumulate, float* loss_out, float* d_z,
                           float* d_weight, void* ws, size_t ws_bytes, void* stream);
/* The same loss term for a MIRRORED edge set -- the layout to_bidirection/process_edges produce (src/utils.py:17-23):
 * every relation range holds its pairs and then the same pairs with rows swapped.  Every (node, relation) segment of
 * the doubled plan would then list each neighbour twice, so the by-target plan (one listing per directed edge, the
 * plan the R-GCN forward already has) is enough: half the gather work, same result up to summation order.
 * tipb_edges_mirrored sets flag_out[0] = 1 iff edge_index/range_list have that layout. */
int tipb_decoder_bce_fused_mirrored(const void* plan_by_target, int64_t n_edges, int64_t n_nodes, int64_t n_rel,
                                    const float* z, const float* weight, int dim, int sign, int accumulate,
                                    float* loss_out, float* d_z, float* d_weight, void* ws, size_t ws_bytes,
                                    void* stream);
int tipb_edges_mirrored(const int64_t* edge_index, const int64_t* range_list, int64_t n_edges, int64_t n_rel,
                        int32_t* flag_out, void* stream);
/* all n_nodes^2 x n_rel scores (BASELINE.json config 5): out[r,i,j] */
int tipb_decoder_sweep(const float* z, const float* weight, int64_t n_nodes, int64_t n_rel, int dim,
                       int apply_sigmoid, float* out, void* stream);
/* 0 = every tensor-core sweep so far ran its barrier protocol to the end (the waits are bounded: a protocol fault shows
 * up here instead of as a hang); synchronises the device, clears the flag */
int tipb_decoder_sweep_status(void);

/* ---------------------------------------------------------------- typed negative sampling
 * Replaces typed_negative_sampling (src/neg_sampling.py:5-26) bit for bit, including the legacy
 * numpy MT19937 stream behind np.random.choice (the reference seeds it at src/layers.py:14), the retry
 * loop's indexing quirk and the float32 row = perm / num_nodes.
 *   mt_state      [624 key words, position] = 625 uint32 on the device, interchangeable with numpy's
 *                 RandomState.get_state()[1:3]; tipb_neg_sample advances it in place.
 *   stream_words  the untempered MT19937 stream starting at the state's key block:
 *                 [0,624) = the key block itself, then n_new further words (tipb_mt19937_generate).
 *                 It depends on the state only, so the host may produce it ahead of time.
 *   member        per-relation bitmaps of the positive pairs (+ their popcounts) from tipb_neg_bitmap_build.
 *   table         per-relation brackets of the stream offsets (8 int64 per relation: lo, W, L, win_off, f_off, k, pred, order --
 *                 row b's `order` = the relation with the b-th longest window, the launch order of the window scan)
 *                 built ON THE HOST from host copies of range_list and the popcounts (tipb_neg_table_build,
 *                 no CUDA call); totals[0..3] = sum_L, sum_W, highest accepted index a window may touch,
 *                 expected accepted values consumed.
 *   status        STICKY device word, never cleared by the library: every failed call ORs its bits in (bit0:
 *                 stream_words too short; bit1: retry-round table overflow (exact mode); bit2: an offset left its
 *                 bracket -- call again with exact_mode = 1).  A failed call leaves mt_state untouched (it consumed
 *                 nothing), so calls that are not checked on the host (CUDA-graph replay) cannot silently drift: the
 *                 caller reads the word whenever it likes and clears it itself. */
int tipb_mt19937_seed(uint32_t* mt_state, uint32_t seed, void* stream);
int64_t tipb_mt19937_stream_words(int64_t n_new); /* rounds n_new up to the generator's granularity (454) */
int tipb_mt19937_generate(const uint32_t* mt_state, uint32_t* stream_words /* [624 + n_new] */, int64_t n_new,
                          void* stream);
/* The same stream produced by many CTAs (MT19937 jump-ahead, csrc/mt_jump.cu).  The stream after the key block is cut
 * into chunks of chunk_words (a multiple of 454); chunk k >= 1 starts from the window x^(k*chunk_words) mod phi applied
 * to the key block.  tipb_mt19937_jump_polys is HOST-ONLY (no CUDA call): polys_host[k-1] = that polynomial for
 * k = 1..n_polys, 624 uint32 each (bit i of the polynomial = bit i%32 of word i/32); copy it to the device once.
 * windows = device scratch of (n_chunks - 1) * 624 uint32.  Output identical to tipb_mt19937_generate. */
int tipb_mt19937_jump_polys(int64_t chunk_words, int64_t n_polys, uint32_t* polys_host /* [n_polys*624] */);
int64_t tipb_mt19937_chunk_count(int64_t n_new, int64_t chunk_words);
int tipb_mt19937_generate_chunked(const uint32_t* mt_state, uint32_t* stream_words /* [624 + n_new] */, int64_t n_new,
                                  int64_t chunk_words, const uint32_t* polys /* device */, int64_t n_polys,
                                  uint32_t* windows, void* stream);
size_t tipb_neg_bitmap_bytes(int64_t n_nodes, int64_t n_rel);
int tipb_neg_bitmap_build(const int64_t* pos_edge_index, const int64_t* range_list, int64_t n_edges,
                          int64_t n_nodes, int64_t n_rel, uint32_t* member, int32_t* popcount /* [n_rel] */,
                          void* stream);
int tipb_neg_table_build(const int64_t* range_list_host, const int32_t* popcount_host, int64_t n_rel,
                         int64_t n_nodes, double z_sigma, int64_t* table_host /* [n_rel*8] */,
                         int64_t* totals_host /* [4] */);
size_t tipb_neg_sample_workspace_bytes(int64_t n_edges, int64_t n_rel, int64_t n_words, int64_t sum_l,
                                       int64_t sum_w);
int tipb_neg_sample(uint32_t* mt_state /* [625] */, const uint32_t* stream_words, int64_t n_words /* 624 + n_new */,
                    const uint32_t* member, const int64_t* range_list, const int64_t* table /* device copy */,
                    int64_t sum_l, int64_t sum_w, int64_t n_edges, int64_t n_nodes, int64_t n_rel, int exact_mode,
                    int64_t* neg_edge_index /* [2,n_edges], nullable */,
                    uint32_t* neg_packed /* [n_edges] (row << 16 | col), nullable; n_nodes <= 65535 */,
                    int32_t* status, void* ws, size_t ws_bytes, void* stream);

/* Relation-sharded sampler (one process per GPU).  The accepted MT19937 stream is one global sequence:
 * relation r starts where relation r-1 (incl. its retries) stopped.  A rank owns the relations [r_lo, r_hi): it holds
 * only THEIR bitmaps (tipb_neg_bitmap_build_range), scans only their windows, and composes them into ONE table
 *   rank_table[x] = start offset of relation r_hi if relation r_lo starts at lo(r_lo) + x   (x < w_max; codes < 0: failed)
 * (shard_begin).  The ranks exchange these tables (all-gather, a few KB; done by the caller, e.g. ncclAllGather),
 * and shard_end walks the `world` tables to this rank's start offset, materialises the pairs of its relations
 * (outputs indexed from e_lo = first edge of relation r_lo) and advances mt_state exactly as the unsharded call does on
 * every rank.  first_rel = device int32 [world + 1]: first relation of every rank.  Same ws for both calls. */
int tipb_neg_bitmap_build_range(const int64_t* pos_edge_index, const int64_t* range_list, int64_t n_edges,
                                int64_t n_nodes, int64_t n_rel, int64_t r_lo, int64_t r_hi, int64_t e_lo, int64_t e_hi,
                                uint32_t* member_local, int32_t* popcount_local /* [r_hi - r_lo] */, void* stream);
int tipb_neg_sample_shard_begin(const uint32_t* mt_state, const uint32_t* stream_words, int64_t n_words,
                                const uint32_t* member_local, const int64_t* table, int64_t sum_l, int64_t sum_w,
                                int64_t n_edges, int64_t n_nodes, int64_t n_rel, int64_t r_lo, int64_t r_hi,
                                int32_t* rank_table /* [w_max] */, int64_t w_max, void* ws, size_t ws_bytes,
                                void* stream);
int tipb_neg_sample_shard_end(uint32_t* mt_state, const uint32_t* stream_words, int64_t n_words,
                              const uint32_t* member_local, const int64_t* range_list, const int64_t* table,
                              int64_t sum_l, int64_t sum_w, int64_t n_edges, int64_t n_nodes, int64_t n_rel,
                              const int32_t* all_tables /* [world, w_max] */, const int32_t* first_rel, int world,
                              int rank, int64_t w_max, int64_t r_lo, int64_t r_hi, int64_t e_lo, int64_t e_hi,
                              int64_t* neg_local /* [2, e_hi - e_lo], nullable */,
                              uint32_t* packed_local /* [e_hi - e_lo], nullable */, int32_t* status, void* ws,
                              size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- fused pair pass (decoder + loss + gradient)
 * Replaces, for the training step, decoder(pos) / decoder(neg) / the two log-means of src/layers.py:335-340 AND what
 * autograd derives from them, without the typed CSR of the freshly sampled negatives: a pair (i, j, r) is scored once
 * by one thread, the 2T pair-ends of a tile are grouped by node inside the CTA (stable, no float atomics).
 * Usable when z (n_nodes x dim fp32) fits in shared memory: tipb_pair_pass_supported; otherwise use the
 * tipb_decoder_bce_fused path above.
 *   pairs    packed (row << 16 | col), all pairs of relation r contiguous
 *   items    int32 [n_items,4] = (relation, first pair, pair count <= tipb_pair_chunk(), slot | pass << 30), the work items
 *            of BOTH passes (pass 0: positive pairs, 1: negative pairs) in one launch, built by the caller from the
 *            relation ranges (static per graph), largest first; a slot receives one item's partial results
 *   slots    the positive pass's slots come first, each pass numbers its slots relation-major;
 *            rel_slot_ptr = int32 [2 * (n_rel + 1)]: slot ranges per relation of the positive, then the negative pass
 *   pos_weight / neg_weight  multiplicity / number of scored entries: 2/E for the first halves of a mirrored edge set
 *            (src/utils.py:17-23: every undirected pair stands for two directed entries), 1/E for negatives */
int tipb_pair_pass_supported(int64_t n_nodes, int dim);
int64_t tipb_pair_chunk(void);
size_t tipb_pair_workspace_bytes(int64_t n_slots_total, int64_t n_nodes, int dim);
int tipb_pack_half_pairs(const int64_t* edge_index, const int64_t* range_list, int64_t n_edges, int64_t n_rel,
                         int64_t n_nodes, uint32_t* packed /* [n_edges / 2] */, int32_t* status, void* stream);
int tipb_unpack_pairs(const uint32_t* packed, int64_t n, int64_t* edge_index /* [2,n] */, void* stream);
int tipb_pair_bce_pass(const uint32_t* pos_pairs, const uint32_t* neg_pairs, const int32_t* items, int64_t n_items,
                       int64_t n_slots_total, int64_t n_nodes, const float* z, const float* weight, int dim,
                       float pos_weight, float neg_weight, void* ws, size_t ws_bytes, void* stream);
int tipb_pair_bce_finish(const int32_t* rel_slot_ptr, int64_t n_slots_total, int64_t n_nodes, int64_t n_rel, int dim,
                         float* loss_out /* [1] */, float* d_z, float* d_weight, void* ws, size_t ws_bytes,
                         void* stream);

/* ---------------------------------------------------------------- per-relation evaluation
 * Replaces TIP.compute_auprc_auroc_ap_by_et / auprc_auroc_ap (src/layers.py:353-375, src/utils.py:86-93), i.e. 861
 * x three scikit-learn calls on host copies: for relation r the 2k scores [pos_score[start:end], neg_score[start:end]]
 * with targets [1]*k + [0]*k give record[0,r] = area under the precision-recall curve (trapezoid, sklearn.metrics.auc
 * of precision_recall_curve), record[1,r] = roc_auc_score, record[2,r] = average_precision_score; ties are grouped
 * per distinct score as scikit-learn does.  range_list must be the cumulative [start,end) table.  record is a
 * DEVICE array of 3*n_rel doubles; a relation with an empty class gives NaN (scikit-learn raises there). */
size_t tipb_eval_workspace_bytes(int64_t n_edges, int64_t n_rel);
int tipb_eval_auprc_auroc_ap(const float* pos_score, const float* neg_score, const int64_t* range_list, int64_t n_edges,
                             int64_t n_rel, double* record /* [3, n_rel] */, void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------- optimiser step
 * torch.optim.Adam(model.parameters(), lr).step() of tip.py:21-30 (defaults: betas 0.9/0.999, eps 1e-8, no weight
 * decay, no amsgrad) for all parameter tensors in one launch.  params/grads/exp_avg/exp_avg_sq are HOST arrays of
 * n_tensors device pointers (fp32, numel[k] elements each); step_dev is one device float holding the number of steps
 * taken so far, advanced by the call (CUDA-graph capturable: no host state). */
int tipb_adam_max_tensors(void);
int tipb_adam_step(int n_tensors, void* const* params, const void* const* grads, void* const* exp_avg,
                   void* const* exp_avg_sq, const int64_t* numel, double lr, double beta1, double beta2, double eps,
                   float* step_dev, void* stream);

/* ---------------------------------------------------------------- dense pieces + ablation operators
 * Small fp32 products that the reference leaves to torch: GCNConv.lin (`self.lin(x)`, torch_geometric 2.0.1, called at
 * src/layers.py:391-395), NNDecoder's two layers (src/layers.py:618-631), HierEncoder / FMEncoder's
 * `torch.matmul(feat, embed)` (src/layers.py:541, 571).  Row-major, CUDA cores (every product is far below 1 GFLOP).
 *   tipb_gemm       c[m,n] = a[m,k] * (trans_b ? b[n,k]^T : b[k,n]) (+ bias[n]) (+ ReLU)
 *   tipb_gemm_tn    out[m,n] = a[k,m]^T * b[k,n]  (weight gradients; split-K summed in slice order: deterministic)
 *   tipb_relu_grad_colsum   g = gy * (relu_out > 0) and d_bias[n] = sum_m g[m,n] (the ReLU mask and bias gradient of
 *                   `F.relu(conv1(...))` / GCNConv.bias, src/layers.py:391-393); relu_out NULL: no mask, g not written
 *   tipb_transpose  out[cols,rows] = in^T  (lin.weight^T of an identity feature matrix, prepare.py:22-23) */
int tipb_gemm(const float* a, const float* b, const float* bias /* nullable */, int64_t m, int64_t n, int64_t k, int trans_b,
              int relu, float* c, void* stream);
size_t tipb_gemm_tn_workspace_bytes(int64_t m, int64_t n);
int tipb_gemm_tn(const float* a, const float* b, int64_t k, int64_t m, int64_t n, float* out, void* ws, size_t ws_bytes,
                 void* stream);
size_t tipb_relu_grad_colsum_workspace_bytes(int64_t n);
int tipb_relu_grad_colsum(const float* gy, const float* relu_out /* nullable */, int64_t m, int64_t n,
                          float* g /* nullable iff relu_out is */, float* d_bias /* nullable */, void* ws, size_t ws_bytes,
                          void* stream);
int tipb_transpose(const float* in, int64_t rows, int64_t cols, float* out, void* stream);
/* FMEncoder.forward glue (src/layers.py:541-547): out[i] = cat(embed_out[i] / d_norm[i], hier_out[i]) (mode 0, 'cat')
 * or embed_out[i] / d_norm[i] + hier_out[i] (mode 1, 'add'; f_embed == f_hier), and its gradient */
int tipb_drug_input_fwd(const float* embed_out, const float* d_norm, const float* hier_out, int64_t n, int f_embed,
                        int f_hier, int mode, float* out, void* stream);
int tipb_drug_input_bwd(const float* grad, const float* d_norm, int64_t n, int f_embed, int f_hier, int mode,
                        float* d_embed_out, float* d_hier_out, void* stream);
/* out_a = a * *scalar_dev, out_b = b * *scalar_dev in one launch (the incoming loss gradient applied to the d_z / d_weight
 * the fused loss kernels computed in their forward pass); tipb_fill_zero = cudaMemsetAsync (zero rows of a cat) */
int tipb_scale2(const float* a, int64_t n_a, const float* b, int64_t n_b, const float* scalar_dev, float* out_a,
                float* out_b, void* stream);
int tipb_fill_zero(void* ptr, size_t bytes, void* stream);
/* NNDecoder.forward (src/layers.py:618-631): sigmoid( relu(z_i W1) . w1_l2[r] + relu(z_j W2) . w2_l2[r] ).  The caller
 * forms the per-node hidden layers and the tables table_a = relu(z W1) w1_l2^T, table_b = relu(z W2) w2_l2^T
 * ([n_nodes, n_rel], tipb_gemm); the per-edge work is the gather of two scalars.  An out-of-range index gives NaN.
 * Backward: d_table_a / d_table_b from the per-edge output gradient, as segment sums over typed CSRs of the scored edges
 * grouped by (edge_index[0], relation) and by (edge_index[1], relation) -- no float atomics.  g_ws: n_edges floats. */
int tipb_nn_decoder_fwd(const float* table_a, const float* table_b, const int64_t* edge_index, const int64_t* edge_type,
                        int64_t n_edges, int64_t n_nodes, int64_t n_rel, int apply_sigmoid, float* out, void* stream);
int tipb_nn_decoder_bwd(const void* plan_by_src, const void* plan_by_dst, int64_t n_edges, int64_t n_nodes, int64_t n_rel,
                        const float* grad_out, const float* score, int apply_sigmoid, float* d_table_a, float* d_table_b,
                        float* g_ws, void* stream);
/* General sparse drug features (data/utils.py:117-132: identity + mono side-effect columns) as the front-end of
 * `torch.matmul(x_drug, self.embed)` (src/layers.py:541): out[n_nodes, f] = S dense with S = COO entries
 * (row, col, values[eid]) given as a typed CSR (n_rel = 1) over (col -> row) "edges"; the gradient wrt `dense` is the
 * same call on the by-source plan (S^T). */
int tipb_spmm_values(const void* plan, int64_t n_entries, int64_t n_nodes, const float* values, const float* dense, int f,
                     float* out, void* stream);

/* ---------------------------------------------------------------- train / test edge split
 * process_edges (src/utils.py:35-65) and process_prot_edge (data/utils.py:212-229) on the device, bit-exact with the
 * reference's use of the global numpy stream: per raw pair one np.random.binomial(1, p) draw = one 53-bit double =
 * two MT19937 words; kept pairs followed by their mirror images per relation (src/utils.py:17-23), relation labels,
 * cumulative ranges (src/utils.py:26-32); the dropped pairs form the test set the same way.
 *   raw_index   int64 [2, n_raw]: the (row < col) pairs of all relations, concatenated in relation order
 *   raw_ptr     int64 [n_rel + 1] (device): first pair of every relation, raw_ptr[n_rel] = n_raw
 *   mt_state / stream_words / n_words   the numpy-compatible stream, as for tipb_neg_sample; needs 2 * n_raw unread words
 *   qn = exp(log(1 - q)), px2 = (q * qn) / (1 - q) with q = min(p, 1 - p), formed by the caller in double exactly as
 *        numpy's random_binomial_inversion forms them; invert = (p > 0.5)
 *   tipb_edge_split_mask: kept_scan int32 [n_raw + 1] = exclusive scan of the keep flags (kept_scan[n_raw] = number of
 *        train pairs); advances mt_state by 2 * n_raw words.  status: 0 ok, bit 1 = not enough words (state untouched),
 *        bit 2 = a draw would have restarted numpy's inversion loop (U within 2^-53 of 1; state untouched)
 *   tipb_edge_split_emit: train_idx int64 [2, 2K], train_et [2K], train_range [n_rel, 2], test_* likewise (K from the host
 *        read of kept_scan[n_raw]: the one synchronisation of this data-preparation step) */
size_t tipb_edge_split_workspace_bytes(int64_t n_raw);
int tipb_edge_split_mask(uint32_t* mt_state, const uint32_t* stream_words, int64_t n_words, int64_t n_raw, double qn,
                         double px2, int invert, int32_t* kept_scan, int32_t* status, void* ws, size_t ws_bytes,
                         void* stream);
int tipb_edge_split_emit(const int64_t* raw_index, const int64_t* raw_ptr, int64_t n_raw, int64_t n_rel,
                         const int32_t* kept_scan, int64_t n_train_pairs, int64_t* train_idx, int64_t* train_et,
                         int64_t* train_range, int64_t* test_idx, int64_t* test_et, int64_t* test_range, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TIPB200_H */
