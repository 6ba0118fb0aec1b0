/* hcomp_head.h -- C ABI of the B200-native HComP-Net per-node prototype head.
 *
 * This is the drop-in boundary for ONE hot path of harishB97/PIPNet: everything between the
 * backbone output and the loss scalar in `PIPNet.forward` (pipnet/pipnet.py:111-171) and the head's
 * share of `calculate_loss` (pipnet/train.py:852-1341), forward and backward.  Plain pointers and
 * sizes only, no torch types.  Every entry point
 *   - takes DEVICE pointers (unless the name says host) and a CUDA stream (`cudaStream_t` as void*),
 *   - never allocates device memory (workspaces are passed in), never synchronises the stream,
 *   - returns 0 on success or a negative HCOMP_E_* code; `hcomp_last_error()` gives the text.
 * There is no CPU fallback: on a machine without an sm_100 device the compute calls fail.
 *
 * Layout vocabulary (host code in pipnet_b200/layout.py builds the tables):
 *   nodes      N internal tree nodes in `root.nodes_with_children()` order (util/node.py:174-185)
 *   P          total prototypes = sum of P_n; flat prototype axis = nodes concatenated in that order
 *   K          total child logits = sum of C_n; flat logit axis, same order
 *   rows       one row per (view, location): M = V*HW rows of C channels, bf16, channels-last
 *   tiles      the padded prototype axis: 128-column tiles of equal-length node segments
 *              (int32 records of HCOMP_TILE_INTS words:
 *               {S, nseg, umma_n, dz_col, spill_n, spill_col0, spill_dst, 0, node[16], len[16], poff[16]})
 *   spill      nodes whose softmax does not run in the GEMM epilogue: wide nodes (P_n > 64) and nodes moved out of a
 *              nearly empty last tile into spare pad columns.  K1 writes their raw logits to a scratch matrix
 *              zs[V*HW, ldz] (tile record: spill_n columns from tile column spill_col0 -> zs columns spill_dst..);
 *              row kernels inside the same entry points finish them (forward) and produce their dZ (backward)
 *   P_pad      128 * number of tiles
 */
#ifndef HCOMP_HEAD_H
#define HCOMP_HEAD_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HCOMP_ABI_VERSION 6
#define HCOMP_TILE_INTS 56
#define HCOMP_TILE_COLS 128
#define HCOMP_MAX_SEGS 16

#define HCOMP_E_ARG (-1)      /* invalid argument / unsupported shape */
#define HCOMP_E_CUDA (-2)     /* CUDA runtime or driver error */
#define HCOMP_E_DEVICE (-3)   /* not an sm_100 device */

/* Static per-model tables (all DEVICE pointers, int32 unless noted). */
typedef struct hcomp_tables {
  int32_t n_nodes, n_protos, n_cols, n_leaves, n_welems, p_max;
  const int32_t* proto_off;    /* [N+1] start of node n on the flat prototype axis            */
  const int32_t* cls_off;      /* [N+1] start of node n on the flat child-logit axis          */
  const int32_t* wc_off;       /* [N+1] start of node n's [C_n, P_n] classifier weights       */
  const int32_t* proto_node;   /* [P]   node of flat prototype p                              */
  const int32_t* col_node;     /* [K]   node of flat child column k                           */
  const int32_t* welem_col;    /* [n_welems] flat child column of classifier weight element   */
  const int32_t* welem_proto;  /* [n_welems] flat prototype of classifier weight element      */
  const float* child_w;        /* [K]   class-loss weight min(d)/d_c (util/node.py:37-41)     */
  const int32_t* path_off;     /* [L+1] root->leaf path of leaf l (leaves in sorted-name order)*/
  const int32_t* path_col;     /* [..]  flat child column taken at each step of the path      */
  const int8_t* anc;           /* [L,N] child label of leaf l at node n, -1 if not below n     */
  const int32_t* col_nleaves;  /* [K]   number of leaves below child column k (descendant-structured losses) */
} hcomp_tables;

/* Spill nodes of a layout (host-built, pipnet_b200/layout.py); pass NULL when the layout has none. */
typedef struct hcomp_spill {
  int32_t n_spill;            /* number of spill nodes                                                        */
  int32_t ldz;                /* columns of zs (multiple of 4)                                                */
  const int32_t* recs_host;   /* HOST [n_spill,8] {node, P_n, poff, zoff, dz_col, S class (0 = wide), dz_width, 0} */
  float* zs;                  /* DEVICE [V*HW, ldz] raw logits: the forward writes, the backward reads        */
  float* stats;               /* DEVICE [n_wide, V*HW, 2] row max / 1 over row sum of the wide nodes (same lifetime);
                                 NULL without wide nodes                                                      */
} hcomp_spill;

/* Block-activity tables of dZ[M, P_c] for the block-sparse backward GEMMs.  With hierarchical labels an image drives only
 * the nodes on its root-to-leaf path, so most (image, node) blocks of dZ are exactly zero; the calls that build K5's
 * scatter table / align coefficients (hcomp_head_bwd_dz, or hcomp_head_chain_bwd with scat_out) mark every block that can
 * be nonzero and hcomp_head_bwd_dx / _dw skip the unmarked k-blocks (same results: the skipped products are exact zeros).
 *   t1[ceil(M/256)][ld1]  (dX)  row tile of 256 rows x 64 compact columns;  ld1 >= P_c/64
 *   t2[ceil(P_c/256)][ld2] (dW) column tile of 256 compact columns x 64 rows;  ld2 >= ceil(M/64)
 * ld1, ld2 multiples of 8, both tables 8-byte aligned and ZERO before the marking call (hcomp_head_prologue can clear
 * them); pcol[P]: compact dZ column of a flat prototype (inverse of row_map_c). */
typedef struct hcomp_dz_blocks {
  uint8_t* t1; int32_t ld1;
  uint8_t* t2; int32_t ld2;
  const int32_t* pcol;
  /* optional (iact may be NULL): work items of the backward recompute kernel itself.  iact[n_tiles][iact_pitch] bytes,
   * entry (prototype tile, chunk): chunk = 32 locations of an image pair, b * ceil(HW/32) + location / 32; iact_pitch a
   * multiple of 8 and >= 8 * ceil(V_first * ceil(HW/32) / 8) + 8; tile_of_node[N] = prototype tile holding node n's
   * segment or -1 (spill nodes).  hcomp_head_bwd_dz skips items (256 locations x one tile) without a marked chunk: no
   * operand loads, no MMAs, zeros stored.  Same zero-before-marking rule as t1 / t2. */
  uint8_t* iact; int32_t iact_pitch;
  const int32_t* tile_of_node;
  /* != 0: the caller promises that dZ is ONLY read by hcomp_head_bwd_dx / _dw calls that get this same struct (so they
   * skip every unmarked block).  hcomp_head_bwd_dz then does not store zero tiles that no marked block overlaps -- dZ
   * is left uninitialised there.  Leave 0 when anything else reads dZ. */
  int32_t dz_only_read_through_tables;
} hcomp_dz_blocks;

int hcomp_abi_version(void);
const char* hcomp_last_error(void);
int hcomp_num_sms(void);
/* Creates the library's per-device helper objects for the CURRENT device up front: two non-blocking side streams and
 * their fork/join events (independent small kernels of one fused call run beside each other on them; event fork/join
 * on the caller's stream, capturable into a CUDA graph).  Optional: the first call that needs them creates them lazily,
 * which is the only allocation-like side effect a compute entry point can have -- call this once per device (and before
 * stream capture) to keep the compute calls free of it.  The helper objects belong to the device, not to a host
 * thread: drive one device from one host thread at a time. */
int hcomp_init(void);
/* number of kernels this library has launched in this process (bench.py reports it as gpu_launches) */
long long hcomp_launch_count(void);
/* K1 / K5 and the dX / dW GEMMs run as CTA pairs (tcgen05 cta_group::2, M = 256 MMAs) by default; 0 selects the
 * 1-CTA kernels (same results; kept for A/B measurements and tests).  Returns the previous setting. */
int hcomp_set_cta_pair(int on);
/* Riders (narrow spill nodes, see the glossary above) are finished in the tail of the fused K1 / K5 launch of their
 * segment class by default; 0 sends them through the stand-alone row kernels instead (same results; A/B measurements
 * and tests).  Returns the previous setting.  The forward tail uses one grid barrier per device: two fused forward
 * launches that both carry riders must not run CONCURRENTLY on different streams of one device. */
int hcomp_set_rider_fold(int on);
/* Data-parallel training: the dW all-reduce (NCCL, side stream) is meant to overlap the dX GEMM, but a persistent GEMM
 * that owns every SM leaves the collective's CTAs nowhere to run.  n > 0 makes hcomp_head_bwd_dx launch on (SMs - n)
 * SMs.  Returns the previous setting (default 0). */
int hcomp_set_reserved_sms(int n);

/* Operand precision of the projection GEMM (K1 / K5):
 *   HCOMP_PREC_BF16   bf16 operands, fp32 accumulate (<= 2e-2 relative on pooled scores / losses for fp32 inputs)
 *   HCOMP_PREC_FP32X3 fp32-accurate: every fp32 operand is split into three bf16 terms (hi+mid+lo = 24 bits) and the six
 *                     leading cross products are accumulated in fp32 by the same tcgen05 kernel (<= 1e-5 relative, 6x the
 *                     MMA work).  x / wp then point at 3 stacked planes: x[3*M, C] (hcomp_split3_f32) and
 *                     wp[3*P_pad, C] (hcomp_pack_weights_split3); plane 0 alone is the plain bf16 operand. */
#define HCOMP_PREC_BF16 0
#define HCOMP_PREC_FP32X3 1

/* ---- operand preparation -------------------------------------------------------------------- */
/* fp32 add-on kernels (flat [P,C]; reference: nn.Conv2d weights built at pipnet/pipnet.py:1207) ->
 * bf16 tile-padded [P_pad,C]; row_map[P_pad] gives the flat row of each padded row or -1. */
int hcomp_pack_weights(const float* w_flat, const int32_t* row_map, int P_pad, int C, void* wp_bf16, void* stream);
/* channels-last fp32 features -> bf16 rows (ConvNeXt-26 output is NHWC in memory). n % 8 == 0. */
int hcomp_cast_f32_to_bf16(const float* src, void* dst_bf16, long long n, void* stream);
/* fp32 rows -> 3 stacked bf16 planes dst[3][n] (hi, mid, lo) for HCOMP_PREC_FP32X3. */
int hcomp_split3_f32(const float* src, void* dst_bf16_3planes, long long n, void* stream);
/* fp32 add-on kernels -> 3 stacked tile-padded bf16 planes wp3[3][P_pad, C] for HCOMP_PREC_FP32X3. */
int hcomp_pack_weights_split3(const float* w_flat, const int32_t* row_map, int P_pad, int C, void* wp3_bf16, void* stream);
/* NCHW-contiguous features (ResNet) -> bf16 rows [V*HW, C]. */
int hcomp_nchw_to_rows_bf16(const void* src, int src_is_bf16, int V, int C, int HW, void* dst_bf16, void* stream);
/* Backbone hand-off: out rows [V*HW, C] (bf16, channels-last) = gamma[c] * keep[v] * y[row, c] + res[row, c],
 * the tail of the last ConvNeXt block (torchvision CNBlock.forward: layer_scale * block(input), stochastic depth, += input;
 * features/convnext_features.py:18-25, util/args.py:503 `features.7.2`) written straight into the feature matrix the
 * projection kernel reads.  y / res: channels-last rows, fp32 or bf16; keep: per-image stochastic-depth factor or NULL. */
int hcomp_scale_residual_rows_bf16(const void* y, int y_is_bf16, const void* res, int res_is_bf16, const float* gamma,
                                   const float* keep, int V, int C, int HW, void* out_bf16, void* stream);
/* tgt[V,N], desc[V_first,N], n_desc[N] from labels (pipnet/train.py:934-937). ys is int64[V]. */
int hcomp_label_tables(const long long* ys, const hcomp_tables* t, int V, int V_first, int8_t* tgt, uint8_t* desc,
                       int32_t* n_desc, void* stream);

/* ---- K1: projection + softmax + max-pool (+ align loss) -------------------------------------- */
/* Replaces conv1x1 -> /tau -> softmax(dim=1) -> AdaptiveMaxPool2d for every node
 * (pipnet/pipnet.py:124-159) and the align_pf reduction (pipnet/train.py:1063-1069).
 * x: bf16 [V*HW, C]; views [0,V_first) are paired with views [V_first, V) (train: V_first = V/2).
 * outputs_zeroed: bit 0 = pooled_packed[V,P] and align_sum[N] are already cleared (else the call clears them); bit 1 = leave
 * the narrow spill nodes ("riders") to hcomp_pool_classify_fwd (no grid barrier / rider tail in the fused kernel).
 * desc/align_sum may be NULL. */
int hcomp_proj_softmax_pool_fwd(const void* x_bf16, const void* wp_bf16, const int32_t* tiles_host,
                                const int32_t* tiles_dev, int n_tiles, int V, int V_first, int HW, int C, int P,
                                int P_pad, int n_nodes, float tau, int precision, int outputs_zeroed,
                                const uint8_t* desc, unsigned long long* pooled_packed, double* align_sum,
                                const hcomp_spill* spill, void* stream);
/* packed -> pooled fp32 [V,P] + argmax int32 [V,P] (flat h*W+w, first occurrence; pipnet/pipnet.py:24-25);
 * thresh > 0 applies the inference rule pooled < thresh -> 0 (pipnet/pipnet.py:168-169). */
int hcomp_unpack_pool(const unsigned long long* packed, long long n, float thresh, float* pooled, int32_t* argmax,
                      void* stream);
int hcomp_align_finalize(const double* align_sum, const int32_t* n_desc, int N, int HW, float* loss, void* stream);

/* ---- K5-K7: backward -------------------------------------------------------------------------- */
/* dZ = S * (G - sum_p G*S) / tau with G = align gradient + g_pooled scattered at argmax; recomputes the
 * logits tile (same GEMM as K1).  scat_ws: int2[V*P], coef_ws: float[V_first*N] workspaces.
 * g_align (per node upstream gradient), desc may be NULL.  argmax == NULL: scat_ws / coef_ws already hold the tables
 * (hcomp_head_chain_bwd wrote them); g_pooled / pooled are then unused.  blk (may be NULL; with the layout's proto_off
 * [N+1] and proto_node [P] tables): the table-building launch also marks the dZ blocks that can be nonzero.
 * dz: bf16 [V*HW, P_c] on the COMPACT column axis: tile t owns columns [tiles[t][3], + used width), used width =
 * segments * S rounded up to 8; within a segment class the full tiles are contiguous (layout.py builds the table,
 * row_map_c[P_c] maps a compact column to its flat prototype or -1).  Every column of dz is written. */
int hcomp_head_bwd_dz(const void* x_bf16, const void* wp_bf16, const int32_t* tiles_host, const int32_t* tiles_dev,
                      int n_tiles, int V, int V_first, int HW, int C, int P, int P_pad, int P_c, int n_nodes, float tau,
                      int precision, const int32_t* argmax, const float* g_pooled, const float* pooled, float thresh,
                      const uint8_t* desc, const int32_t* n_desc, const float* g_align, void* scat_ws, float* coef_ws,
                      void* dz_bf16, const hcomp_spill* spill, const hcomp_dz_blocks* blk, const int32_t* proto_off,
                      const int32_t* proto_node, void* stream);
/* dX[rows,C] (bf16) = dZ[rows,P_c] * Wpc[P_c,C]  (Wpc: bf16 kernels packed on the compact axis with row_map_c). */
int hcomp_head_bwd_dx(const void* dz_bf16, const void* wpc_bf16, long long rows, int P_c, int C, void* dx_bf16,
                      const hcomp_dz_blocks* blk /* NULL: dense */, void* stream);
/* dW[P,C] (fp32, ACCUMULATED into; clear it first) += dZ^T * X, padding columns dropped via row_map_c. */
int hcomp_head_bwd_dw(const void* dz_bf16, const void* x_bf16, const int32_t* row_map_c, long long rows, int P_c, int C,
                      float* dw_flat, const hcomp_dz_blocks* blk /* NULL: dense */, void* stream);

/* ---- K2: per-node non-negative classifier (pipnet/pipnet.py:1035-1036) ------------------------- */
int hcomp_classifier_fwd(const float* pooled, const float* wc, const float* bias, const hcomp_tables* t, int V,
                         float* out, void* stream);
/* g_pooled (accumulate flag), g_wc[n_welems], g_bias[K] (may be NULL). */
int hcomp_classifier_bwd(const float* g_out, const float* pooled, const float* wc, const hcomp_tables* t, int V,
                         float* g_pooled, int accumulate, float* g_wc, float* g_bias, void* stream);

/* ---- losses (one call forward, one backward) -------------------------------------------------- */
/* Per-node loss terms of calculate_loss for the shipped recipe, on the flat axes:
 *   align  [N] as produced by hcomp_align_finalize (NULL = term off)              pipnet/train.py:1063-1074
 *   tanh   -1/2 sum_views mean_p log(tanh(sum_desc pooled)+eps)                   pipnet/train.py:1076-1087
 *   orth   ||W_rel W_rel^T - I||_F over prototypes with a classifier weight > 1e-3 pipnet/train.py:1136-1151, :1408-1412
 *   class  weighted NLL(log_softmax(log1p(out^m)))   m = `multiplier` (net._multiplier) pipnet/train.py:1153-1163, util/custom_losses.py:22-34
 * Nodes without a descendant in the batch contribute nothing (pipnet/train.py:941-942).
 * stats[4*N] = per-node {align, tanh, orth, class}; total = sum_k weights[k] * sum_n stats[k][n]
 * (weights_host: 4 floats on the HOST, already divided by N); n_correct[N] = argmax accuracy counters.
 * ws: hcomp_head_losses_ws_floats(t) floats, rel: uint8[P]; both are kept for the backward. */
#define HCOMP_LOSS_TANH 1
#define HCOMP_LOSS_ORTH 2
#define HCOMP_LOSS_CLASS 4
#define HCOMP_LOSS_SPARSITY 8   /* class term on log1p(out^multiplier) (default recipe, multiplier = 2) instead of out */
long long hcomp_head_losses_ws_floats(const hcomp_tables* t);
int hcomp_head_losses_fwd(const float* pooled, const float* out, const float* align, const float* w_flat, const float* wc,
                          const int8_t* tgt, const int32_t* n_desc, const hcomp_tables* t, int V, int V_first, int C,
                          int flags, const float* weights_host, float eps, float multiplier, float* total, float* stats,
                          int32_t* n_correct, float* ws, uint8_t* rel, void* stream);
/* g_total: device scalar.  gvec: float[4*N] workspace; on return gvec[0..N) is the gradient w.r.t. align.
 * g_pooled [V,P], g_out [V,K], g_w [P,C] are fully written (zero where a term is off); any may be NULL. */
int hcomp_head_losses_bwd(const float* g_total, const float* out, const float* w_flat, const int8_t* tgt,
                          const int32_t* n_desc, const float* stats, const hcomp_tables* t, int V, int V_first, int C,
                          int flags, const float* weights_host, float eps, float multiplier, const float* ws,
                          const uint8_t* rel, float* gvec, float* g_pooled, float* g_out, float* g_w, void* stream);

/* ---- fused chains (ABI v6): the same arithmetic as the entry points above in four multi-role launches ------------- */
/* The small per-step work between the four large kernels is launch-latency bound (27 launches, 36 us of a 0.31 ms
 * cub27 step); these calls do it in ONE launch each.  Results equal the separate calls (same arithmetic per element;
 * block-level sums may associate differently).
 *
 * hcomp_head_prologue: everything K1 needs --
 *   rows > 0      : hcomp_pack_weights(w_flat, row_map, rows, C, wp_bf16)
 *   packed != NULL: clears pooled_packed[n_packed] and align_sum[n_align] (then call K1 with outputs_zeroed = 1)
 *   ys != NULL    : hcomp_label_tables(ys, t, V, V_first, tgt, desc, n_desc), n_desc needs no clearing.
 *   zero_extra    : one more buffer to clear (16-byte aligned, multiple of 16 bytes), e.g. the hcomp_dz_blocks tables. */
int hcomp_head_prologue(const float* w_flat, const int32_t* row_map, int rows, int C, void* wp_bf16,
                        unsigned long long* packed, long long n_packed, double* align_sum, int n_align, const long long* ys,
                        const hcomp_tables* t, int V, int V_first, int8_t* tgt, uint8_t* desc, int32_t* n_desc,
                        void* zero_extra, long long zero_extra_bytes, void* stream);
/* hcomp_unpack_pool + hcomp_align_finalize (align / align_sum / n_desc may be NULL) + hcomp_classifier_fwd (out / wc may
 * be NULL; the inference threshold applies to the classifier's input like pipnet/pipnet.py:168-170).
 * deferred (may be NULL): the layout's spill record when the forward call was told to leave its narrow spill nodes
 * ("riders") unfinished (outputs_zeroed bit 1 of hcomp_proj_softmax_pool_fwd; only narrow records, at most 4).  One block
 * per (rider, image pair) then runs the rider's softmax / max-pool / align rows from the raw-logit scratch matrix, unpacks
 * that pair's pooled entries of the node and applies the node's classifier; the fused forward kernel loses its grid
 * barrier and rider tail.  Needs V_first, tau, desc (may be NULL without the align term) and `counter`: DEVICE uint32[5],
 * zero on entry, left zero (word 1 + i: last-block-done of rider i). */
int hcomp_pool_classify_fwd(const unsigned long long* packed, double* align_sum, const int32_t* n_desc, const float* wc,
                            const float* bias, const hcomp_tables* t, int V, int HW, float thresh, float* pooled,
                            int32_t* argmax, float* align, float* out, const hcomp_spill* deferred, int V_first, float tau,
                            const uint8_t* desc, unsigned int* counter, void* stream);
/* Workspace of the chained losses: hcomp_head_losses_ws_floats(t) + V * N floats (row log-sum-exp of the class term). */
long long hcomp_head_chain_ws_floats(const hcomp_tables* t, int V);
/* The weights-only part of the orth term (Gram matrices E_n = W_rel W_rel^T - I and the relevance mask; ||E_n||^2 is a
 * role of the chain kernel) into ws / rel; may run on ANY stream before hcomp_head_chain_fwd (e.g. beside the forward
 * finish) -- then pass HCOMP_LOSS_ORTH_READY there. */
#define HCOMP_LOSS_ORTH_READY 16
int hcomp_orth_gram(const float* w_flat, const float* wc, const hcomp_tables* t, int C, float* ws, uint8_t* rel, void* stream);
/* hcomp_head_losses_fwd in one launch (plus two for the orth term unless HCOMP_LOSS_ORTH_READY).  counter: one DEVICE
 * uint32 that is zero on entry and left zero (last-block-done combine); one call at a time per counter. */
int hcomp_head_chain_fwd(const float* pooled, const float* out, const float* align, const float* w_flat, const float* wc,
                         const int8_t* tgt, const int32_t* n_desc, const hcomp_tables* t, int V, int V_first, int C,
                         int flags, const float* weights_host, float eps, float multiplier, float* total, float* stats,
                         int32_t* n_correct, float* ws, uint8_t* rel, unsigned int* counter, void* stream);
/* Backward of hcomp_head_chain_fwd CHAINED THROUGH THE CLASSIFIER (out = classifier(pooled, wc, bias)): one launch gives
 * g_pooled [V,P] (tanh term + class term through relu(Wc)), g_wc [n_welems], g_bias [K], g_align [N]; the orth term's
 * g_w [P,C] runs beside it.  d loss / d out is never materialised.  Any output may be NULL.
 * scat_out / coef_out (optional; with argmax [V,P], thresh and desc, HW of the forward): the same launch also writes the
 * scatter table int2[V*P] and the align coefficients float[V_first*N] that hcomp_head_bwd_dz builds from g_pooled /
 * g_align; pass them to hcomp_head_bwd_dz as scat_ws / coef_ws with argmax = NULL ("tables are ready") -- valid only when
 * exactly this g_pooled and g_align reach the head backward (no other gradient was added in between). */
int hcomp_head_chain_bwd(const float* g_total, const float* pooled, const float* out, const float* w_flat, const float* wc,
                         const int8_t* tgt, const int32_t* n_desc, const float* stats, const hcomp_tables* t, int V,
                         int V_first, int C, int flags, const float* weights_host, float eps, float multiplier,
                         const float* ws, const uint8_t* rel, float* g_pooled, float* g_wc, float* g_bias, float* g_align,
                         float* g_w, const int32_t* argmax, float thresh, const uint8_t* desc, int HW, void* scat_out,
                         float* coef_out, const hcomp_dz_blocks* blk /* marks for scat_out / coef_out, may be NULL */,
                         void* stream);

/* ---- descendant-structured loss terms switched on by the shipped scripts ------------------------ */
/* (run_pipnet_20protos_multi_runs_seed42.sh: --tanh_desc "y|0.05" --minimize_contrasting_set 'y'
 *  --mask_prune_overspecific 'y|0|1.1'), batched over all nodes on the flat pooled table [V,P]:
 *   TANH_DESC   per leaf below a node: -1/2 sum_views mean_{p in R(child)} log(tanh(sum_{rows of the leaf} pooled)+eps),
 *               R(child) = classifier row > 1e-3; mean over the node's leaves (absent leaves give log(eps))
 *                                                                                  pipnet/train.py:1089-1133
 *   CONTRAST    per child: max over the node's descendants that are NOT below the child of the child's prototypes
 *               (classifier row > 1e-5); mean over all (child, prototype) entries; TOPK = 1   pipnet/train.py:1017-1060
 *   MASK_PRUNE  per child with a leaf in the batch: presence <- softmax((presence + gumbel) / tau) (soft Gumbel softmax,
 *               applied to the previous child's OUTPUT, :978); overspecificity = -sum_{p in R} score[p] * presence[p,1]
 *               with score = prod over the child's leaves in the batch of clamp(max_rows pooled * boost, max=1)
 *               (boost <= 0: plain product; GEOMETRIC: prod of max^(1/#leaves); SG_SCORE: no gradient through the
 *               score); mask_l1 = sum_{p in R} presence[p,1]; both / total relevant prototypes      pipnet/train.py:946-1015
 * Nodes without a descendant in the batch contribute nothing (:941-942).
 * ys: int64[V] leaf index per row (sorted leaf-name order); presence: float[P,2]; gumbel: float[n_welems,2] noise
 * -log(Exp(1)) indexed like the classifier weights (node, child, prototype); weights_host: 4 HOST floats
 * {tanh_desc_weight/N, contrast_weight/N, 2.0/N, 0.5/N}.
 * stats[4*N]: per node {tanh_desc, contrast (unweighted means), overspecificity, mask_l1 (weighted, as the reference
 * stores them :1006-1010)}; loss[1] = sum_n w0*stats0 + w1*stats1 + stats2 + stats3.
 * ws: hcomp_desc_losses_ws_bytes(t, V) bytes, kept for the backward.  Halves for TANH_DESC are rows [0,V_first) and
 * [V_first,V): equal to the reference's .chunk(2) of a leaf's rows for two-view batches ys = cat[ys, ys]. */
#define HCOMP_DESC_TANH_DESC 1
#define HCOMP_DESC_CONTRAST 2
#define HCOMP_DESC_MASK_PRUNE 4
#define HCOMP_DESC_GEOMETRIC 8
#define HCOMP_DESC_SG_SCORE 16
long long hcomp_desc_losses_ws_bytes(const hcomp_tables* t, int V);
int hcomp_desc_losses_fwd(const float* pooled, const float* wc, const float* presence, const float* gumbel,
                          const long long* ys, const int8_t* tgt, const int32_t* n_desc, const hcomp_tables* t, int V,
                          int V_first, int flags, const float* weights_host, float eps, float boost, float gumbel_tau,
                          void* ws, float* stats, float* loss, void* stream);
/* g_loss: device scalar.  g_pooled [V,P] and g_presence [P,2] are fully written; either may be NULL. */
int hcomp_desc_losses_bwd(const float* g_loss, const float* pooled, const float* wc, const float* presence,
                          const float* gumbel, const long long* ys, const int8_t* tgt, const int32_t* n_desc,
                          const hcomp_tables* t, int V, int V_first, int flags, const float* weights_host, float eps,
                          float boost, float gumbel_tau, const void* ws, float* g_pooled, float* g_presence, void* stream);

/* ---- predictions (util/node.py:300-385, pipnet/pipnet.py:173-185) ------------------------------ */
/* probs_ws: float[V*K]; joint: float[V*L] (columns in sorted leaf-name order); pred: int64[V] argmax.
 * prob_override: float[K] or NULL; a node whose first column holds a value >= 0 uses these child probabilities for every
 * sample instead of the softmax (leave_out_classes util/node.py:319-323, fully masked class under
 * apply_overspecificity_mask :335-359; pipnet_b200/pipnet.py builds the table). */
int hcomp_joint_leaf(const float* out, const hcomp_tables* t, int V, float tau, const float* prob_override,
                     float* probs_ws, float* joint, long long* pred, void* stream);

/* ---- visualisation feed (util/vis_hpipnet.py:62-127, :184-290) ---------------------------------- */
/* Top-k activations per (prototype, leaf) over a pass through the data, from the pooled scores / argmax of the fused
 * forward (replaces save_images_topk's per-node, batch-1 loop with Python heaps).  Tables [P, L, k]: t_score descending,
 * t_img = caller's image id (-1 = empty slot; initialise to -1), t_loc = flat argmax location h*W+w.  An image enters the
 * list of (p, its leaf) if the leaf is below p's node, p has a relevant class (classifier column > 1e-3) and the image's
 * child class at that node is relevant to p (or is NOT, with find_non_descendants).  k <= 32.  ws_2v: int32[2*V]. */
int hcomp_topk_update(const float* pooled, const int32_t* argmax, const long long* ys, const long long* img_ids,
                      const float* wc, const hcomp_tables* t, int V, int k, int find_non_descendants, int32_t* ws_2v,
                      float* t_score, long long* t_img, int32_t* t_loc, void* stream);
/* full softmax map of ONE node, fp32 [V, P_n, HW]; w_node: fp32 [P_n, C]. */
int hcomp_materialize_map(const void* x_bf16, const float* w_node, int V, int HW, int C, int P_n, float tau, float* map,
                          void* stream);

/* ---- self-test hook: plain tcgen05 GEMM used by tests (D = A*B, bf16 in, fp32/bf16 out) --------- */
/* a_mn/b_mn: operand stored with its M (resp. N) axis contiguous.  out_mode 0 bf16, 1 fp32, 2 fp32 red.add. */
int hcomp_gemm_bf16(const void* a, const void* b, int M, int N, int K, int a_mn, int b_mn, int out_mode, int splits,
                    void* out, long long ldo, void* stream);

/* ---- data-parallel gradient exchange (reference: DistributedDataParallel's mean of the per-rank
 *      gradients, main_dist.py:330) ---------------------------------------------------------------------------- */
/* In-place MEAN all-reduce of one flat fp32 buffer that every rank of one NVSwitch box allocated symmetrically (same
 * size, mapped into every peer; torch.distributed._symmetric_memory or cuMem* + cuMulticast*).  `ctas` CTAs (the SMs
 * the concurrent dX GEMM leaves free, hcomp_set_reserved_sms) run: cross-rank barrier -> each rank reduces its 1/world
 * shard (mc != NULL: in the switch, multimem.ld_reduce at the multicast address `mc`; else peer loads through
 * `peers_dev`, a DEVICE array of `world` buffer pointers) and scales it by 1/world -> broadcasts it (multimem.st / peer
 * stores) -> cross-rank barrier.  `pads_dev`: DEVICE array of `world` pointers to the ranks' zero-initialised 32-bit
 * signal pads; this call uses channels [channel_base, channel_base + ctas), i.e. words
 * [channel_base * world, (channel_base + ctas) * world) of every pad, and leaves them zero.  Every rank must make the
 * same call (same n, ctas, channel_base) in the same stream order; the buffer must not be written by the caller between
 * the producers' completion and the end of this call.  n: floats, a multiple of 4. */
int hcomp_allreduce_mean_symm(float* local, float* mc, const void* peers_dev, const void* pads_dev, int rank, int world,
                              long long n, int channel_base, int ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HCOMP_HEAD_H */
