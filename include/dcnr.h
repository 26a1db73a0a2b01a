/* dcnr.h -- C ABI of libdcnr_sm100a.so, the B200-native DCN-R hot path.
 *
 * The reference (pure Python) has no FFI; its boundary for this path is two Python object
 * protocols used at fixed call sites:
 *   - DCN_RecSys(...).forward / autograd backward      train.py:125-170, :223-225   main.py:93-127, :320-322
 *   - NearestNeighbors(metric='cosine').fit/kneighbors main.py:268-269, :200, :300
 * Every entry point below is what a ctypes binding on the reference side would call for one of
 * those lines; the line(s) each one replaces are cited on the declaration.  INTEGRATION.md shows
 * the reference-side stub.
 *
 * Conventions
 *   - every function returns 0 (DCNR_OK) or a negative dcnr_status; the message for the last
 *     failure on the calling thread is dcnr_last_error_string().  Nothing aborts or throws.
 *   - all pointers are DEVICE pointers unless the parameter name ends in _host; the caller owns every
 *     buffer (outputs and workspaces included); the library keeps no mutable global state.
 *   - stream is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *     unless documented.  Entry points are re-entrant.
 *   - matrices are row-major fp32; "ld" arguments are row strides in elements.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     DCNR_ERR_CUDA.
 */
#ifndef DCNR_H_
#define DCNR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCNR_ABI_VERSION 4
#define DCNR_MAX_CAT 8     /* categorical tables (reference uses 2: city, hotel_type; train.py:290) */
#define DCNR_MAX_RES 8     /* ResBlocks          (search space 1..4; train.py:183) */
#define DCNR_MAX_CROSS 8   /* CrossLayers        (search space 1..6; train.py:182) */
#define DCNR_PAD 32        /* internal row padding of x0 / dx0, in floats */

typedef enum dcnr_status {
    DCNR_OK = 0,
    DCNR_ERR_INVALID = -1,      /* bad argument / unsupported shape */
    DCNR_ERR_CUDA = -2,         /* a CUDA runtime call or launch failed (message has the code) */
    DCNR_ERR_WORKSPACE = -3,    /* workspace too small */
    DCNR_ERR_INDEX = -4         /* an embedding id was out of range (only from dcnr_check_ids) */
} dcnr_status;

typedef enum dcnr_precision {
    DCNR_PREC_FP32 = 0,         /* CUDA-core FMA GEMMs, IEEE fp32 (parity path) */
    DCNR_PREC_TF32X3 = 1,       /* tcgen05 kind::tf32, 3-term error-compensated split, fp32 accumulate */
    DCNR_PREC_TF32 = 2,         /* tcgen05 kind::tf32 single pass (stated tolerance, not parity) */
    DCNR_PREC_BF16 = 3,         /* eval: fused tower on tcgen05 kind::f16 with bf16 operands, fp32 accumulate (stated
                                 * tolerance, never the default); training and shapes the fused tower does not take run
                                 * DCNR_PREC_TF32 */
    DCNR_PREC_FP16X3 = 4        /* eval: fused tower on tcgen05 kind::f16 with a 3-term error-compensated fp16 split (hi.hi +
                                 * lo.hi + hi.lo, fp32 accumulate) -- fp32 parity at twice the tf32x3 tensor rate; training and
                                 * shapes the fused tower does not take run DCNR_PREC_TF32X3 */
} dcnr_precision;

typedef void *dcnr_stream_t;

int dcnr_abi_version(void);
const char *dcnr_last_error_string(void);
/* Number of kernels this library has launched (process-wide) since the last reset
 * (bench.py's gpu_launches). */
int64_t dcnr_launch_count(int reset);
/* Device-side timing of the dense-layer GEMM launches (measurement aid, off by default): between _begin and _end every
 * GEMM launch is bracketed by a CUDA event pair on its stream; _end waits for them and returns the summed duration, the
 * number of launches and their algorithmic flops (2 m n k each).  bench.py reports roofline.achieved from it. */
int dcnr_gemm_timing_begin(void);
int dcnr_gemm_timing_end(double *total_ms, int64_t *launches, double *total_flops);

/* ------------------------------------------------------------------------------------------
 * Model description: shapes of DCN_RecSys(n_users, n_items, cat_dims, n_num_features, params)
 * (train.py:126-153).  in_dim = 2*emb_dim + sum(cat_width) + n_num; in_dim_pad = round_up(in_dim, 32).
 * ------------------------------------------------------------------------------------------ */
typedef struct dcnr_dims {
    int32_t emb_dim, n_cat, n_num, hidden, n_cross, n_res;
    int32_t in_dim, in_dim_pad;
    int64_t n_users, n_items;
    int64_t cat_rows[DCNR_MAX_CAT];
    int32_t cat_width[DCNR_MAX_CAT];
    float dropout_p;            /* nn.Dropout(p) inside each ResBlock (train.py:108) */
    float bn_eps, bn_momentum;  /* nn.BatchNorm1d defaults 1e-5 / 0.1 */
    int32_t precision;          /* dcnr_precision for the dense layers */
    int32_t dp_sparse_tables;   /* with comm: 1 = build the user / item table gradients from the all-gathered (id, gradient
                                 * row) pairs of ALL ranks (identical dense gradients on every rank, single-device summation
                                 * order, ~150 B per sample on the wire); 0 = local gradients, all-reduce them yourself */
    int64_t dp_batch_cap;       /* with comm and dp_sparse_tables: the largest LOCAL batch of any rank in this step (equal on all
                                 * ranks).  The (id, gradient row) all-gather moves dp_batch_cap rows per rank -- shorter local
                                 * batches are padded with zero rows -- so ranks may hold different batch sizes (the short last
                                 * batch of train.py:196's DataLoader).  0 = every rank has exactly `batch` rows */
    uint64_t *dropout_step;     /* optional DEVICE counter: its value is added to dropout_seed and every dcnr_forward_train
                                 * increments it, so a captured CUDA graph of the training step draws a fresh dropout mask
                                 * on every replay (NULL: the seed argument alone decides the mask) */
    int32_t *eval_flags;        /* optional DEVICE int, OR-ed by dcnr_forward_eval and dcnr_forward_train (never cleared): bit 0 = an embedding id was out of
                                 * range (the row was read as row 0; torch raises IndexError), bit 1 = an activation left the
                                 * fp16 range of DCNR_PREC_FP16X3 (re-run the batch with DCNR_PREC_TF32X3).  NULL: not reported */
    const void *tower_pack;     /* optional DEVICE buffer written by dcnr_tower_prepare for THESE parameters and this precision:
                                 * dcnr_forward_eval then skips its per-call weight preparation (one launch less per call; it
                                 * matters for single requests).  NULL: prepared in the workspace on every call.  The caller
                                 * re-prepares after any parameter or running-statistics update */
    void *comm;                 /* data-parallel group (dcnr_comm_create) or NULL.  When set, train-mode BatchNorm
                                 * statistics and the BatchNorm backward reductions cover the batches of ALL ranks, so an
                                 * N-rank step equals the reference's single-device step on the concatenated batch */
} dcnr_dims;

/* Parameters in the reference's state_dict layouts, fp32, device memory. */
typedef struct dcnr_params {
    const float *user_table;                 /* user_embedding.weight  [n_users, E] */
    const float *item_table;                 /* item_embedding.weight  [n_items, E] */
    const float *cat_table[DCNR_MAX_CAT];    /* cat_embeddings.i.weight [rows_i, width_i] */
    const float *w0, *b0;                    /* initial_deep_layer.{weight [H,D], bias [H]} */
    const float *res_w1[DCNR_MAX_RES], *res_b1[DCNR_MAX_RES];   /* res_blocks.r.layer1 [H,H],[H] */
    const float *res_g1[DCNR_MAX_RES], *res_be1[DCNR_MAX_RES];  /* res_blocks.r.bn1.{weight,bias} */
    float *res_rm1[DCNR_MAX_RES], *res_rv1[DCNR_MAX_RES];       /* bn1.running_{mean,var} (updated in train) */
    int64_t *res_nbt1[DCNR_MAX_RES];                            /* bn1.num_batches_tracked */
    const float *res_w2[DCNR_MAX_RES], *res_b2[DCNR_MAX_RES];
    const float *res_g2[DCNR_MAX_RES], *res_be2[DCNR_MAX_RES];
    float *res_rm2[DCNR_MAX_RES], *res_rv2[DCNR_MAX_RES];
    int64_t *res_nbt2[DCNR_MAX_RES];
    const float *cross_w[DCNR_MAX_CROSS];    /* cross_network.l.w.weight [1,D] */
    const float *cross_b[DCNR_MAX_CROSS];    /* cross_network.l.b        [D]   */
    const float *wf, *bf;                    /* final_linear.{weight [1,H+D], bias [1]} */
} dcnr_params;

/* Gradients, same shapes as the parameters (embedding-table gradients are DENSE like
 * nn.Embedding(sparse=False); the library zero-fills them).  Any pointer may be NULL = not wanted. */
typedef struct dcnr_grads {
    float *user_table, *item_table, *cat_table[DCNR_MAX_CAT];
    float *w0, *b0;
    float *res_w1[DCNR_MAX_RES], *res_b1[DCNR_MAX_RES], *res_g1[DCNR_MAX_RES], *res_be1[DCNR_MAX_RES];
    float *res_w2[DCNR_MAX_RES], *res_b2[DCNR_MAX_RES], *res_g2[DCNR_MAX_RES], *res_be2[DCNR_MAX_RES];
    float *cross_w[DCNR_MAX_CROSS], *cross_b[DCNR_MAX_CROSS];
    float *wf, *bf;
} dcnr_grads;

/* One batch in the tensor schema of prepare_data / preprocess_for_ranking
 * (train.py:69-78, main.py:221-230). */
typedef struct dcnr_batch {
    const int64_t *user_ids;      /* [B] */
    const int64_t *item_ids;      /* [B] */
    const int64_t *cat_features;  /* [B, n_cat] row-major */
    const float *num_features;    /* [B, n_num] */
    int64_t batch;
} dcnr_batch;

/* ------------------------------------------------------------------------------------------
 * Whole-model entry points (what DCN_RecSys.forward / loss.backward() dispatch to).
 * ------------------------------------------------------------------------------------------ */

/* Bytes of workspace for a batch of `batch` rows.  kind: 0 = eval forward, 1 = train forward
 * (holds everything saved for backward), 2 = backward scratch.  Needs a CUDA device. */
int64_t dcnr_workspace_bytes(const dcnr_dims *dims, int64_t batch, int kind);

/* eval()/no_grad forward: running-stat BatchNorm folded into the GEMM epilogues, dropout identity.
 * Replaces main.py:320-322 (ranking call) and train.py:229-233 (validation forward).
 * logits: [batch] fp32. */
int dcnr_forward_eval(const dcnr_dims *dims, const dcnr_params *params, const dcnr_batch *batch,
                      float *logits, void *workspace, int64_t workspace_bytes, dcnr_stream_t stream);

/* train() forward: batch-statistic BatchNorm (running stats + num_batches_tracked updated),
 * dropout with a Philox keep-mask derived from (seed, row, column) -- or, when drop_keep_mask is
 * non-NULL, an injected {0,1} uint8 mask [n_res, batch, hidden] (parity tests).
 * Everything backward needs is left in `saved` (dcnr_workspace_bytes(kind=1)).
 * Replaces train.py:223. */
int dcnr_forward_train(const dcnr_dims *dims, const dcnr_params *params, const dcnr_batch *batch,
                       uint64_t dropout_seed, const uint8_t *drop_keep_mask, float *logits, void *saved,
                       int64_t saved_bytes, dcnr_stream_t stream);

/* Backward of dcnr_forward_train given dL/dlogits [batch].  Replaces train.py:225 (loss.backward()).
 * Deterministic: every reduction has a fixed order (sorted-segment embedding scatter, fixed-chunk
 * column sums, split-K in chunk order). */
int dcnr_backward(const dcnr_dims *dims, const dcnr_params *params, const dcnr_batch *batch,
                  const float *grad_logits, const void *saved, int64_t saved_bytes, const dcnr_grads *grads,
                  void *scratch, int64_t scratch_bytes, dcnr_stream_t stream);

/* nn.BCEWithLogitsLoss() (mean) forward + dL/dlogits (train.py:206,224).  loss: 1 float;
 * grad_logits may be NULL.  scratch: at least 4096 floats. */
int dcnr_bce_with_logits(const float *logits, const float *labels, int64_t batch, float *loss,
                         float *grad_logits, float *scratch, dcnr_stream_t stream);

/* One fused dense Adam (decoupled_weight_decay = 0) / AdamW (= 1) update of a flat fp32 tensor,
 * torch.optim semantics (train.py:201-204, :226); step counts from 1.  The hyper-parameters are doubles (Python floats)
 * so that 1 - beta, lr / bias_correction are formed exactly like torch forms them.   */
int dcnr_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, double lr,
                   double beta1, double beta2, double eps, double weight_decay, int decoupled_weight_decay,
                   int64_t step, dcnr_stream_t stream);

/* Sets *flag_host != 0 (and returns DCNR_ERR_INDEX) if any id is outside its table
 * (torch raises IndexError there).  Synchronises the stream.  err_flag: 1 device int. */
int dcnr_check_ids(const dcnr_dims *dims, const dcnr_batch *batch, int32_t *err_flag, dcnr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Operator-level entry points (each is one stage of the path; the whole-model calls chain them).
 * ------------------------------------------------------------------------------------------ */

/* x0[b] = [U[u_b] | I[i_b] | C0[c_b0] | ... | num_b | 0-pad]   -- train.py:156-159 / main.py:116-119.
 * x0: [batch, ldx0] with ldx0 >= in_dim; columns in_dim..ldx0-1 are written as zero when
 * ldx0 == in_dim_pad.  Bit-exact copy of the table rows. */
int dcnr_embed_concat_fwd(const dcnr_dims *dims, const dcnr_params *params, const dcnr_batch *batch,
                          float *x0, int64_t ldx0, dcnr_stream_t stream);

/* Dense embedding-table gradients from dx0 [batch, lddx]: stable sort of (id, row) then a
 * per-segment sum in batch order -- the implicit embedding_dense_backward at train.py:225.
 * Only the table pointers of `grads` are used.  scratch: dcnr_workspace_bytes(kind=2). */
int dcnr_embed_scatter_bwd(const dcnr_dims *dims, const dcnr_batch *batch, const float *dx0, int64_t lddx,
                           const dcnr_grads *grads, void *scratch, int64_t scratch_bytes,
                           dcnr_stream_t stream);

/* n_layers CrossLayers, y_{l+1} = y_l * (1 + y_l . w_l) + b_l  -- train.py:96-99, :167-168.
 * w_host / b_host: HOST arrays of n_layers device pointers ([dim] each).  x, y: [batch, ld]. */
int dcnr_cross_fwd(const float *x, int64_t ldx, int64_t batch, int32_t dim, int32_t n_layers,
                   const float *const *w_host, const float *const *b_host, float *y, int64_t ldy,
                   dcnr_stream_t stream);

/* Backward of dcnr_cross_fwd with forward recomputation from x.  gy: [batch, ldg].
 * gx: [batch, ldgx]; gw_host / gb_host: HOST arrays of device pointers ([dim] each, overwritten).
 * scratch: dcnr_cross_bwd_scratch_bytes(). */
int64_t dcnr_cross_bwd_scratch_bytes(int64_t batch, int32_t dim, int32_t n_layers);
int dcnr_cross_bwd(const float *x, int64_t ldx, int64_t batch, int32_t dim, int32_t n_layers,
                   const float *const *w_host, const float *const *b_host, const float *gy, int64_t ldg,
                   float *gx, int64_t ldgx, float *const *gw_host, float *const *gb_host, void *scratch,
                   int64_t scratch_bytes, dcnr_stream_t stream);

/* y = x W^T + bias   (nn.Linear; train.py:143,161; :105,109,114,118).  x [m,k], W [n,k], y [m,n].
 * Optional fused epilogue, applied in this order when the pointer is non-NULL:
 *   v = acc * col_scale[n] ; v += bias[n] ; v += residual[m,n] ; v = max(v, 0) if relu. */
int dcnr_linear_fwd(const float *x, int64_t ldx, const float *w, int64_t ldw, const float *bias,
                    const float *col_scale, const float *residual, int64_t ldr, int relu, float *y, int64_t ldy,
                    int64_t m, int32_t n, int32_t k, int32_t precision, void *workspace, int64_t workspace_bytes,
                    dcnr_stream_t stream);
/* Workspace of dcnr_linear_fwd / dcnr_linear_dgrad / dcnr_cross_v2_fwd for an [n, k] weight: room for its tensor-core
 * operand form (tf32 hi / lo split, transposed for the dgrad).  0 for DCNR_PREC_FP32 (workspace may then be NULL). */
int64_t dcnr_linear_workspace_bytes(int32_t n, int32_t k, int32_t precision);

/* DCN-v2 ("full-matrix") cross layer -- OPT-IN variant (BASELINE.json north_star item 2), not the reference's
 * rank-1 CrossLayer (train.py:96-99, served by dcnr_cross_fwd above):
 *     y = x0 * (x W^T + bias) + x        elementwise *, x0 / x / y [m, d], W [d, d] (nn.Linear layout), bias [d] or NULL.
 * One tcgen05 GEMM with bias, Hadamard and residual fused into the epilogue when d % 32 == 0 and precision != fp32
 * (rows padded with zeros to d = round_up(D, 32), W zero-padded to [d, d]); the CUDA-core GEMM otherwise. */
int dcnr_cross_v2_fwd(const float *x0, int64_t ldx0, const float *x, int64_t ldx, const float *w, int64_t ldw,
                      const float *bias, float *y, int64_t ldy, int64_t m, int32_t d, int32_t precision,
                      void *workspace, int64_t workspace_bytes, dcnr_stream_t stream);
/* Elementwise part of its backward for upstream g = dL/dy and u = x W^T + bias (dcnr_linear_fwd):
 *     gm = g * x0 ;  dx0 = (accumulate ? dx0 : 0) + g * u.
 * The rest is dx = gm W + g (dcnr_linear_dgrad with residual g) and dW = gm^T x, db = sum gm (dcnr_linear_wgrad). */
int dcnr_cross_v2_bwd_prep(const float *g, int64_t ldg, const float *x0, int64_t ldx0, const float *u, int64_t ldu,
                           float *gm, int64_t ldgm, float *dx0, int64_t lddx0, int accumulate, int64_t m, int32_t d,
                           dcnr_stream_t stream);

/* dx = dy W (+ residual)   -- autograd of nn.Linear wrt its input.  dy [m,n], W [n,k], dx [m,k]. */
int dcnr_linear_dgrad(const float *dy, int64_t lddy, const float *w, int64_t ldw, const float *residual,
                      int64_t ldr, float *dx, int64_t lddx, int64_t m, int32_t n, int32_t k, int32_t precision,
                      void *workspace, int64_t workspace_bytes, dcnr_stream_t stream);

/* dW = dy^T x, db = column sums of dy  -- autograd of nn.Linear wrt weight and bias.
 * dw [n, lddw] (only the first k columns are written), db [n] (may be NULL).
 * Split over the batch in fixed chunks, reduced in chunk order (deterministic).
 * scratch: dcnr_linear_wgrad_scratch_bytes(). */
int64_t dcnr_linear_wgrad_scratch_bytes(int64_t m, int32_t n, int32_t k);
int dcnr_linear_wgrad(const float *dy, int64_t lddy, const float *x, int64_t ldx, float *dw, int64_t lddw,
                      float *db, int64_t m, int32_t n, int32_t k, int32_t precision, void *scratch,
                      int64_t scratch_bytes, dcnr_stream_t stream);

/* Train-mode BatchNorm1d statistics of z [m,n]: mean[n], rstd[n] = 1/sqrt(biased_var + eps);
 * if running_mean != NULL also running_mean/var (unbiased var, momentum) and num_batches_tracked.
 * (train.py:115,119).  scratch: dcnr_bn_scratch_bytes(). */
int64_t dcnr_bn_scratch_bytes(int64_t m, int32_t n);
int dcnr_bn_stats(const float *z, int64_t ldz, int64_t m, int32_t n, float eps, float momentum,
                  float *mean, float *rstd, float *running_mean, float *running_var,
                  int64_t *num_batches_tracked, void *scratch, int64_t scratch_bytes, dcnr_stream_t stream);

/* out = relu(gamma*(z-mean)*rstd + beta + residual) * keep * post_scale   (train.py:115-121).
 * residual may be NULL; keep: injected uint8 mask [m,n] or NULL; when keep == NULL and drop_p > 0
 * the Philox mask of (seed, layer_tag, row, col) is used.  post_scale = 1/(1-drop_p). */
int dcnr_bn_act_fwd(const float *z, int64_t ldz, const float *mean, const float *rstd, const float *gamma,
                    const float *beta, const float *residual, int64_t ldr, const uint8_t *keep, float drop_p,
                    uint64_t seed, uint32_t layer_tag, float *out, int64_t ldo, int64_t m, int32_t n,
                    dcnr_stream_t stream);

/* Backward through [ReLU (+dropout scale)] and train-mode BatchNorm:
 *   dy = g * post_scale * (out > 0);  dgamma = sum dy*xhat;  dbeta = sum dy;
 *   dz = gamma*rstd*(dy - dbeta/m - xhat*dgamma/m);  dbias = column sums of dz.
 * g [m,n] upstream gradient; out = the forward output of dcnr_bn_act_fwd; z = its input.
 * dz [m,n] (may alias g); dy_out (may be NULL or alias g) receives dy (the identity-path
 * gradient of a ResBlock, train.py:120).  dbias may be NULL. */
int dcnr_bn_act_bwd(const float *g, int64_t ldg, const float *out, int64_t ldo, const float *z, int64_t ldz,
                    const float *mean, const float *rstd, const float *gamma, float post_scale, float *dz,
                    int64_t lddz, float *dy_out, int64_t lddy, float *dgamma, float *dbeta, float *dbias,
                    int64_t m, int32_t n, void *scratch, int64_t scratch_bytes, dcnr_stream_t stream);

/* Fused eval-mode deep tower -- initial_deep_layer, every ResBlock with folded BatchNorm, and the deep half of the final
 * dot in ONE persistent tcgen05 kernel; activations never leave the SM (train.py:161-170 in eval(), main.py:120-127):
 *     logits[b] = wf[0:H] . tower(x0[b]) + logit_cross[b] + bf
 * x0 [m, ldx0] is the padded output of dcnr_embed_concat_fwd (ldx0 >= in_dim_pad, pad columns zero); logit_cross [m] (the
 * cross half, may be NULL).  precision: DCNR_PREC_FP16X3 or DCNR_PREC_BF16.  Needs hidden == 256, 1..4 ResBlocks
 * (dcnr_tower_eval_supported).  workspace: dcnr_tower_eval_workspace_bytes().  flags: see dcnr_dims.eval_flags (may be NULL).
 * options: bits 8.. = cap on the number of CTAs (0 = one per SM) -- measurement
 * and test aids.  flags[1..2] receive a diagnostic record if a pipeline wait times out (the kernel then traps), so flags must
 * point at >= 3 ints. */
int dcnr_tower_eval_supported(const dcnr_dims *dims);
int64_t dcnr_tower_eval_workspace_bytes(const dcnr_dims *dims);
/* The weight pack of the fused tower (pre-split fp16 / bf16 weights, folded BatchNorm scale / shift) for dcnr_dims.tower_pack. */
int64_t dcnr_tower_pack_bytes(const dcnr_dims *dims);
int dcnr_tower_prepare(const dcnr_dims *dims, const dcnr_params *params, int32_t precision, void *pack, int64_t pack_bytes,
                       dcnr_stream_t stream);
int dcnr_tower_eval(const dcnr_dims *dims, const dcnr_params *params, const float *x0, int64_t ldx0,
                    const float *logit_cross, float *logits, int64_t m, int32_t precision, int32_t options, int32_t *flags,
                    void *workspace, int64_t workspace_bytes, dcnr_stream_t stream);

/* logit[b] = wf[0:H] . deep[b] + extra[b] + bf   -- the deep half of train.py:169-170 (the cross
 * half, wf[H:H+D] . cross[b], is produced by the fused gather+cross kernel into `extra`).
 * extra may be NULL; bf: 1 float on the device, may be NULL. */
int dcnr_rowdot_fwd(const float *a, int64_t lda, const float *w, const float *extra, const float *bf,
                    float *out, int64_t m, int32_t n, dcnr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Cosine top-k (NearestNeighbors(metric='cosine', algorithm='brute'); main.py:268-269, :200, :300).
 * Arithmetic contract (bit-exact with oracle/knn_oracle.c): sequential-fma row norms and dot
 * products, dist = clip(1 - sim, 0, 2), total order (dist ascending, index ascending).
 * ------------------------------------------------------------------------------------------ */

/* fit(): out[r] = in[r] / max-safe ||in[r]||  (zero rows stay zero).  in/out: [n, d]. */
int dcnr_knn_normalize(const float *in, float *out, int64_t n, int32_t d, dcnr_stream_t stream);

int64_t dcnr_knn_scratch_bytes(int64_t n, int32_t d, int32_t n_queries, int32_t k);
/* kneighbors(): catalog_hat [n,d] and queries_hat [n_queries,d] both pre-normalised.
 * dist_out [n_queries,k] fp32, idx_out [n_queries,k] int64 = idx_base + local row, sorted by the
 * contract order; entries beyond n are (inf, -1). */
int dcnr_knn_topk(const float *catalog_hat, int64_t n, int32_t d, const float *queries_hat, int32_t n_queries,
                  int32_t k, int64_t idx_base, float *dist_out, int64_t *idx_out, void *scratch,
                  int64_t scratch_bytes, dcnr_stream_t stream);

/* The same result for a BATCH of queries, found with a tensor-core shortlist: tcgen05 kind::tf32 scores of every (row, query)
 * pair decide which rows are re-scored with the exact arithmetic above (error bound 2e-3 on unit vectors, thresholds from exact
 * top-k of row samples: no member of the true top-k can be dropped -- csrc/topk_tc.cu has the argument).  One catalog pass for
 * up to 1 024 queries (d = 16).  Shapes: d in {16, 24, 32, 48, 64} (the reference's emb_dim search space, train.py:180), 2^18 <= n <= 2^24, k <= 256
 * (dcnr_knn_tc_supported).
 * status: optional DEVICE int, OR-ed: bit 0 = a per-query shortlist overflowed (pathological duplicates; the outputs are
 * then NOT valid and the caller re-runs the batch with dcnr_knn_topk). */
int dcnr_knn_tc_supported(int64_t n, int32_t d, int32_t n_queries, int32_t k);
int64_t dcnr_knn_tc_scratch_bytes(int64_t n, int32_t d, int32_t n_queries, int32_t k);
int dcnr_knn_topk_tc(const float *catalog_hat, int64_t n, int32_t d, const float *queries_hat, int32_t n_queries,
                     int32_t k, int64_t idx_base, float *dist_out, int64_t *idx_out, void *scratch,
                     int64_t scratch_bytes, int32_t *status, dcnr_stream_t stream);

/* Cross-shard merge: parts [n_parts, n_queries, k] (each row sorted, (inf,-1) padded) -> [n_queries, k]
 * in the same total order, so the result does not depend on the shard count. */
int dcnr_knn_merge(const float *dist_parts, const int64_t *idx_parts, int32_t n_parts, int32_t n_queries,
                   int32_t k, float *dist_out, int64_t *idx_out, dcnr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * MMR re-rank (rerank_with_mmr, main.py:133-169), one CTA per request.
 * scores / emb_idx: the candidates of all requests back to back, each request in ranked (score-descending) order;
 * offsets: [n_requests + 1] prefix offsets; emb_idx: row of item_emb per candidate, < 0 = item unknown to the id map
 * (skipped, main.py:150).  order_out: [n_requests, top_k] positions INSIDE each request's candidate list (-1 padded),
 * count_out: [n_requests].  max_candidates: the largest request (host value, sizes the shared memory).
 * Bit-exact with oracle/mmr_oracle.c (sequential-fma cosine, fp32 mmr, first maximum wins).
 * ------------------------------------------------------------------------------------------ */
int dcnr_mmr_rerank(const float *item_emb, int64_t n_items, int32_t d, const float *scores, const int64_t *emb_idx,
                    const int32_t *offsets, int32_t n_requests, float lambda, int32_t top_k, int32_t max_candidates,
                    int32_t *order_out, int32_t *count_out, dcnr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Data parallelism (no counterpart in the reference, which is single-process).
 * One process per GPU; the communicator is NCCL over NVLink / NVSwitch, reached through dlopen.
 * ------------------------------------------------------------------------------------------ */

/* Rank 0 creates a 128-byte id and distributes it (e.g. torch.distributed broadcast); every rank then calls
 * dcnr_comm_create with the same id.  The handle goes into dcnr_dims.comm and the calls below. */
int dcnr_comm_unique_id(uint8_t *id_host /* [128] */);
int dcnr_comm_create(const uint8_t *id_host, int32_t rank, int32_t world, void **comm_out);
int dcnr_comm_destroy(void *comm);
int dcnr_comm_info(const void *comm, int32_t *rank, int32_t *world);
/* 1 when the ranks of `comm` mapped each other's memory (CUDA IPC over NVLink) at creation: the small BatchNorm exchanges of a
 * training step then run as one peer-to-peer kernel each instead of an NCCL all-gather. */
int dcnr_comm_uses_peer_memory(const void *comm);
/* enable = 0 parks the peer-memory path (the exchanges go through NCCL), 1 restores it when the mapping exists.  Must be called
 * with the same value on every rank, between steps.  Both paths fold the ranks in rank order in float64: identical results. */
int dcnr_comm_set_peer_memory(void *comm, int32_t enable);
/* In-place sum over the ranks (gradient all-reduce after dcnr_backward). */
int dcnr_comm_allreduce_f32(void *comm, float *buf, int64_t count, dcnr_stream_t stream);
/* recv[r*bytes_per_rank ..] = rank r's send buffer (sparse embedding-gradient segments, top-k lists). */
int dcnr_comm_allgather(void *comm, const void *send, void *recv, int64_t bytes_per_rank, dcnr_stream_t stream);
/* Variable all-to-all (row-sharded embedding tables: ids to the owning rank, rows back; gradient rows to the
 * owner in backward).  The four arrays are HOST arrays with one entry per rank (bytes / byte offsets). */
int dcnr_comm_alltoallv(void *comm, const void *send, const int64_t *send_bytes_host, const int64_t *send_off_host,
                        void *recv, const int64_t *recv_bytes_host, const int64_t *recv_off_host, dcnr_stream_t stream);

/* out[i] = table[ids[i]] for a [rows, width] fp32 table (the owner-side lookup of a row-sharded embedding
 * exchange; ids are LOCAL row numbers).  Bit-exact copy. */
int dcnr_gather_rows(const float *table, int64_t rows, int32_t width, const int64_t *ids, int64_t n, float *out,
                     dcnr_stream_t stream);
/* Dense gradient of ONE table from per-sample gradient rows g [n, width] and LOCAL ids [n]: the sorted-segment
 * scatter-add of dcnr_embed_scatter_bwd for a single table (owner side of the sharded exchange).
 * scratch: dcnr_workspace_bytes(kind = 2) for a batch of n rows. */
int dcnr_scatter_rows(const int64_t *ids, int64_t n, int64_t rows, int32_t width, const float *g, int64_t ldg,
                      float *grad_table, void *scratch, int64_t scratch_bytes, dcnr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DCNR_H_ */
