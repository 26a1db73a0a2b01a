// Internal kernel-launcher interfaces shared by the translation units of libdcnr_sm100a.so.
#pragma once

#include <algorithm>
#include <cstring>

#include "common.cuh"

namespace dcnr {

// ---- gather + cross (embed_cross.cu) ---------------------------------------------------------
struct GatherSeg {
    const float *table;
    const int64_t *ids;      // id of row b is ids[b * id_stride]
    int64_t rows;
    int32_t id_stride, width, col0, pad_;
};
struct GatherArgs {
    GatherSeg seg[2 + DCNR_MAX_CAT];
    const float *num;        // [B, n_num]
    int32_t n_seg, n_num, num_col0, D;
};
struct CrossArgs {
    const float *w[DCNR_MAX_CROSS];
    const float *b[DCNR_MAX_CROSS];
    int32_t L, D;
};

int make_gather_args(const dcnr_dims *dims, const dcnr_params *params, const dcnr_batch *batch, GatherArgs *out);

// x0 = gather(ga) (or x_in), cross layers applied in registers.  Any of x0_out / y_out /
// logit_part may be NULL.  logit_part[b] = wf_cross . cross_out[b].
int launch_embed_cross_fwd(const GatherArgs *ga, const float *x_in, int64_t ldx_in, int64_t B,
                           const CrossArgs &ca, int32_t dim_pad, float *x0_out, int64_t ldx0, float *y_out,
                           int64_t ldy, const float *wf_cross, float *logit_part, int32_t *err_flag,
                           cudaStream_t stream);

int64_t cross_bwd_partial_floats(int64_t B, int32_t dim_pad, int32_t L);
// upstream gradient either gy [B, ldg] or dlogit[b] * wf_cross (then gwf = d wf_cross is produced)
int launch_cross_bwd(const float *x, int64_t ldx, int64_t B, const CrossArgs &ca, int32_t dim_pad, const float *gy,
                     int64_t ldg, const float *dlogit, const float *wf_cross, float *gx, int64_t ldgx,
                     int accumulate, float *const *gw, float *const *gb, float *gwf, float *partials,
                     cudaStream_t stream);

// ---- dense layers (gemm_simt.cu / gemm_tc.cu) -------------------------------------------------
struct GemmEpilogue {
    const float *col_scale;   // [n] or NULL
    const float *bias;        // [n] or NULL
    const float *residual;    // [m, ldr] or NULL
    int64_t ldr;
    int relu;
    const float *hadamard = nullptr;   // [m, ldh] or NULL: v = (acc*col_scale + bias) * hadamard + residual (DCN-v2 cross layer)
    int64_t ldh = 0;
};
// C[m,n] = sum_k A(m,k) * B(k,n) with
//   a_kmajor: A(m,k) = A[m*lda + k]   else A(m,k) = A[k*lda + m]
//   b_kmajor: B(k,n) = B[n*ldb + k]   else B(k,n) = B[k*ldb + n]
// split_k > 1: blockIdx.z handles a contiguous k-range and writes its own [m,n] slab at
// C + z * m * ldc (no epilogue); the caller reduces the slabs in order.
int launch_gemm_simt(const float *A, int64_t lda, bool a_kmajor, const float *B, int64_t ldb, bool b_kmajor,
                     float *C, int64_t ldc, int64_t m, int64_t n, int64_t k, int split_k,
                     const GemmEpilogue &epi, cudaStream_t stream);

bool gemm_tc_supported(int precision, bool a_kmajor, bool b_kmajor, int64_t lda, int64_t ldb, int64_t ldc, int64_t m,
                       int64_t n, int64_t k, int split_k);
// TF32X3: B is the tf32-rounded (hi) half of the weight and B_lo its exact remainder (launch_split_tf32)
int launch_gemm_tc(int precision, const float *A, int64_t lda, bool a_kmajor, const float *B, int64_t ldb, bool b_kmajor,
                   float *C, int64_t ldc, int64_t m, int64_t n, int64_t k, int split_k, const GemmEpilogue &epi,
                   cudaStream_t stream, const float *B_lo, const float *dot_w = nullptr, float *dot_out = nullptr);
int gemm_tc_n_tiles(int64_t n, int precision, int64_t k);   // partial row dots per row a FusedDot gets for an [n, k] layer
// hi = rn_tf32(src), lo = src - hi, dense [rows, cols] (or transposed: [cols, rows]); lo may be NULL when transposing
int launch_split_tf32(const float *src, int64_t lds, float *hi, float *lo, int32_t rows, int32_t cols, bool transpose,
                      cudaStream_t stream);
// A pre-processed weight operand for the tensor-core path: `w` [n, k] row-major dense (ld = k).
struct WeightOp {
    const float *hi;   // tf32-rounded weight (TF32X3) or the (transposed) weight itself (TF32)
    const float *lo;   // exact remainder (TF32X3), else NULL
    int64_t ld;
};
// Optional fusion of the final row dot into the GEMM epilogue (tensor-core path only):
// dot_out[tile][m] = sum over the tile's columns of epilogue(C)[m,n] * dot_w[n]; C itself is not stored.
struct FusedDot {
    const float *w;
    float *out;       // [gemm_tc_n_tiles(n, precision, k)][m]
};
// precision dispatch (linear.cu).  wop == NULL -> B is used as is (CUDA-core path unless precision == TF32).
int gemm_any(int precision, const float *A, int64_t lda, bool a_kmajor, const float *B, int64_t ldb, bool b_kmajor,
             float *C, int64_t ldc, int64_t m, int64_t n, int64_t k, int split_k, const GemmEpilogue &epi,
             cudaStream_t stream, const WeightOp *wop = nullptr, const FusedDot *dot = nullptr);
// true when gemm_any would run this forward GEMM on the tensor-core kernel (so a FusedDot may be passed)
bool gemm_any_uses_tc(int precision, int64_t lda, int64_t ldc, int64_t m, int64_t n, int64_t k, const WeightOp *wop,
                      int64_t ldb);
// tcgen05 weight gradient (gemm_wgrad_tc.cu): dW[n, k] = dy^T x with both operands MN-major, per-slab partials
bool wgrad_tc_supported(int precision, int64_t lddy, int64_t ldx, int64_t m, int32_t n, int32_t k);
int wgrad_tc_slabs(int64_t m, int32_t n);
int launch_wgrad_tc(int precision, const float *dy, int64_t lddy, const float *x, int64_t ldx, float *slabs, int64_t m,
                    int32_t n, int32_t k, cudaStream_t stream, const float *center = nullptr);
int64_t wgrad_scratch_floats(int64_t m, int32_t n, int32_t k);
// dW[n, 0:k_valid] (ld lddw) = dy^T x over k (padded) columns; db = column sums of dy (may be NULL)
// center [k] + dy_colsum [n] (both or neither): the tensor-core path computes dy^T (x - center) + dy_colsum (x) center, which is
// the same dW with far less cancellation when the columns of x have means that are large against their spread and dy sums to
// ~0 over the batch (the layer after a BatchNorm backward); other paths ignore them.
int launch_linear_wgrad(int precision, const float *dy, int64_t lddy, const float *x, int64_t ldx, float *dw,
                        int64_t lddw, float *db, int64_t m, int32_t n, int32_t k, int32_t k_valid, float *scratch,
                        cudaStream_t stream, const float *center = nullptr, const float *dy_colsum = nullptr);
// out[c] = mean of a[0:min(m, 512), c]: a cheap estimate of the column means (one launch), n % 32 == 0
int launch_col_mean_sample(const float *a, int64_t lda, int64_t m, int32_t n, float *out, cudaStream_t stream);

// ---- tensor-core shortlist + exact re-score top-k for query batches (topk_tc.cu) ----
bool knn_tc_supported(int64_t n, int32_t d, int32_t n_queries, int32_t k);
int64_t knn_tc_scratch_bytes(int64_t n, int32_t d, int32_t n_queries, int32_t k);
int launch_knn_tc(const float *cat, int64_t n, int32_t d, const float *queries, int32_t n_queries, int32_t k, int64_t idx_base,
                  float *dist_out, int64_t *idx_out, void *scratch, int64_t scratch_bytes, int32_t *status, cudaStream_t st);

// ---- fused eval tower (tower_eval.cu): initial layer + ResBlocks + deep half of the final dot in ONE persistent kernel ----
bool tower_eval_supported(const dcnr_dims *d);                 // hidden 256, 1..4 ResBlocks, in_dim_pad <= 256
int64_t tower_pack_bytes(const dcnr_dims *d);
int launch_tower_prepare(const dcnr_dims *d, const dcnr_params *p, void *pack, int precision, cudaStream_t stream);
int launch_tower_eval(const dcnr_dims *d, const float *x0, int64_t ldx0, const float *logit_cross, const float *bf,
                      const void *pack, float *out, int64_t M, int32_t *flags, int precision, int options,
                      cudaStream_t stream);
// dense-layer precision the GEMM kernels run for a requested mode: the fp16x3 / bf16 modes exist only in the fused eval
// tower; everywhere else (training, unsupported shapes, operator-level calls) they run as tf32x3 / tf32
static inline int gemm_precision(int precision) {
    return precision == DCNR_PREC_FP16X3 ? DCNR_PREC_TF32X3 : (precision == DCNR_PREC_BF16 ? DCNR_PREC_TF32 : precision);
}

// ---- data-parallel communicator (comm.cu); `comm` is the opaque handle of dcnr_comm_create or NULL ----
int comm_world(const void *comm);
int comm_rank(const void *comm);
int comm_allgather(const void *comm, const void *send, void *recv, int64_t bytes_per_rank, cudaStream_t stream);
// recv[r * n + i] = rank r's send[i]: the small SyncBN exchanges (one peer-memory kernel over NVLink when available, else NCCL)
int comm_allgather_f64(const void *comm, const double *send, double *recv, int n, cudaStream_t stream);
constexpr int kMaxWorld = 64;

// ---- batch norm / elementwise (bn.cu) -------------------------------------------------------
// `comm` != NULL: statistics (forward) and the two backward column sums are taken over the batches of ALL ranks
// (SyncBN), so an N-rank step equals the single-device step on the concatenated batch.
int64_t bn_scratch_floats(int64_t m, int32_t n);
int launch_bn_stats(const float *z, int64_t ldz, int64_t m, int32_t n, float eps, float momentum, float *mean,
                    float *rstd, float *running_mean, float *running_var, int64_t *nbt, float *scratch,
                    cudaStream_t stream, const void *comm = nullptr);
int launch_bn_act_fwd(const float *z, int64_t ldz, const float *mean, const float *rstd, const float *gamma,
                      const float *beta, const float *residual, int64_t ldr, const uint8_t *keep, float drop_p,
                      uint64_t seed, uint32_t layer_tag, float *out, int64_t ldo, int64_t m, int32_t n,
                      cudaStream_t stream, const uint64_t *seed_step = nullptr);
int launch_inc_u64(uint64_t *p, cudaStream_t stream);
int launch_bn_act_bwd(const float *g, int64_t ldg, const float *out, int64_t ldo, const float *z, int64_t ldz,
                      const float *mean, const float *rstd, const float *gamma, float post_scale, float *dz,
                      int64_t lddz, float *dy_out, int64_t lddy, float *dgamma, float *dbeta, float *dbias,
                      int64_t m, int32_t n, float *scratch, cudaStream_t stream, const void *comm = nullptr);
// eval: scale[c] = gamma*rsqrt(rv+eps), shift[c] = beta + (b_lin - rm)*scale   (folded in double)
int launch_bn_fold(const float *gamma, const float *beta, const float *rm, const float *rv, const float *lin_bias,
                   float eps, float *scale, float *shift, int32_t n, cudaStream_t stream);
// dst[r, 0:cols_pad] = src[r, 0:cols] zero padded
int launch_pad_rows(const float *src, int64_t lds, float *dst, int64_t ldd, int32_t rows, int32_t cols,
                    int32_t cols_pad, cudaStream_t stream);
// column sums of a [m,n] matrix (deterministic); scratch: bn_scratch_floats
int launch_colsum(const float *a, int64_t lda, int64_t m, int32_t n, float *out, float *scratch, cudaStream_t stream);

// ---- head (head.cu) ---------------------------------------------------------------------------
int launch_rowdot_fwd(const float *a, int64_t lda, const float *w, const float *extra, const float *bf,
                      float *out, int64_t m, int32_t n, cudaStream_t stream);
int launch_combine_logits(const float *parts, int n_parts, int64_t m, const float *extra, const float *bf, float *out,
                          cudaStream_t stream);
// dh[b,:] = dlogit[b]*w ; dw[c] = sum_b dlogit[b]*a[b,c] ; dbf = sum_b dlogit[b]
int launch_rowdot_bwd(const float *dlogit, const float *a, int64_t lda, const float *w, float *dh, int64_t lddh,
                      float *dw, float *dbf, int64_t m, int32_t n, float *scratch, cudaStream_t stream);

// out[j] = base[j] + sum_n v[n] * W[n*ldw + j]  (base may be NULL)
int launch_vecmat_add(const float *v, const float *W, int64_t ldw, int32_t n_rows, int32_t n_cols, const float *base,
                      float *out, cudaStream_t stream);

// ---- stable radix sort of (key, value) pairs (radix_sort.cu) ----------------------------------------------------
int64_t radix_sort_scratch_bytes(int64_t n);
int launch_radix_sort_pairs(uint32_t *k0, uint32_t *v0, uint32_t *k1, uint32_t *v1, int64_t n, int end_bit, void *scratch,
                            uint32_t **k_out, uint32_t **v_out, cudaStream_t stream);

// ---- embedding backward (embed_bwd.cu) ----------------------------------------------------------
int64_t scatter_scratch_bytes(int64_t B);
int launch_embed_scatter_pair(const int64_t *ids0, int64_t stride0, int64_t rows0, float *grad0, int32_t col0,
                              const int64_t *ids1, int64_t stride1, int64_t rows1, float *grad1, int32_t col1, int64_t B,
                              int32_t width, const float *dx0, int64_t lddx, void *scratch, int64_t scratch_bytes,
                              cudaStream_t stream);
// ids_out[b] = (user_ids[b], item_ids[b]); rows_out[b] = dx0[b, 0 : 2*emb_dim] for b < B, (0, 0) / zero rows up to cap
// (the payload of the sparse gradient all-gather; cap = the largest local batch of any rank)
int launch_pack_embed_grads(const int64_t *user_ids, const int64_t *item_ids, const float *dx0, int64_t lddx, int64_t B,
                            int64_t cap, int32_t emb_dim, int64_t *ids_out, float *rows_out, cudaStream_t stream);
int launch_embed_scatter(const int64_t *ids, int64_t id_stride, int64_t B, int64_t n_rows, int32_t width,
                         const float *dx0, int64_t lddx, int32_t col0, float *grad_table, void *scratch,
                         int64_t scratch_bytes, cudaStream_t stream);

}  // namespace dcnr
