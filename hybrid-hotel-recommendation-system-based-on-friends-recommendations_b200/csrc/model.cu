// Whole-model orchestration: DCN_RecSys.forward in eval and train mode and its backward,
// chained from the operator launchers on one stream with no host synchronisation.
//
// Reference lines: DCN_RecSys.forward train.py:155-170 (main.py:114-127); training step
// train.py:223-225; ranking call main.py:320-322.
#include "kernels.cuh"

namespace dcnr {

constexpr int64_t kEvalChunkRows = 1 << 20;   // rows per pass of the eval forward (bounds the workspace)
constexpr int64_t kEvalChunkRowsFused = 1 << 22;   // fused tower: only x0 (256 B / row) is staged, so a pass can be 4x longer
                                                   // (fewer launch ramps and tile-wave tails: 55 tiles per SM per 2^20 rows)

static int check_dims(const dcnr_dims *d) {
    DCNR_REQUIRE(d != nullptr, "null dims");
    DCNR_REQUIRE(d->hidden >= 4 && d->hidden % 4 == 0, "hidden_dim %d must be a multiple of 4", d->hidden);
    DCNR_REQUIRE(d->n_res >= 0 && d->n_res <= DCNR_MAX_RES, "n_res_blocks %d > %d", d->n_res, DCNR_MAX_RES);
    DCNR_REQUIRE(d->n_cross >= 0 && d->n_cross <= DCNR_MAX_CROSS, "n_cross_layers %d > %d", d->n_cross, DCNR_MAX_CROSS);
    DCNR_REQUIRE(d->n_cat >= 0 && d->n_cat <= DCNR_MAX_CAT, "n_cat %d > %d", d->n_cat, DCNR_MAX_CAT);
    DCNR_REQUIRE(d->in_dim >= 1 && d->in_dim <= 256, "input_dim %d unsupported (max 256)", d->in_dim);
    DCNR_REQUIRE(d->in_dim_pad == (int32_t)round_up(d->in_dim, DCNR_PAD), "in_dim_pad must be round_up(in_dim, 32)");
    DCNR_REQUIRE(d->dropout_p >= 0.f && d->dropout_p < 1.f, "dropout must be in [0,1)");
    return DCNR_OK;
}

// Tensor-core operand forms of the dense-layer weights, carved from a workspace.
//   forward : [n, k] hi/lo split (TF32X3 only; TF32 reads the raw weight)
//   dgrad   : W^T [k, n] (hi, and lo for TF32X3) because the tcgen05 kernel takes K-major operands
struct WeightOps {
    WeightOp w0, w1[DCNR_MAX_RES], w2[DCNR_MAX_RES];
    bool on = false;
    static int64_t floats(const dcnr_dims *d) {
        return 2 * ((int64_t)d->hidden * d->in_dim_pad + 2 * (int64_t)d->n_res * d->hidden * d->hidden) + 64;
    }
    const WeightOp *get(const WeightOp &w) const { return on ? &w : nullptr; }
    int prepare(const dcnr_dims *d, const dcnr_params *p, const float *w0p, float *buf, bool transpose, cudaStream_t st) {
        const int prec = gemm_precision(d->precision);
        on = prec == DCNR_PREC_TF32X3 || (transpose && prec == DCNR_PREC_TF32);
        if (!on) return DCNR_OK;
        const bool lo = prec == DCNR_PREC_TF32X3;
        const int H = d->hidden, Dp = d->in_dim_pad;
        auto one = [&](WeightOp &op, const float *w, int64_t ldw, int rows, int cols) -> int {
            const int64_t n = (int64_t)rows * cols;
            op.hi = buf;
            op.lo = lo ? buf + n : nullptr;
            op.ld = transpose ? rows : cols;
            DCNR_TRY(launch_split_tf32(w, ldw, buf, lo ? buf + n : nullptr, rows, cols, transpose, st));
            buf += 2 * n;
            return DCNR_OK;
        };
        DCNR_TRY(one(w0, w0p, Dp, H, Dp));
        for (int r = 0; r < d->n_res; ++r) {
            DCNR_TRY(one(w1[r], p->res_w1[r], H, H, H));
            DCNR_TRY(one(w2[r], p->res_w2[r], H, H, H));
        }
        return DCNR_OK;
    }
};

struct TrainSaved {
    float *x0p, *w0p, *logit_cross;
    float *h[DCNR_MAX_RES + 1];
    float *z1[DCNR_MAX_RES], *d1[DCNR_MAX_RES], *z2[DCNR_MAX_RES];
    float *stats[DCNR_MAX_RES];   // mean1, rstd1, mean2, rstd2 : 4*H each block
    float *bn_scratch, *wsplit;
    void layout(const dcnr_dims *d, int64_t B, Arena &a) {
        const int64_t H = d->hidden, Dp = d->in_dim_pad;
        x0p = a.take<float>(B * Dp);
        w0p = a.take<float>(H * Dp);
        logit_cross = a.take<float>(B);
        for (int r = 0; r <= d->n_res; ++r) h[r] = a.take<float>(B * H);
        for (int r = 0; r < d->n_res; ++r) {
            z1[r] = a.take<float>(B * H);
            d1[r] = a.take<float>(B * H);
            z2[r] = a.take<float>(B * H);
            stats[r] = a.take<float>(4 * H);
        }
        bn_scratch = a.take<float>(bn_scratch_floats(B, (int32_t)H));
        wsplit = a.take<float>(WeightOps::floats(d));
    }
};

// true when the user / item table gradients are built from the GLOBAL batch: every rank all-gathers the (ids, gradient
// rows) of all ranks and runs the same sorted-segment scatter-add, so the dense gradients come out identical on all
// ranks, in the single-device summation order, for ~150 B per sample instead of a 70 MB dense all-reduce per step
static bool dp_sparse_tables(const dcnr_dims *d) { return comm_world(d->comm) > 1 && d->dp_sparse_tables != 0; }

struct BwdScratch {
    float *ga, *gb, *gc, *dx0, *cross_partials, *wgrad, *bn, *wsplit, *vec_a, *vec_b, *vec_c;
    void *scatter;
    int64_t scatter_bytes;
    int64_t *pack_ids, *all_ids;       // [B][2] (user, item) of this rank / [world*B][2] of all ranks
    float *pack_rows, *all_rows;       // [B][2E] / [world*B][2E]
    void layout(const dcnr_dims *d, int64_t B, Arena &a, bool with_scatter) {
        const int64_t H = d->hidden, Dp = d->in_dim_pad;
        ga = a.take<float>(B * H);
        gb = a.take<float>(B * H);
        gc = a.take<float>(B * H);
        dx0 = a.take<float>(B * Dp);
        cross_partials = a.take<float>(cross_bwd_partial_floats(B, (int32_t)Dp, d->n_cross));
        wgrad = a.take<float>(std::max(wgrad_scratch_floats(B, (int32_t)H, (int32_t)H),
                                       wgrad_scratch_floats(B, (int32_t)H, (int32_t)Dp)));
        bn = a.take<float>(bn_scratch_floats(B, (int32_t)H));
        wsplit = a.take<float>(WeightOps::floats(d));
        vec_a = a.take<float>(H);
        vec_b = a.take<float>(H);
        vec_c = a.take<float>(std::max<int64_t>(H, Dp));
        const int world = dp_sparse_tables(d) ? comm_world(d->comm) : 1;
        const int64_t cap = (world > 1 && d->dp_batch_cap > B) ? d->dp_batch_cap : B;    // rows per rank in the sparse exchange
        scatter_bytes = with_scatter ? scatter_scratch_bytes(2 * cap * world) : 0;      // user + item share one sort
        scatter = a.take<char>(scatter_bytes);
        pack_ids = all_ids = nullptr;
        pack_rows = all_rows = nullptr;
        if (with_scatter && world > 1) {
            pack_ids = a.take<int64_t>(cap * 2);
            all_ids = a.take<int64_t>(cap * 2 * world);
            pack_rows = a.take<float>(cap * 2 * d->emb_dim);
            all_rows = a.take<float>(cap * 2 * d->emb_dim * world);
        }
    }
};

struct EvalWs {
    float *x0p, *w0p, *logit_cross, *ha, *hb, *ht, *fold, *wsplit, *dot_parts;   // fold: per block scale1, shift1, scale2, shift2
    char *tower_pack;
    // the fused tower (fp16x3 / bf16 on a supported shape) needs x0, the cross half of the logit and its weight pack only
    static bool fused(const dcnr_dims *d) {
        return (d->precision == DCNR_PREC_FP16X3 || d->precision == DCNR_PREC_BF16) && tower_eval_supported(d);
    }
    void layout(const dcnr_dims *d, int64_t rows, Arena &a) {
        const int64_t H = d->hidden, Dp = d->in_dim_pad;
        x0p = a.take<float>(rows * Dp);
        w0p = a.take<float>(H * Dp);
        logit_cross = a.take<float>(rows);
        tower_pack = nullptr;
        if (fused(d)) {
            tower_pack = a.take<char>(tower_pack_bytes(d));
            ha = hb = ht = fold = wsplit = dot_parts = nullptr;
            return;
        }
        ha = a.take<float>(rows * H);
        hb = a.take<float>(rows * H);
        ht = a.take<float>(rows * H);
        fold = a.take<float>((int64_t)std::max(d->n_res, 1) * 4 * H);
        wsplit = a.take<float>(WeightOps::floats(d));
        dot_parts = a.take<float>(rows * 4);
    }
};


static CrossArgs cross_args(const dcnr_dims *d, const dcnr_params *p) {
    CrossArgs ca;
    memset(&ca, 0, sizeof(ca));
    ca.L = d->n_cross;
    ca.D = d->in_dim;
    for (int l = 0; l < d->n_cross; ++l) { ca.w[l] = p->cross_w[l]; ca.b[l] = p->cross_b[l]; }
    return ca;
}

static dcnr_batch slice_batch(const dcnr_dims *d, const dcnr_batch *b, int64_t r0, int64_t rows) {
    dcnr_batch s = *b;
    s.user_ids = b->user_ids + r0;
    s.item_ids = b->item_ids + r0;
    s.cat_features = b->cat_features ? b->cat_features + r0 * d->n_cat : nullptr;
    s.num_features = b->num_features ? b->num_features + r0 * d->n_num : nullptr;
    s.batch = rows;
    return s;
}

}  // namespace dcnr

using namespace dcnr;

extern "C" int64_t dcnr_workspace_bytes(const dcnr_dims *dims, int64_t batch, int kind) {
    if (check_dims(dims) != DCNR_OK || batch < 0) return -1;
    Arena a(nullptr, 0);
    if (kind == 0) {
        EvalWs w;
        w.layout(dims, std::min<int64_t>(std::max<int64_t>(batch, 1), EvalWs::fused(dims) ? kEvalChunkRowsFused : kEvalChunkRows), a);
    } else if (kind == 1) {
        TrainSaved s;
        s.layout(dims, std::max<int64_t>(batch, 1), a);
    } else if (kind == 2) {
        BwdScratch s;
        s.layout(dims, std::max<int64_t>(batch, 1), a, true);
    } else {
        set_error("workspace kind %d unknown", kind);
        return -1;
    }
    return a.used + 256;
}

extern "C" int dcnr_forward_eval(const dcnr_dims *dims, const dcnr_params *params, const dcnr_batch *batch,
                                 float *logits, void *workspace, int64_t workspace_bytes, dcnr_stream_t stream) {
    DCNR_TRY(check_dims(dims));
    DCNR_REQUIRE(params && batch && logits && workspace, "null argument");
    const int64_t B = batch->batch;
    if (B <= 0) return DCNR_OK;
    const int64_t chunk = std::min<int64_t>(B, EvalWs::fused(dims) ? kEvalChunkRowsFused : kEvalChunkRows);
    Arena a(workspace, workspace_bytes);
    EvalWs w;
    w.layout(dims, chunk, a);
    if (!a.ok()) {
        set_error("eval workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)a.used);
        return DCNR_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    const int H = dims->hidden, D = dims->in_dim, Dp = dims->in_dim_pad, prec = gemm_precision(dims->precision);
    const CrossArgs ca = cross_args(dims, params);
    if (EvalWs::fused(dims)) {
        // gather + concat + cross network (K1), then ONE persistent kernel for the whole deep tower and the final dot
        const char *pack = static_cast<const char *>(dims->tower_pack);
        if (pack == nullptr) {
            DCNR_TRY(launch_tower_prepare(dims, params, w.tower_pack, dims->precision, st));
            pack = w.tower_pack;
        }
        for (int64_t r0 = 0; r0 < B; r0 += chunk) {
            const int64_t rows = std::min(chunk, B - r0);
            const dcnr_batch sb = slice_batch(dims, batch, r0, rows);
            GatherArgs ga;
            DCNR_TRY(make_gather_args(dims, params, &sb, &ga));
            DCNR_TRY(launch_embed_cross_fwd(&ga, nullptr, 0, rows, ca, Dp, w.x0p, Dp, nullptr, 0, params->wf + H,
                                            w.logit_cross, dims->eval_flags, st));
            DCNR_TRY(launch_tower_eval(dims, w.x0p, Dp, w.logit_cross, params->bf, pack, logits + r0, rows,
                                       dims->eval_flags, dims->precision, 0, st));
        }
        return DCNR_OK;
    }
    DCNR_TRY(launch_pad_rows(params->w0, D, w.w0p, Dp, H, D, Dp, st));
    for (int r = 0; r < dims->n_res; ++r) {
        float *f = w.fold + (int64_t)r * 4 * H;
        DCNR_TRY(launch_bn_fold(params->res_g1[r], params->res_be1[r], params->res_rm1[r], params->res_rv1[r],
                                params->res_b1[r], dims->bn_eps, f, f + H, H, st));
        DCNR_TRY(launch_bn_fold(params->res_g2[r], params->res_be2[r], params->res_rm2[r], params->res_rv2[r],
                                params->res_b2[r], dims->bn_eps, f + 2 * H, f + 3 * H, H, st));
    }
    WeightOps wo;
    DCNR_TRY(wo.prepare(dims, params, w.w0p, w.wsplit, false, st));
    for (int64_t r0 = 0; r0 < B; r0 += chunk) {
        const int64_t rows = std::min(chunk, B - r0);
        const dcnr_batch sb = slice_batch(dims, batch, r0, rows);
        GatherArgs ga;
        DCNR_TRY(make_gather_args(dims, params, &sb, &ga));
        DCNR_TRY(launch_embed_cross_fwd(&ga, nullptr, 0, rows, ca, Dp, w.x0p, Dp, nullptr, 0, params->wf + H,
                                        w.logit_cross, dims->eval_flags, st));
        // The last GEMM of the tower never writes its output: the deep half of the final dot
        // (train.py:169-170) is taken in its epilogue when that GEMM runs on the tensor-core kernel.
        const int last = 2 * dims->n_res;                  // GEMM index: 0 = initial layer, 2r+1 / 2r+2 = block r
        auto run = [&](int idx, const float *A, int64_t lda, const float *W, int64_t ldw, const WeightOp &wop,
                       const GemmEpilogue &epi, float *out, int32_t K, bool *fused) -> int {
            *fused = false;
            if (idx == last && gemm_tc_n_tiles(H, prec, K) <= 4 && gemm_any_uses_tc(prec, lda, H, rows, H, K, wo.get(wop), ldw)) {
                FusedDot fd{params->wf, w.dot_parts};
                *fused = true;
                return gemm_any(prec, A, lda, true, W, ldw, true, nullptr, H, rows, H, K, 1, epi, st, wo.get(wop), &fd);
            }
            return gemm_any(prec, A, lda, true, W, ldw, true, out, H, rows, H, K, 1, epi, st, wo.get(wop));
        };
        bool fused = false;
        GemmEpilogue e0{nullptr, params->b0, nullptr, 0, 0};
        DCNR_TRY(run(0, w.x0p, Dp, w.w0p, Dp, wo.w0, e0, w.ha, Dp, &fused));
        float *h = w.ha, *hn = w.hb;
        for (int r = 0; r < dims->n_res; ++r) {
            const float *f = w.fold + (int64_t)r * 4 * H;
            GemmEpilogue e1{f, f + H, nullptr, 0, 1};
            DCNR_TRY(run(2 * r + 1, h, H, params->res_w1[r], H, wo.w1[r], e1, w.ht, H, &fused));
            GemmEpilogue e2{f + 2 * H, f + 3 * H, h, H, 1};
            DCNR_TRY(run(2 * r + 2, w.ht, H, params->res_w2[r], H, wo.w2[r], e2, hn, H, &fused));
            std::swap(h, hn);
        }
        if (fused)
            DCNR_TRY(launch_combine_logits(w.dot_parts, gemm_tc_n_tiles(H, prec, H), rows, w.logit_cross, params->bf, logits + r0, st));
        else
            DCNR_TRY(launch_rowdot_fwd(h, H, params->wf, w.logit_cross, params->bf, logits + r0, rows, H, st));
    }
    return DCNR_OK;
}

extern "C" int dcnr_forward_train(const dcnr_dims *dims, const dcnr_params *params, const dcnr_batch *batch,
                                  uint64_t dropout_seed, const uint8_t *drop_keep_mask, float *logits, void *saved,
                                  int64_t saved_bytes, dcnr_stream_t stream) {
    DCNR_TRY(check_dims(dims));
    DCNR_REQUIRE(params && batch && logits && saved, "null argument");
    const int64_t B = batch->batch;
    // (with a data-parallel group the statistics cover the GLOBAL batch, so a single local row is fine)
    DCNR_REQUIRE(B >= 2 || dims->n_res == 0 || (B == 1 && comm_world(dims->comm) > 1),
                 "Expected more than 1 value per channel when training");
    Arena a(saved, saved_bytes);
    TrainSaved s;
    s.layout(dims, B, a);
    if (!a.ok()) {
        set_error("train workspace too small (%lld < %lld)", (long long)saved_bytes, (long long)a.used);
        return DCNR_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    const int H = dims->hidden, D = dims->in_dim, Dp = dims->in_dim_pad, prec = gemm_precision(dims->precision);
    DCNR_TRY(launch_pad_rows(params->w0, D, s.w0p, Dp, H, D, Dp, st));
    WeightOps wo;
    DCNR_TRY(wo.prepare(dims, params, s.w0p, s.wsplit, false, st));
    GatherArgs ga;
    DCNR_TRY(make_gather_args(dims, params, batch, &ga));
    const CrossArgs ca = cross_args(dims, params);
    DCNR_TRY(launch_embed_cross_fwd(&ga, nullptr, 0, B, ca, Dp, s.x0p, Dp, nullptr, 0, params->wf + H, s.logit_cross,
                                    dims->eval_flags, st));
    GemmEpilogue e0{nullptr, params->b0, nullptr, 0, 0};
    DCNR_TRY(gemm_any(prec, s.x0p, Dp, true, s.w0p, Dp, true, s.h[0], H, B, H, Dp, 1, e0, st, wo.get(wo.w0)));
    for (int r = 0; r < dims->n_res; ++r) {
        float *mean1 = s.stats[r], *rstd1 = mean1 + H, *mean2 = mean1 + 2 * H, *rstd2 = mean1 + 3 * H;
        GemmEpilogue e1{nullptr, params->res_b1[r], nullptr, 0, 0};
        DCNR_TRY(gemm_any(prec, s.h[r], H, true, params->res_w1[r], H, true, s.z1[r], H, B, H, H, 1, e1, st, wo.get(wo.w1[r])));
        DCNR_TRY(launch_bn_stats(s.z1[r], H, B, H, dims->bn_eps, dims->bn_momentum, mean1, rstd1, params->res_rm1[r],
                                 params->res_rv1[r], params->res_nbt1[r], s.bn_scratch, st, dims->comm));
        const uint8_t *keep = drop_keep_mask ? drop_keep_mask + (int64_t)r * B * H : nullptr;
        DCNR_TRY(launch_bn_act_fwd(s.z1[r], H, mean1, rstd1, params->res_g1[r], params->res_be1[r], nullptr, 0, keep,
                                   dims->dropout_p, dropout_seed, (uint32_t)r, s.d1[r], H, B, H, st, dims->dropout_step));
        GemmEpilogue e2{nullptr, params->res_b2[r], nullptr, 0, 0};
        DCNR_TRY(gemm_any(prec, s.d1[r], H, true, params->res_w2[r], H, true, s.z2[r], H, B, H, H, 1, e2, st, wo.get(wo.w2[r])));
        DCNR_TRY(launch_bn_stats(s.z2[r], H, B, H, dims->bn_eps, dims->bn_momentum, mean2, rstd2, params->res_rm2[r],
                                 params->res_rv2[r], params->res_nbt2[r], s.bn_scratch, st, dims->comm));
        DCNR_TRY(launch_bn_act_fwd(s.z2[r], H, mean2, rstd2, params->res_g2[r], params->res_be2[r], s.h[r], H, nullptr,
                                   0.f, 0, 0, s.h[r + 1], H, B, H, st));
    }
    DCNR_TRY(launch_rowdot_fwd(s.h[dims->n_res], H, params->wf, s.logit_cross, params->bf, logits, B, H, st));
    if (dims->dropout_step != nullptr) DCNR_TRY(launch_inc_u64(dims->dropout_step, st));
    return DCNR_OK;
}

extern "C" int dcnr_backward(const dcnr_dims *dims, const dcnr_params *params, const dcnr_batch *batch,
                             const float *grad_logits, const void *saved, int64_t saved_bytes, const dcnr_grads *grads,
                             void *scratch, int64_t scratch_bytes, dcnr_stream_t stream) {
    DCNR_TRY(check_dims(dims));
    DCNR_REQUIRE(params && batch && grad_logits && saved && grads && scratch, "null argument");
    const int64_t B = batch->batch;
    Arena a(const_cast<void *>(saved), saved_bytes);
    TrainSaved s;
    s.layout(dims, B, a);
    const bool want_tables = grads->user_table || grads->item_table ||
                             std::any_of(grads->cat_table, grads->cat_table + dims->n_cat, [](float *p) { return p; });
    Arena b(scratch, scratch_bytes);
    BwdScratch w;
    w.layout(dims, B, b, want_tables);
    if (!a.ok() || !b.ok()) {
        set_error("backward workspace too small (saved %lld/%lld, scratch %lld/%lld)", (long long)saved_bytes,
                  (long long)a.used, (long long)scratch_bytes, (long long)b.used);
        return DCNR_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    const int H = dims->hidden, D = dims->in_dim, Dp = dims->in_dim_pad, R = dims->n_res, prec = gemm_precision(dims->precision);
    const float post = dims->dropout_p > 0.f ? 1.f / (1.f - dims->dropout_p) : 1.f;

    WeightOps wt;      // transposed weights for the tensor-core dgrads
    DCNR_TRY(wt.prepare(dims, params, s.w0p, w.wsplit, true, st));
    // logit = wf[0:H].h_R + wf[H:].c_L + bf
    float *g = w.ga, *g2 = w.gb, *g3 = w.gc;
    DCNR_TRY(launch_rowdot_bwd(grad_logits, s.h[R], H, params->wf, g, H, grads->wf, grads->bf, B, H, w.bn, st));
    for (int r = R - 1; r >= 0; --r) {
        const float *mean1 = s.stats[r], *rstd1 = mean1 + H, *mean2 = mean1 + 2 * H, *rstd2 = mean1 + 3 * H;
        // out = relu(BN2(z2) + h_in): dy2 (kept in g for the identity path), dz2 -> g2
        DCNR_TRY(launch_bn_act_bwd(g, H, s.h[r + 1], H, s.z2[r], H, mean2, rstd2, params->res_g2[r], 1.f, g2, H, g, H,
                                   grads->res_g2[r], grads->res_be2[r], grads->res_b2[r], B, H, w.bn, st, dims->comm));
        if (grads->res_w2[r])
            DCNR_TRY(launch_linear_wgrad(prec, g2, H, s.d1[r], H, grads->res_w2[r], H, nullptr, B, H, H, H, w.wgrad, st));
        GemmEpilogue none{nullptr, nullptr, nullptr, 0, 0};
        DCNR_TRY(gemm_any(prec, g2, H, true, params->res_w2[r], H, false, g3, H, B, H, H, 1, none, st, wt.get(wt.w2[r])));   // dd1
        // d1 = dropout(relu(BN1(z1))): dz1 in place in g3
        // (block 0: colsum(dz1) is needed for the initial layer's bias gradient even when res_b1 itself is not wanted)
        float *db1 = (r == 0 && grads->b0 != nullptr && grads->res_b1[r] == nullptr) ? w.vec_a : grads->res_b1[r];
        DCNR_TRY(launch_bn_act_bwd(g3, H, s.d1[r], H, s.z1[r], H, mean1, rstd1, params->res_g1[r], post, g3, H, nullptr,
                                   0, grads->res_g1[r], grads->res_be1[r], db1, B, H, w.bn, st, dims->comm));
        if (grads->res_w1[r]) {
            // Block 0 reads h0 = x0 W0^T + b0, the one activation of the tower that no BatchNorm / ReLU has re-centred: its
            // column means are several times its spread (5x on the synthetic states), and dz1 sums to ~0 over the batch, so
            // dz1^T h0 cancels heavily.  The tensor-core kernel takes h0 - mu (mu = column means of a 512-row sample) and the
            // exact rank-1 remainder colsum(dz1) (x) mu is added with the slabs: 4.7e-6 -> 6.9e-7 against float64 on the
            // operands of a real step (profiles/r02_wgrad_centering.md).
            const float *mu = nullptr;
            if (r == 0 && db1 != nullptr && prec == DCNR_PREC_TF32X3) {
                DCNR_TRY(launch_col_mean_sample(s.h[0], H, B, H, w.vec_c, st));
                mu = w.vec_c;
            }
            DCNR_TRY(launch_linear_wgrad(prec, g3, H, s.h[r], H, grads->res_w1[r], H, nullptr, B, H, H, H, w.wgrad, st, mu,
                                         mu != nullptr ? db1 : nullptr));
        }
        GemmEpilogue idn{nullptr, nullptr, g, H, 0};                                                      // + dy2
        DCNR_TRY(gemm_any(prec, g3, H, true, params->res_w1[r], H, false, g2, H, B, H, H, 1, idn, st, wt.get(wt.w1[r])));
        std::swap(g, g2);
    }
    // initial_deep_layer: h0 = x0 W0^T + b0
    // Bias gradient of the initial layer: db0 = colsum(g).  With ResBlocks g = dz1 W1 + dy2 (block 0: colsum(dz1) is the layer1
    // bias gradient, dy2 what `g2` holds after the last swap), so by linearity db0 = colsum(dz1) W1 + colsum(dy2): the batch sum
    // runs over dy2 (elementwise products) instead of over GEMM outputs.  The tensor core's accumulate truncates toward zero
    // (profiles/r02_acc_probe.md), so a GEMM output carries a small error that is sign-correlated along the batch; a later
    // BatchNorm backward removes its column mean, but nothing does for the initial layer, where this cancellation-heavy sum
    // collected it (3.4e-5 against a reference fp32 noise of 9e-6 at B = 4096).  (Rewriting dW0 = W1^T (dz1^T x0) + dy2^T x0 the
    // same way was measured WORSE, 1.6e-5 vs 1.3e-5 at B = 65 536; dW0 is handled by centring x0, below.)
    const bool b0_split = grads->b0 != nullptr && R > 0;
    if (b0_split) {
        const float *db1 = grads->res_b1[0] != nullptr ? grads->res_b1[0] : w.vec_a;
        DCNR_TRY(launch_colsum(g2, H, B, H, w.vec_b, w.bn, st));
        DCNR_TRY(launch_vecmat_add(db1, params->res_w1[0], H, H, H, w.vec_b, grads->b0, st));
    }
    // dW0 = g^T x0 collects the same batch-coherent error through the column means of x0 (the scaled numerics average 0.5):
    // the tensor-core kernel takes x0 - mu and the rank-1 remainder uses the ACCURATE column sums of g from above,
    // dW0 = g^T (x0 - mu) + db0 (x) mu -- identical in exact arithmetic, and the coherent part of g's error now meets a
    // zero-mean operand (1.3e-5 -> see profiles/r02_parity_65536.md at B = 65 536).
    const float *mu0 = nullptr;
    if (b0_split && grads->w0 != nullptr && prec == DCNR_PREC_TF32X3) {
        DCNR_TRY(launch_col_mean_sample(s.x0p, Dp, B, (int32_t)Dp, w.vec_c, st));
        mu0 = w.vec_c;
    }
    if (grads->w0 || (grads->b0 && !b0_split))
        DCNR_TRY(launch_linear_wgrad(prec, g, H, s.x0p, Dp, grads->w0, D, b0_split ? nullptr : grads->b0, B, H, Dp, D, w.wgrad, st,
                                     mu0, mu0 != nullptr ? grads->b0 : nullptr));
    GemmEpilogue none{nullptr, nullptr, nullptr, 0, 0};
    DCNR_TRY(gemm_any(prec, g, H, true, s.w0p, Dp, false, w.dx0, Dp, B, Dp, H, 1, none, st, wt.get(wt.w0)));
    // cross network (recomputed from x0), accumulated onto dx0
    const CrossArgs ca = cross_args(dims, params);
    DCNR_TRY(launch_cross_bwd(s.x0p, Dp, B, ca, Dp, nullptr, 0, grad_logits, params->wf + H, w.dx0, Dp, 1, grads->cross_w,
                              grads->cross_b, grads->wf ? grads->wf + H : nullptr, w.cross_partials, st));
    if (want_tables && dp_sparse_tables(dims)) {
        // data parallel: user / item gradients from the global batch (rank-major = the concatenated batch order);
        // the tiny categorical tables stay local and are summed by the dense gradient all-reduce
        // Every rank contributes `cap` rows (its own batch, zero-padded when shorter): equal NCCL counts on all ranks even for
        // the ragged last batch of an epoch.  dp_batch_cap == 0 declares equal batch sizes.
        const int world = comm_world(dims->comm), E = dims->emb_dim;
        DCNR_REQUIRE(dims->dp_batch_cap == 0 || dims->dp_batch_cap >= B, "dp_batch_cap %lld < local batch %lld",
                     (long long)dims->dp_batch_cap, (long long)B);
        const int64_t cap = dims->dp_batch_cap > B ? dims->dp_batch_cap : B;
        DCNR_TRY(launch_pack_embed_grads(batch->user_ids, batch->item_ids, w.dx0, Dp, B, cap, E, w.pack_ids, w.pack_rows, st));
        DCNR_TRY(comm_allgather(dims->comm, w.pack_ids, w.all_ids, cap * 2 * (int64_t)sizeof(int64_t), st));
        DCNR_TRY(comm_allgather(dims->comm, w.pack_rows, w.all_rows, cap * 2 * E * (int64_t)sizeof(float), st));
        DCNR_TRY(launch_embed_scatter_pair(w.all_ids, 2, dims->n_users, grads->user_table, 0, w.all_ids + 1, 2, dims->n_items,
                                           grads->item_table, E, cap * world, E, w.all_rows, 2 * E, w.scatter, w.scatter_bytes, st));
        dcnr_grads cat_only = *grads;
        cat_only.user_table = cat_only.item_table = nullptr;
        DCNR_TRY(dcnr_embed_scatter_bwd(dims, batch, w.dx0, Dp, &cat_only, w.scatter, w.scatter_bytes, stream));
    } else if (want_tables) {
        DCNR_TRY(dcnr_embed_scatter_bwd(dims, batch, w.dx0, Dp, grads, w.scatter, w.scatter_bytes, stream));
    }
    return DCNR_OK;
}

extern "C" int dcnr_tower_eval_supported(const dcnr_dims *dims) { return dims != nullptr && check_dims(dims) == DCNR_OK && tower_eval_supported(dims) ? 1 : 0; }

extern "C" int64_t dcnr_tower_pack_bytes(const dcnr_dims *dims) {
    if (dims == nullptr || !tower_eval_supported(dims)) return -1;
    return tower_pack_bytes(dims);
}

extern "C" int dcnr_tower_prepare(const dcnr_dims *dims, const dcnr_params *params, int32_t precision, void *pack,
                                  int64_t pack_bytes, dcnr_stream_t stream) {
    DCNR_TRY(check_dims(dims));
    DCNR_REQUIRE(params && pack, "null argument");
    DCNR_REQUIRE(precision == DCNR_PREC_FP16X3 || precision == DCNR_PREC_BF16, "the fused tower runs FP16X3 or BF16");
    DCNR_REQUIRE(tower_eval_supported(dims), "the fused tower needs hidden_dim 256 and 1..4 ResBlocks");
    DCNR_REQUIRE(pack_bytes >= tower_pack_bytes(dims) && ((uintptr_t)pack & 255) == 0, "pack buffer too small or unaligned");
    return launch_tower_prepare(dims, params, static_cast<char *>(pack), precision, as_stream(stream));
}

extern "C" int64_t dcnr_tower_eval_workspace_bytes(const dcnr_dims *dims) {
    if (dims == nullptr || !tower_eval_supported(dims)) return -1;
    return tower_pack_bytes(dims) + 256;
}

extern "C" int dcnr_tower_eval(const dcnr_dims *dims, const dcnr_params *params, const float *x0, int64_t ldx0,
                               const float *logit_cross, float *logits, int64_t m, int32_t precision, int32_t options,
                               int32_t *flags, void *workspace, int64_t workspace_bytes, dcnr_stream_t stream) {
    DCNR_TRY(check_dims(dims));
    DCNR_REQUIRE(params && x0 && logits && workspace, "null argument");
    DCNR_REQUIRE(tower_eval_supported(dims), "the fused tower needs hidden_dim 256 and 1..4 ResBlocks");
    Arena a(workspace, workspace_bytes);
    char *pack = a.take<char>(tower_pack_bytes(dims));
    if (!a.ok()) {
        set_error("tower workspace too small (%lld < %lld)", (long long)workspace_bytes, (long long)a.used);
        return DCNR_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    DCNR_TRY(launch_tower_prepare(dims, params, pack, precision, st));
    return launch_tower_eval(dims, x0, ldx0, logit_cross, params->bf, pack, logits, m, flags, precision, options, st);
}

