// K1/K2: fused embedding gather + concat (+ cross network + cross half of the final dot), and
// the cross-network backward with forward recomputation.
//
// Reference lines: gather+concat train.py:156-159 (main.py:116-119); CrossLayer train.py:96-99
// (main.py:67-70) applied n_cross times train.py:167-168; final dot train.py:169-170.
//
// Layout: one row of x0 is handled by an aligned group of 8 lanes; lane l owns the float4 column
// quads {32*j + 4*l .. +3 : j < NV}, NV = in_dim_pad/32.  A warp therefore touches 4 rows per
// instruction and every float4 store of the padded row hits full 128-byte lines.  The row lives in
// registers across all cross layers (dot products by 3-step xor shuffles), so x0 is written once
// and the cross activations never touch HBM.
#include "kernels.cuh"

namespace dcnr {

constexpr int kThreads = 256;
constexpr int kGroupsPerCta = kThreads / 8;

struct __align__(8) SmemSeg {
    const float *table;
    const int64_t *ids;
    int64_t rows;
    int32_t id_stride, width;
};

template <int NV>
__global__ void __launch_bounds__(kThreads)
k_embed_cross_fwd(GatherArgs ga, const float *__restrict__ x_in, int64_t ldx_in, int64_t B, CrossArgs ca,
                  float *__restrict__ x0_out, int64_t ldx0, float *__restrict__ y_out, int64_t ldy,
                  const float *__restrict__ wf_cross, float *__restrict__ logit_part, int32_t *err_flag) {
    constexpr int DP = NV * 32;
    constexpr int NE = NV * 4;
    extern __shared__ __align__(16) float smem[];
    float *sw = smem;                    // [L][DP]
    float *sb = sw + ca.L * DP;          // [L][DP]
    float *swf = sb + ca.L * DP;         // [DP]
    __shared__ SmemSeg sseg[2 + DCNR_MAX_CAT];

    const int tid = threadIdx.x;
    for (int i = tid; i < ca.L * DP; i += kThreads) {
        int l = i / DP, c = i % DP;
        sw[i] = c < ca.D ? ca.w[l][c] : 0.f;
        sb[i] = c < ca.D ? ca.b[l][c] : 0.f;
    }
    for (int c = tid; c < DP; c += kThreads) swf[c] = (wf_cross != nullptr && c < ca.D) ? wf_cross[c] : 0.f;
    if (tid < ga.n_seg) {
        sseg[tid].table = ga.seg[tid].table;
        sseg[tid].ids = ga.seg[tid].ids;
        sseg[tid].rows = ga.seg[tid].rows;
        sseg[tid].id_stride = ga.seg[tid].id_stride;
        sseg[tid].width = ga.seg[tid].width;
    }
    __syncthreads();

    const int lane8 = tid & 7;
    // per-element column metadata (row invariant): kind >= 0 table segment, -1 numeric, -2 pad
    int kind[NE], off[NE];
#pragma unroll
    for (int k = 0; k < NE; ++k) {
        int col = 32 * (k >> 2) + 4 * lane8 + (k & 3);
        kind[k] = -2;
        off[k] = 0;
        if (x_in == nullptr) {
            if (col >= ga.num_col0 && col < ga.D) {
                kind[k] = -1;
                off[k] = col - ga.num_col0;
            } else {
                for (int s = 0; s < ga.n_seg; ++s) {
                    int c0 = ga.seg[s].col0;
                    if (col >= c0 && col < c0 + ga.seg[s].width) {
                        kind[k] = s;
                        off[k] = col - c0;
                    }
                }
            }
        }
    }

    const bool x0_vec = x0_out != nullptr && ldx0 >= DP && (ldx0 & 3) == 0;
    const bool y_vec = y_out != nullptr && ldy >= DP && (ldy & 3) == 0;
    const int D = (x_in != nullptr) ? ca.D : ga.D;
    const int64_t group0 = (int64_t)blockIdx.x * kGroupsPerCta + (tid >> 3);
    const int64_t stride = (int64_t)gridDim.x * kGroupsPerCta;
    // warp-uniform trip count: all 4 groups of a warp iterate together (shuffles need full warps)
    const int64_t warp_first = group0 - ((tid >> 3) & 3);
    bool bad_id = false;

    // a column quad that lies inside one table segment at a 16-byte aligned offset is fetched with one id load and
    // one 128-bit row load (user / item embeddings: 8 of the 16 quads of a P0 row); the rest goes element by element
    int qseg[NV], qoff[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        qseg[j] = -1;
        qoff[j] = 0;
        if (x_in == nullptr && kind[4 * j] >= 0 && kind[4 * j] == kind[4 * j + 3] && (off[4 * j] & 3) == 0 &&
            (ga.seg[kind[4 * j]].width & 3) == 0 && (reinterpret_cast<uintptr_t>(ga.seg[kind[4 * j]].table) & 15) == 0) {
            qseg[j] = kind[4 * j];
            qoff[j] = off[4 * j];
        }
    }
    // the element-wise metadata is only needed on the mixed quads: keep it in shared memory, not in 2 * NE registers
    __shared__ int smeta[8][NV * 4];
    if ((tid >> 3) == 0) {
#pragma unroll
        for (int k = 0; k < NE; ++k) smeta[lane8][k] = ((kind[k] + 2) & 0xff) | (off[k] << 8);
    }
    __syncthreads();
    // ids of the vector quads are fetched one row AHEAD, so a row costs one exposed memory round trip (the table
    // rows) instead of two dependent ones (id, then row)
    int64_t idv[NV];
    auto load_ids = [&](int64_t r, int64_t (&dst)[NV]) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            dst[j] = 0;
            if (qseg[j] >= 0 && r < B) dst[j] = __ldg(sseg[qseg[j]].ids + r * sseg[qseg[j]].id_stride);
        }
    };
    if (x_in == nullptr) load_ids(warp_first + ((tid >> 3) & 3), idv);
    for (int64_t base = warp_first; base < B; base += stride) {
        const int64_t row = base + ((tid >> 3) & 3);
        const bool active = row < B;
        float x[NE];
        if (x_in != nullptr) {
#pragma unroll
            for (int k = 0; k < NE; ++k) {
                int col = 32 * (k >> 2) + 4 * lane8 + (k & 3);
                x[k] = (active && col < D) ? __ldg(x_in + row * ldx_in + col) : 0.f;
            }
        } else {
            int64_t idn[NV];
            load_ids(row + stride, idn);
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                if (qseg[j] >= 0) {
                    float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (active) {
                        const SmemSeg &s = sseg[qseg[j]];
                        int64_t id = idv[j];
                        if ((uint64_t)id >= (uint64_t)s.rows) {
                            bad_id = true;
                            id = 0;
                        }
                        v4 = ldg4(s.table + id * s.width + qoff[j]);
                    }
                    x[4 * j] = v4.x; x[4 * j + 1] = v4.y; x[4 * j + 2] = v4.z; x[4 * j + 3] = v4.w;
                    continue;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int k = 4 * j + e;
                    const int meta = smeta[lane8][k];
                    const int kd = (meta & 0xff) - 2, of = meta >> 8;
                    float v = 0.f;
                    if (active) {
                        if (kd >= 0) {
                            const SmemSeg &s = sseg[kd];
                            int64_t id = __ldg(s.ids + row * s.id_stride);
                            if ((uint64_t)id >= (uint64_t)s.rows) {
                                bad_id = true;
                                id = 0;
                            }
                            v = __ldg(s.table + id * s.width + of);
                        } else if (kd == -1) {
                            v = __ldg(ga.num + row * ga.n_num + of);
                        }
                    }
                    x[k] = v;
                }
            }
#pragma unroll
            for (int j = 0; j < NV; ++j) idv[j] = idn[j];
            if (x0_out != nullptr && active) {
                if (x0_vec) {
#pragma unroll
                    for (int j = 0; j < NV; ++j)
                        st4(x0_out + row * ldx0 + 32 * j + 4 * lane8,
                            make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]));
                } else {
#pragma unroll
                    for (int k = 0; k < NE; ++k) {
                        int col = 32 * (k >> 2) + 4 * lane8 + (k & 3);
                        if (col < ldx0 && col < DP) x0_out[row * ldx0 + col] = x[k];
                    }
                }
            }
        }
        // cross layers: x <- x * (1 + x.w_l) + b_l
        for (int l = 0; l < ca.L; ++l) {
            const float *wl = sw + l * DP, *bl = sb + l * DP;
            float p = 0.f;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                float4 w4 = *reinterpret_cast<const float4 *>(wl + 32 * j + 4 * lane8);
                p = fmaf(x[4 * j + 0], w4.x, p);
                p = fmaf(x[4 * j + 1], w4.y, p);
                p = fmaf(x[4 * j + 2], w4.z, p);
                p = fmaf(x[4 * j + 3], w4.w, p);
            }
            const float s = group8_sum(p);
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                float4 b4 = *reinterpret_cast<const float4 *>(bl + 32 * j + 4 * lane8);
                x[4 * j + 0] = fmaf(x[4 * j + 0], s, x[4 * j + 0]) + b4.x;
                x[4 * j + 1] = fmaf(x[4 * j + 1], s, x[4 * j + 1]) + b4.y;
                x[4 * j + 2] = fmaf(x[4 * j + 2], s, x[4 * j + 2]) + b4.z;
                x[4 * j + 3] = fmaf(x[4 * j + 3], s, x[4 * j + 3]) + b4.w;
            }
        }
        if (y_out != nullptr && active) {
            if (y_vec) {
#pragma unroll
                for (int j = 0; j < NV; ++j)
                    st4(y_out + row * ldy + 32 * j + 4 * lane8,
                        make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]));
            } else {
#pragma unroll
                for (int k = 0; k < NE; ++k) {
                    int col = 32 * (k >> 2) + 4 * lane8 + (k & 3);
                    if (col < D && col < ldy) y_out[row * ldy + col] = x[k];
                }
            }
        }
        if (logit_part != nullptr) {
            float p = 0.f;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                float4 w4 = *reinterpret_cast<const float4 *>(swf + 32 * j + 4 * lane8);
                p = fmaf(x[4 * j + 0], w4.x, p);
                p = fmaf(x[4 * j + 1], w4.y, p);
                p = fmaf(x[4 * j + 2], w4.z, p);
                p = fmaf(x[4 * j + 3], w4.w, p);
            }
            const float s = group8_sum(p);
            if (active && lane8 == 0) logit_part[row] = s;
        }
    }
    if (bad_id && err_flag != nullptr) atomicExch(err_flag, 1);
}

int make_gather_args(const dcnr_dims *dims, const dcnr_params *params, const dcnr_batch *batch, GatherArgs *out) {
    DCNR_REQUIRE(dims->n_cat >= 0 && dims->n_cat <= DCNR_MAX_CAT, "n_cat %d > %d", dims->n_cat, DCNR_MAX_CAT);
    DCNR_REQUIRE(batch->user_ids && batch->item_ids && (dims->n_cat == 0 || batch->cat_features) &&
                     (dims->n_num == 0 || batch->num_features),
                 "null batch tensor");
    memset(out, 0, sizeof(*out));
    int col = 0, n = 0;
    auto add = [&](const float *table, const int64_t *ids, int64_t stride, int64_t rows, int32_t width) {
        GatherSeg &s = out->seg[n++];
        s.table = table; s.ids = ids; s.id_stride = (int32_t)stride; s.rows = rows; s.width = width; s.col0 = col;
        col += width;
    };
    add(params ? params->user_table : nullptr, batch->user_ids, 1, dims->n_users, dims->emb_dim);
    add(params ? params->item_table : nullptr, batch->item_ids, 1, dims->n_items, dims->emb_dim);
    for (int i = 0; i < dims->n_cat; ++i)
        add(params ? params->cat_table[i] : nullptr, batch->cat_features + i, dims->n_cat, dims->cat_rows[i],
            dims->cat_width[i]);
    out->n_seg = n;
    out->num = batch->num_features;
    out->n_num = dims->n_num;
    out->num_col0 = col;
    out->D = col + dims->n_num;
    DCNR_REQUIRE(out->D == dims->in_dim, "in_dim %d != 2E + sum(cat_width) + n_num = %d", dims->in_dim, out->D);
    DCNR_REQUIRE(dims->in_dim_pad == (int32_t)round_up(dims->in_dim, DCNR_PAD), "in_dim_pad must be round_up(in_dim, %d)",
                 DCNR_PAD);
    return DCNR_OK;
}

int launch_embed_cross_fwd(const GatherArgs *ga, const float *x_in, int64_t ldx_in, int64_t B,
                           const CrossArgs &ca, int32_t dim_pad, float *x0_out, int64_t ldx0, float *y_out,
                           int64_t ldy, const float *wf_cross, float *logit_part, int32_t *err_flag,
                           cudaStream_t stream) {
    if (B <= 0) return DCNR_OK;
    GatherArgs g0;
    if (ga == nullptr) {
        memset(&g0, 0, sizeof(g0));
        ga = &g0;
    }
    const int nv = dim_pad / 32;
    DCNR_REQUIRE(dim_pad % 32 == 0 && nv >= 1 && nv <= 8, "in_dim_pad %d unsupported (must be 32..256, multiple of 32)",
                 dim_pad);
    const size_t smem = (size_t)(2 * ca.L + 1) * dim_pad * sizeof(float);
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(B, kGroupsPerCta), (int64_t)sm_count() * 8);
#define DCNR_CASE(NVV)                                                                                      \
    case NVV:                                                                                               \
        k_embed_cross_fwd<NVV><<<grid, kThreads, smem, stream>>>(*ga, x_in, ldx_in, B, ca, x0_out, ldx0,   \
                                                                  y_out, ldy, wf_cross, logit_part, err_flag); \
        break;
    switch (nv) {
        DCNR_CASE(1) DCNR_CASE(2) DCNR_CASE(3) DCNR_CASE(4) DCNR_CASE(5) DCNR_CASE(6) DCNR_CASE(7) DCNR_CASE(8)
    }
#undef DCNR_CASE
    DCNR_LAUNCHED();
    return DCNR_OK;
}

// ------------------------------------------------------------------------------------------------
// Cross backward.  For one row with layer inputs c_0 = x, c_{l+1} = c_l (1 + s_l) + b_l, s_l = c_l.w_l
// and upstream g = dL/dc_L:
//     t_l = g . c_l ;  gw_l += t_l c_l ;  gb_l += g ;  g <- g (1 + s_l) + w_l t_l
// c_l is recomputed from x (O(L^2) cheap ALU work instead of saving L activations).  Per-lane
// accumulators for gw/gb are reduced over the CTA's 32 row groups in group order and written as
// one partial row per CTA; a finalize kernel adds the CTAs in order (deterministic).
// ------------------------------------------------------------------------------------------------
template <int NV, int LMAX>
__global__ void __launch_bounds__(kThreads)
k_cross_bwd(const float *__restrict__ x_in, int64_t ldx, int64_t B, CrossArgs ca, const float *__restrict__ gy,
            int64_t ldg, const float *__restrict__ dlogit, const float *__restrict__ wf_cross,
            float *__restrict__ gx, int64_t ldgx, int accumulate, float *__restrict__ partials) {
    constexpr int DP = NV * 32;
    constexpr int NE = NV * 4;
    extern __shared__ __align__(16) float smem[];
    float *sw = smem;                    // [L][DP]
    float *sb = sw + ca.L * DP;          // [L][DP]
    float *swf = sb + ca.L * DP;         // [DP]
    float *red = swf + DP;               // [kGroupsPerCta][DP]
    const int tid = threadIdx.x;
    const int L = ca.L;
    for (int i = tid; i < L * DP; i += kThreads) {
        int l = i / DP, c = i % DP;
        sw[i] = c < ca.D ? ca.w[l][c] : 0.f;
        sb[i] = c < ca.D ? ca.b[l][c] : 0.f;
    }
    for (int c = tid; c < DP; c += kThreads) swf[c] = (wf_cross != nullptr && c < ca.D) ? wf_cross[c] : 0.f;
    __syncthreads();

    const int lane8 = tid & 7;
    const int grp = tid >> 3;
    float acc_w[LMAX][NE], acc_b[LMAX][NE], acc_f[NE];
#pragma unroll
    for (int l = 0; l < LMAX; ++l)
#pragma unroll
        for (int k = 0; k < NE; ++k) acc_w[l][k] = acc_b[l][k] = 0.f;
#pragma unroll
    for (int k = 0; k < NE; ++k) acc_f[k] = 0.f;

    const bool x_vec = ldx >= DP && (ldx & 3) == 0;
    const bool gx_vec = ldgx >= DP && (ldgx & 3) == 0;
    const int64_t group0 = (int64_t)blockIdx.x * kGroupsPerCta + grp;
    const int64_t stride = (int64_t)gridDim.x * kGroupsPerCta;
    const int64_t warp_first = group0 - (grp & 3);

    auto layer_fwd = [&](float (&c)[NE], int l) {
        const float *wl = sw + l * DP, *bl = sb + l * DP;
        float p = 0.f;
#pragma unroll
        for (int k = 0; k < NE; ++k) p = fmaf(c[k], wl[32 * (k >> 2) + 4 * lane8 + (k & 3)], p);
        const float s = group8_sum(p);
#pragma unroll
        for (int k = 0; k < NE; ++k) c[k] = fmaf(c[k], s, c[k]) + bl[32 * (k >> 2) + 4 * lane8 + (k & 3)];
    };

    for (int64_t base = warp_first; base < B; base += stride) {
        const int64_t row = base + (grp & 3);
        const bool active = row < B;
        float x[NE], g[NE], c[NE];
        if (x_vec) {
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                float4 v = active ? ldg4(x_in + row * ldx + 32 * j + 4 * lane8) : make_float4(0.f, 0.f, 0.f, 0.f);
                x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
            }
#pragma unroll
            for (int k = 0; k < NE; ++k)
                if (32 * (k >> 2) + 4 * lane8 + (k & 3) >= ca.D) x[k] = 0.f;   // pad columns may hold anything
        } else {
#pragma unroll
            for (int k = 0; k < NE; ++k) {
                int col = 32 * (k >> 2) + 4 * lane8 + (k & 3);
                x[k] = (active && col < ca.D) ? __ldg(x_in + row * ldx + col) : 0.f;
            }
        }
        if (gy != nullptr) {
#pragma unroll
            for (int k = 0; k < NE; ++k) {
                int col = 32 * (k >> 2) + 4 * lane8 + (k & 3);
                g[k] = (active && col < ca.D) ? __ldg(gy + row * ldg + col) : 0.f;
            }
        } else {
            const float dl = active ? __ldg(dlogit + row) : 0.f;
#pragma unroll
            for (int k = 0; k < NE; ++k) c[k] = x[k];
            for (int l = 0; l < L; ++l) layer_fwd(c, l);
#pragma unroll
            for (int k = 0; k < NE; ++k) {
                g[k] = dl * swf[32 * (k >> 2) + 4 * lane8 + (k & 3)];
                acc_f[k] = fmaf(dl, c[k], acc_f[k]);           // d wf_cross = sum_b dlogit_b * c_L[b]
            }
        }
#pragma unroll
        for (int l = LMAX - 1; l >= 0; --l) {
            if (l < L) {
#pragma unroll
                for (int k = 0; k < NE; ++k) c[k] = x[k];
#pragma unroll
                for (int j = 0; j < LMAX; ++j)
                    if (j < l) layer_fwd(c, j);
                const float *wl = sw + l * DP;
                float ps = 0.f, pt = 0.f;
#pragma unroll
                for (int k = 0; k < NE; ++k) {
                    ps = fmaf(c[k], wl[32 * (k >> 2) + 4 * lane8 + (k & 3)], ps);
                    pt = fmaf(g[k], c[k], pt);
                }
                const float s = group8_sum(ps);
                const float t = group8_sum(pt);
#pragma unroll
                for (int k = 0; k < NE; ++k) {
                    acc_w[l][k] = fmaf(t, c[k], acc_w[l][k]);
                    acc_b[l][k] += g[k];
                    g[k] = fmaf(g[k], s, g[k]) + wl[32 * (k >> 2) + 4 * lane8 + (k & 3)] * t;
                }
            }
        }
        if (active) {
            if (gx_vec) {
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    float *p = gx + row * ldgx + 32 * j + 4 * lane8;
                    float4 v = make_float4(g[4 * j], g[4 * j + 1], g[4 * j + 2], g[4 * j + 3]);
                    if (accumulate) {
                        float4 o = *reinterpret_cast<const float4 *>(p);
                        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                    }
                    st4(p, v);
                }
            } else {
#pragma unroll
                for (int k = 0; k < NE; ++k) {
                    int col = 32 * (k >> 2) + 4 * lane8 + (k & 3);
                    if (col < ca.D && col < ldgx) {
                        float v = g[k];
                        if (accumulate) v += gx[row * ldgx + col];
                        gx[row * ldgx + col] = v;
                    }
                }
            }
        }
    }

    // CTA reduction in group order, one vector (w_l, b_l, wf) at a time
    float *prow = partials + (int64_t)blockIdx.x * (2 * L + 1) * DP;
    for (int v = 0; v < 2 * L + 1; ++v) {
        const int l = v >> 1;
#pragma unroll
        for (int k = 0; k < NE; ++k) {
            float val = acc_f[k];
#pragma unroll
            for (int ll = 0; ll < LMAX; ++ll)
                if (ll == l && v < 2 * L) val = (v & 1) ? acc_b[ll][k] : acc_w[ll][k];
            red[grp * DP + 32 * (k >> 2) + 4 * lane8 + (k & 3)] = val;
        }
        __syncthreads();
        for (int c0 = tid; c0 < DP; c0 += kThreads) {
            float s = 0.f;
            for (int gq = 0; gq < kGroupsPerCta; ++gq) s += red[gq * DP + c0];
            prow[v * DP + c0] = s;
        }
        __syncthreads();
    }
}

int cross_bwd_grid(int64_t B) {
    return (int)std::min<int64_t>(std::max<int64_t>(ceil_div(B, kGroupsPerCta), 1), (int64_t)sm_count() * 2);
}

int64_t cross_bwd_partial_floats(int64_t B, int32_t dim_pad, int32_t L) {
    return (int64_t)cross_bwd_grid(B) * (2 * L + 1) * dim_pad;
}

int launch_cross_bwd(const float *x, int64_t ldx, int64_t B, const CrossArgs &ca, int32_t dim_pad, const float *gy,
                     int64_t ldg, const float *dlogit, const float *wf_cross, float *gx, int64_t ldgx,
                     int accumulate, float *const *gw, float *const *gb, float *gwf, float *partials,
                     cudaStream_t stream) {
    if (B <= 0) return DCNR_OK;
    const int nv = dim_pad / 32;
    const int L = ca.L;
    DCNR_REQUIRE(dim_pad % 32 == 0 && nv >= 1 && nv <= 8, "in_dim_pad %d unsupported", dim_pad);
    DCNR_REQUIRE(L >= 0 && L <= DCNR_MAX_CROSS, "n_cross %d unsupported", L);
    const int grid = cross_bwd_grid(B);
    const size_t smem = (size_t)((2 * L + 1) * dim_pad + kGroupsPerCta * dim_pad) * sizeof(float);
#define DCNR_CASE2(NVV, LM)                                                                                   \
    k_cross_bwd<NVV, LM><<<grid, kThreads, smem, stream>>>(x, ldx, B, ca, gy, ldg, dlogit, wf_cross, gx, ldgx, \
                                                           accumulate, partials)
#define DCNR_CASE(NVV)                                  \
    case NVV:                                           \
        if (L <= 4) DCNR_CASE2(NVV, 4);                 \
        else DCNR_CASE2(NVV, 8);                        \
        break;
    switch (nv) {
        DCNR_CASE(1) DCNR_CASE(2) DCNR_CASE(3) DCNR_CASE(4) DCNR_CASE(5) DCNR_CASE(6) DCNR_CASE(7) DCNR_CASE(8)
    }
#undef DCNR_CASE
#undef DCNR_CASE2
    DCNR_LAUNCHED();
    SegPtrs seg;
    memset(&seg, 0, sizeof(seg));
    int n = 0;
    for (int l = 0; l < L; ++l) {
        seg.out[n] = gw ? gw[l] : nullptr; seg.offset[n] = (2 * l) * dim_pad; seg.len[n] = ca.D; ++n;
        seg.out[n] = gb ? gb[l] : nullptr; seg.offset[n] = (2 * l + 1) * dim_pad; seg.len[n] = ca.D; ++n;
    }
    seg.out[n] = gwf; seg.offset[n] = 2 * L * dim_pad; seg.len[n] = ca.D; ++n;
    seg.n = n;
    return launch_sum_partials(partials, grid, (int64_t)(2 * L + 1) * dim_pad, seg, stream);
}

}  // namespace dcnr

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace dcnr;

extern "C" int dcnr_embed_concat_fwd(const dcnr_dims *dims, const dcnr_params *params, const dcnr_batch *batch,
                                     float *x0, int64_t ldx0, dcnr_stream_t stream) {
    DCNR_REQUIRE(dims && params && batch && x0, "null argument");
    DCNR_REQUIRE(ldx0 >= dims->in_dim, "ldx0 %lld < in_dim %d", (long long)ldx0, dims->in_dim);
    GatherArgs ga;
    DCNR_TRY(make_gather_args(dims, params, batch, &ga));
    CrossArgs ca;
    memset(&ca, 0, sizeof(ca));
    ca.D = dims->in_dim;
    return launch_embed_cross_fwd(&ga, nullptr, 0, batch->batch, ca, dims->in_dim_pad, x0, ldx0, nullptr, 0, nullptr,
                                  nullptr, nullptr, as_stream(stream));
}

extern "C" int dcnr_cross_fwd(const float *x, int64_t ldx, int64_t batch, int32_t dim, int32_t n_layers,
                              const float *const *w_host, const float *const *b_host, float *y, int64_t ldy,
                              dcnr_stream_t stream) {
    DCNR_REQUIRE(x && y && (n_layers == 0 || (w_host && b_host)), "null argument");
    DCNR_REQUIRE(n_layers >= 0 && n_layers <= DCNR_MAX_CROSS, "n_layers %d > %d", n_layers, DCNR_MAX_CROSS);
    DCNR_REQUIRE(dim >= 1 && dim <= 256 && ldx >= dim && ldy >= dim, "bad dim/ld");
    CrossArgs ca;
    memset(&ca, 0, sizeof(ca));
    ca.L = n_layers;
    ca.D = dim;
    for (int l = 0; l < n_layers; ++l) { ca.w[l] = w_host[l]; ca.b[l] = b_host[l]; }
    return launch_embed_cross_fwd(nullptr, x, ldx, batch, ca, (int32_t)round_up(dim, DCNR_PAD), nullptr, 0, y, ldy,
                                  nullptr, nullptr, nullptr, as_stream(stream));
}

extern "C" int64_t dcnr_cross_bwd_scratch_bytes(int64_t batch, int32_t dim, int32_t n_layers) {
    return round_up(cross_bwd_partial_floats(batch, (int32_t)round_up(dim, DCNR_PAD), n_layers) * 4, 256);
}

extern "C" int dcnr_cross_bwd(const float *x, int64_t ldx, int64_t batch, int32_t dim, int32_t n_layers,
                              const float *const *w_host, const float *const *b_host, const float *gy, int64_t ldg,
                              float *gx, int64_t ldgx, float *const *gw_host, float *const *gb_host, void *scratch,
                              int64_t scratch_bytes, dcnr_stream_t stream) {
    DCNR_REQUIRE(x && gy && gx && w_host && b_host && scratch, "null argument");
    DCNR_REQUIRE(n_layers >= 1 && n_layers <= DCNR_MAX_CROSS, "n_layers %d out of range", n_layers);
    DCNR_REQUIRE(dim >= 1 && dim <= 256 && ldx >= dim && ldg >= dim && ldgx >= dim, "bad dim/ld");
    if (scratch_bytes < dcnr_cross_bwd_scratch_bytes(batch, dim, n_layers)) {
        set_error("cross_bwd scratch too small");
        return DCNR_ERR_WORKSPACE;
    }
    CrossArgs ca;
    memset(&ca, 0, sizeof(ca));
    ca.L = n_layers;
    ca.D = dim;
    for (int l = 0; l < n_layers; ++l) { ca.w[l] = w_host[l]; ca.b[l] = b_host[l]; }
    return launch_cross_bwd(x, ldx, batch, ca, (int32_t)round_up(dim, DCNR_PAD), gy, ldg, nullptr, nullptr, gx, ldgx,
                            0, gw_host, gb_host, nullptr, reinterpret_cast<float *>(scratch), as_stream(stream));
}
