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
    int32_t id_stride, width, col0, pad_;
};

template <int NV, int R>
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
        sseg[tid].col0 = ga.seg[tid].col0;
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
    auto load_ids = [&](int64_t r, int64_t (&dst)[NV]) {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            dst[j] = 0;
            if (qseg[j] >= 0 && r < B) dst[j] = __ldg(sseg[qseg[j]].ids + r * sseg[qseg[j]].id_stride);
        }
    };
    // R rows per 8-lane group per iteration (rows base + r * stride): every load of all R rows is issued before the first
    // use, so a warp keeps R x 4 rows of gathers in flight (the kernel is latency-bound: bytes in flight per SM set the rate)
    int64_t idv[R][NV];
    if (x_in == nullptr) {
#pragma unroll
        for (int r = 0; r < R; ++r) load_ids(warp_first + ((tid >> 3) & 3) + r * stride, idv[r]);
    }
    for (int64_t base = warp_first; base < B; base += R * stride) {
        int64_t row[R];
        bool active[R];
        float x[R][NE];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            row[r] = base + ((tid >> 3) & 3) + r * stride;
            active[r] = row[r] < B;
        }
        if (x_in != nullptr) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
#pragma unroll
                for (int k = 0; k < NE; ++k) {
                    int col = 32 * (k >> 2) + 4 * lane8 + (k & 3);
                    x[r][k] = (active[r] && col < D) ? __ldg(x_in + row[r] * ldx_in + col) : 0.f;
                }
            }
        } else {
            int64_t idn[R][NV];
#pragma unroll
            for (int r = 0; r < R; ++r) load_ids(row[r] + R * stride, idn[r]);
#pragma unroll
            for (int r = 0; r < R; ++r) {
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    if (qseg[j] >= 0) {
                        float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (active[r]) {
                            const SmemSeg &sg = sseg[qseg[j]];
                            int64_t id = idv[r][j];
                            if ((uint64_t)id >= (uint64_t)sg.rows) {
                                bad_id = true;
                                id = 0;
                            }
                            v4 = ldg4(sg.table + id * sg.width + qoff[j]);
                        }
                        x[r][4 * j] = v4.x; x[r][4 * j + 1] = v4.y; x[r][4 * j + 2] = v4.z; x[r][4 * j + 3] = v4.w;
                        continue;
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int k = 4 * j + e;
                        const int meta = smeta[lane8][k];
                        const int kd = (meta & 0xff) - 2, of = meta >> 8;
                        float v = 0.f;
                        if (active[r]) {
                            if (kd >= 0) {
                                const SmemSeg &sg = sseg[kd];
                                int64_t id = __ldg(sg.ids + row[r] * sg.id_stride);
                                if ((uint64_t)id >= (uint64_t)sg.rows) {
                                    bad_id = true;
                                    id = 0;
                                }
                                v = __ldg(sg.table + id * sg.width + of);
                            } else if (kd == -1) {
                                v = __ldg(ga.num + row[r] * ga.n_num + of);
                            }
                        }
                        x[r][k] = v;
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
#pragma unroll
                for (int j = 0; j < NV; ++j) idv[r][j] = idn[r][j];
                if (x0_out != nullptr && active[r]) {
                    if (x0_vec) {
#pragma unroll
                        for (int j = 0; j < NV; ++j)
                            st4(x0_out + row[r] * ldx0 + 32 * j + 4 * lane8,
                                make_float4(x[r][4 * j], x[r][4 * j + 1], x[r][4 * j + 2], x[r][4 * j + 3]));
                    } else {
#pragma unroll
                        for (int k = 0; k < NE; ++k) {
                            int col = 32 * (k >> 2) + 4 * lane8 + (k & 3);
                            if (col < ldx0 && col < DP) x0_out[row[r] * ldx0 + col] = x[r][k];
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            // cross layers: x <- x * (1 + x.w_l) + b_l
            for (int l = 0; l < ca.L; ++l) {
                const float *wl = sw + l * DP, *bl = sb + l * DP;
                float p = 0.f;
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    float4 w4 = *reinterpret_cast<const float4 *>(wl + 32 * j + 4 * lane8);
                    p = fmaf(x[r][4 * j + 0], w4.x, p);
                    p = fmaf(x[r][4 * j + 1], w4.y, p);
                    p = fmaf(x[r][4 * j + 2], w4.z, p);
                    p = fmaf(x[r][4 * j + 3], w4.w, p);
                }
                const float s = group8_sum(p);
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    float4 b4 = *reinterpret_cast<const float4 *>(bl + 32 * j + 4 * lane8);
                    x[r][4 * j + 0] = fmaf(x[r][4 * j + 0], s, x[r][4 * j + 0]) + b4.x;
                    x[r][4 * j + 1] = fmaf(x[r][4 * j + 1], s, x[r][4 * j + 1]) + b4.y;
                    x[r][4 * j + 2] = fmaf(x[r][4 * j + 2], s, x[r][4 * j + 2]) + b4.z;
                    x[r][4 * j + 3] = fmaf(x[r][4 * j + 3], s, x[r][4 * j + 3]) + b4.w;
                }
            }
            if (y_out != nullptr && active[r]) {
                if (y_vec) {
#pragma unroll
                    for (int j = 0; j < NV; ++j)
                        st4(y_out + row[r] * ldy + 32 * j + 4 * lane8,
                            make_float4(x[r][4 * j], x[r][4 * j + 1], x[r][4 * j + 2], x[r][4 * j + 3]));
                } else {
#pragma unroll
                    for (int k = 0; k < NE; ++k) {
                        int col = 32 * (k >> 2) + 4 * lane8 + (k & 3);
                        if (col < D && col < ldy) y_out[row[r] * ldy + col] = x[r][k];
                    }
                }
            }
            if (logit_part != nullptr) {
                float p = 0.f;
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    float4 w4 = *reinterpret_cast<const float4 *>(swf + 32 * j + 4 * lane8);
                    p = fmaf(x[r][4 * j + 0], w4.x, p);
                    p = fmaf(x[r][4 * j + 1], w4.y, p);
                    p = fmaf(x[r][4 * j + 2], w4.z, p);
                    p = fmaf(x[r][4 * j + 3], w4.w, p);
                }
                const float s = group8_sum(p);
                if (active[r] && lane8 == 0) logit_part[row[r]] = s;
            }
        }
    }
    if (bad_id && err_flag != nullptr) atomicOr(err_flag, 1);
}

// ------------------------------------------------------------------------------------------------
// Staged form of the gather (the default when the row layout allows it).  The element-by-element path
// above spends ~250 of its ~350 warp instructions per 4 rows on the columns that are not 16-byte
// aligned table quads (categorical rows of odd width, the numerics, the pad): the kernel was
// issue-bound at ~50 % of HBM bandwidth.  Here a warp takes 32 consecutive rows per step:
//   1. lane i loads the ids of row i (coalesced, bounds-checked once per id);
//   2. the "mixed" columns [mix0, DP) of the 32 rows are assembled in a shared-memory tile with
//      flat, coalesced loops (numerics: 32 x n_num contiguous floats; a categorical table: 32 x width
//      elements, the row id fetched by shuffle) -- ~4 instructions per 32 elements;
//   3. 8 passes of 4 rows: the aligned prefix segments (user / item rows) come straight from the
//      tables as 128-bit loads (ids by shuffle), the rest as float4 reads of the tile; then exactly
//      the register-resident cross layers / stores of the form above.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxVecSeg = 4;

template <int NV, int PF, int MB>
__global__ void __launch_bounds__(kThreads, MB)
k_embed_cross_fwd_staged(GatherArgs ga, int n_vec, int mix0, int64_t B, CrossArgs ca, float *__restrict__ x0_out,
                         int64_t ldx0, float *__restrict__ y_out, int64_t ldy, const float *__restrict__ wf_cross,
                         float *__restrict__ logit_part, int32_t *err_flag) {
    constexpr int DP = NV * 32;
    constexpr int NE = NV * 4;
    constexpr unsigned kFull = 0xffffffffu;
    extern __shared__ __align__(16) float smem[];
    float *sw = smem;                    // [L][DP]
    float *sb = sw + ca.L * DP;          // [L][DP]
    float *swf = sb + ca.L * DP;         // [DP]
    const int TW = DP - mix0, TS = TW + 4;          // tile row: the mixed columns, +4 floats so that rows spread over the banks
    float *tiles = swf + DP;             // [warps][32][TS]
    __shared__ SmemSeg sseg[2 + DCNR_MAX_CAT];

    const int tid = threadIdx.x;
    for (int i = tid; i < ca.L * DP; i += kThreads) {
        int l = i / DP, c = i % DP;
        sw[i] = c < ca.D ? ca.w[l][c] : 0.f;
        sb[i] = c < ca.D ? ca.b[l][c] : 0.f;
    }
    for (int c = tid; c < DP; c += kThreads) swf[c] = (wf_cross != nullptr && c < ca.D) ? wf_cross[c] : 0.f;
    for (int i = tid; i < (kThreads / 32) * 32 * TS; i += kThreads) tiles[i] = 0.f;     // pad columns stay zero
    if (tid < ga.n_seg) {
        sseg[tid].table = ga.seg[tid].table;
        sseg[tid].ids = ga.seg[tid].ids;
        sseg[tid].rows = ga.seg[tid].rows;
        sseg[tid].id_stride = ga.seg[tid].id_stride;
        sseg[tid].width = ga.seg[tid].width;
        sseg[tid].col0 = ga.seg[tid].col0;
    }
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31, lane8 = lane & 7, grp = lane >> 3;
    float *tile = tiles + warp * 32 * TS;
    // row-invariant source of each of this lane's column quads: an aligned table segment (index, offset) or the tile (-1, column)
    int qseg[NV], qoff[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int col = 32 * j + 4 * lane8;
        qseg[j] = -1;
        qoff[j] = col - mix0;
        if (col < mix0) {
            for (int sgi = 0; sgi < n_vec; ++sgi)
                if (col >= sseg[sgi].col0 && col < sseg[sgi].col0 + sseg[sgi].width) {
                    qseg[j] = sgi;
                    qoff[j] = col - sseg[sgi].col0;
                }
        }
    }
    const bool x0_vec = x0_out != nullptr && ldx0 >= DP && (ldx0 & 3) == 0;
    const bool y_vec = y_out != nullptr && ldy >= DP && (ldy & 3) == 0;
    const int D = ga.D;
    const int n_num = ga.n_num, num_c0 = ga.num_col0 - mix0;
    const float inv_num = n_num > 0 ? 1.f / (float)n_num : 0.f;
    bool bad_id = false;

    const int64_t n_blocks = (B + 31) >> 5;
    const int64_t blk_step = (int64_t)gridDim.x * (kThreads / 32);
    constexpr int kMixPref = 2;          // ids of the first two narrow tables are fetched one step ahead as well
    // ids of one 32-row step: lane i loads the ids of row 32 * blk + i (coalesced), bounds-checked here
    auto load_ids = [&](int64_t blk, int (&v)[kMaxVecSeg], int (&c)[kMixPref]) {
        const int64_t r = (blk << 5) + lane;
        const bool have = blk < n_blocks && r < B;
#pragma unroll
        for (int sgi = 0; sgi < kMaxVecSeg; ++sgi) {
            v[sgi] = 0;
            if (sgi < n_vec && have) {
                const int64_t id = __ldcs(sseg[sgi].ids + r * sseg[sgi].id_stride);
                if ((uint64_t)id >= (uint64_t)sseg[sgi].rows) bad_id = true;
                else v[sgi] = (int)id;
            }
        }
#pragma unroll
        for (int m = 0; m < kMixPref; ++m) {
            c[m] = 0;
            const int sgi = n_vec + m;
            if (sgi < ga.n_seg && have) {
                const int64_t id = __ldcs(sseg[sgi].ids + r * sseg[sgi].id_stride);
                if ((uint64_t)id >= (uint64_t)sseg[sgi].rows) bad_id = true;
                else c[m] = (int)id;
            }
        }
    };
    int vid[kMaxVecSeg], cpre[kMixPref];
    const int64_t blk0 = (int64_t)blockIdx.x * (kThreads / 32) + warp;
    load_ids(blk0, vid, cpre);
    for (int64_t blk = blk0; blk < n_blocks; blk += blk_step) {
        const int64_t base = blk << 5;
        const int64_t myrow = base + lane;
        const bool have = myrow < B;
        // 2a. the remaining (narrow / unaligned) tables, element by element in flat order
        for (int sgi = n_vec; sgi < ga.n_seg; ++sgi) {
            const SmemSeg &sg = sseg[sgi];
            int cid = 0;
            if (sgi - n_vec < kMixPref) {
                cid = sgi == n_vec ? cpre[0] : cpre[1];
            } else if (have) {
                const int64_t id = __ldcs(sg.ids + myrow * sg.id_stride);
                if ((uint64_t)id >= (uint64_t)sg.rows) bad_id = true;
                else cid = (int)id;
            }
            const int w = sg.width, c0 = sg.col0 - mix0;
            const float inv_w = 1.f / (float)w;
            for (int t = 0; t < w; ++t) {
                const int idx = t * 32 + lane;
                const int r = (int)(((float)idx + 0.5f) * inv_w);       // exact: idx <= 32 * w, w <= 256
                const int c = idx - r * w;
                const int rid = __shfl_sync(kFull, cid, r);
                tile[r * TS + c0 + c] = __ldg(sg.table + (int64_t)rid * w + c);      // rows past B read row 0 (never stored)
            }
        }
        // 2b. the numerics: 32 x n_num contiguous floats
        if (n_num > 0) {
            const float *np = ga.num + base * n_num;
            const int64_t lim = (B - base) * n_num;
            for (int t = 0; t < n_num; ++t) {
                const int idx = t * 32 + lane;
                const int r = (int)(((float)idx + 0.5f) * inv_num);
                const int c = idx - r * n_num;
                tile[r * TS + num_c0 + c] = idx < lim ? __ldcs(np + idx) : 0.f;
            }
        }
        __syncwarp();
        // ids of the next step now, so that they are in registers when this step's rows are done
        int nvid[kMaxVecSeg], ncpre[kMixPref];
        load_ids(blk + blk_step, nvid, ncpre);
        // 3. four rows per pass, 8 lanes per row.  The table-row loads of PF passes are issued together (one exposed
        //    memory round trip per PF passes instead of one per pass).
#pragma unroll 1
        for (int it0 = 0; it0 < 8; it0 += PF) {
        float4 pre[PF][NV];
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int r = (it0 + u) * 4 + grp;
            int rid[kMaxVecSeg];
#pragma unroll
            for (int sgi = 0; sgi < kMaxVecSeg; ++sgi) rid[sgi] = sgi < n_vec ? __shfl_sync(kFull, vid[sgi], r) : 0;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                if (qseg[j] >= 0) {
                    const SmemSeg &sg = sseg[qseg[j]];
                    const int id = qseg[j] == 0 ? rid[0] : (qseg[j] == 1 ? rid[1] : (qseg[j] == 2 ? rid[2] : rid[3]));
                    pre[u][j] = ldg4(sg.table + (int64_t)id * sg.width + qoff[j]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < PF; ++u) {
            const int r = (it0 + u) * 4 + grp;
            const int64_t row = base + r;
            const bool active = row < B;
            float x[NE];
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const float4 v4 = qseg[j] >= 0 ? pre[u][j] : *reinterpret_cast<const float4 *>(tile + r * TS + qoff[j]);
                x[4 * j] = v4.x; x[4 * j + 1] = v4.y; x[4 * j + 2] = v4.z; x[4 * j + 3] = v4.w;
            }
            if (x0_out != nullptr && active) {
                if (x0_vec) {
#pragma unroll
                    for (int j = 0; j < NV; ++j)
                        __stcs(reinterpret_cast<float4 *>(x0_out + row * ldx0 + 32 * j + 4 * lane8), make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]));
                } else {
#pragma unroll
                    for (int k = 0; k < NE; ++k) {
                        int col = 32 * (k >> 2) + 4 * lane8 + (k & 3);
                        if (col < ldx0 && col < DP) x0_out[row * ldx0 + col] = x[k];
                    }
                }
            }
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
                const float sdot = group8_sum(p);
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    float4 b4 = *reinterpret_cast<const float4 *>(bl + 32 * j + 4 * lane8);
                    x[4 * j + 0] = fmaf(x[4 * j + 0], sdot, x[4 * j + 0]) + b4.x;
                    x[4 * j + 1] = fmaf(x[4 * j + 1], sdot, x[4 * j + 1]) + b4.y;
                    x[4 * j + 2] = fmaf(x[4 * j + 2], sdot, x[4 * j + 2]) + b4.z;
                    x[4 * j + 3] = fmaf(x[4 * j + 3], sdot, x[4 * j + 3]) + b4.w;
                }
            }
            if (y_out != nullptr && active) {
                if (y_vec) {
#pragma unroll
                    for (int j = 0; j < NV; ++j)
                        __stcs(reinterpret_cast<float4 *>(y_out + row * ldy + 32 * j + 4 * lane8), make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]));
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
                const float sdot = group8_sum(p);
                if (active && lane8 == 0) logit_part[row] = sdot;
            }
        }
        }
        __syncwarp();       // the tile is rewritten by the next step
#pragma unroll
        for (int sgi = 0; sgi < kMaxVecSeg; ++sgi) vid[sgi] = nvid[sgi];
#pragma unroll
        for (int m = 0; m < kMixPref; ++m) cpre[m] = ncpre[m];
    }
    if (bad_id && err_flag != nullptr) atomicOr(err_flag, 1);
}


// ------------------------------------------------------------------------------------------------
// cp.async-pipelined form of the staged gather (narrow rows; the default when it fits).  The staged
// kernel above is latency-bound (ncu: long-scoreboard stalls, 4 CTAs / SM at 64 registers, and every
// attempt to keep more row loads in flight per thread cost occupancy).  Here the loads need no
// registers: a warp owns a ring of S shared-memory tiles of 32 FULL padded rows, and for step k + 1
// it issues every byte of the 32 rows as cp.async (16-byte chunks for the aligned table rows, 4-byte
// elements for narrow tables and the numerics, zero-fill past the batch) while step k is read back
// from its tile, pushed through the register-resident cross layers and stored.  Ids are fetched one
// step further ahead (plain loads, needed to form the cp.async addresses).
// ------------------------------------------------------------------------------------------------
constexpr int kPipeWarps = 4;          // default warps per CTA (W below)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src, uint32_t src_bytes) {      // src_bytes 0 -> zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NV, int S, int W, int RS>      // RS rows per step (32 or 16: half the tile, twice the warps per SM)
__global__ void __launch_bounds__(32 * W)
k_embed_cross_fwd_pipe(GatherArgs ga, int n_vec, int64_t B, CrossArgs ca, float *__restrict__ x0_out, int64_t ldx0,
                       float *__restrict__ y_out, int64_t ldy, const float *__restrict__ wf_cross,
                       float *__restrict__ logit_part, int32_t *err_flag) {
    constexpr int DP = NV * 32;
    constexpr int NE = NV * 4;
    constexpr int TS = DP + 4;           // tile row pitch in floats (+4: consecutive rows start 4 banks apart)
    constexpr unsigned kFull = 0xffffffffu;
    constexpr int kT = 32 * W;
    constexpr int kMixPref = 2;
    extern __shared__ __align__(16) float smem[];
    float *sw = smem;                    // [L][DP]
    float *sb = sw + ca.L * DP;          // [L][DP]
    float *swf = sb + ca.L * DP;         // [DP]
    float *tiles = swf + DP;             // [warps][S][RS][TS]
    __shared__ SmemSeg sseg[2 + DCNR_MAX_CAT];

    const int tid = threadIdx.x;
    for (int i = tid; i < ca.L * DP; i += kT) {
        int l = i / DP, c = i % DP;
        sw[i] = c < ca.D ? ca.w[l][c] : 0.f;
        sb[i] = c < ca.D ? ca.b[l][c] : 0.f;
    }
    for (int c = tid; c < DP; c += kT) swf[c] = (wf_cross != nullptr && c < ca.D) ? wf_cross[c] : 0.f;
    for (int i = tid; i < W * S * RS * TS; i += kT) tiles[i] = 0.f;          // pad columns stay zero
    if (tid < ga.n_seg) {
        sseg[tid].table = ga.seg[tid].table;
        sseg[tid].ids = ga.seg[tid].ids;
        sseg[tid].rows = ga.seg[tid].rows;
        sseg[tid].id_stride = ga.seg[tid].id_stride;
        sseg[tid].width = ga.seg[tid].width;
        sseg[tid].col0 = ga.seg[tid].col0;
    }
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31, lane8 = lane & 7, grp = lane >> 3;
    float *wtile = tiles + (size_t)warp * S * RS * TS;
    const uint32_t wtile_s = (uint32_t)__cvta_generic_to_shared(wtile);
    const bool x0_vec = x0_out != nullptr && ldx0 >= DP && (ldx0 & 3) == 0;
    const bool y_vec = y_out != nullptr && ldy >= DP && (ldy & 3) == 0;
    const int D = ga.D;
    const int n_num = ga.n_num, num_c0 = ga.num_col0;
    const float inv_num = n_num > 0 ? 1.f / (float)n_num : 0.f;
    bool bad_id = false;

    const int64_t n_blocks = (B + RS - 1) / RS;
    const int64_t blk_step = (int64_t)gridDim.x * W;
    auto load_ids = [&](int64_t blk, int (&v)[kMaxVecSeg], int (&c)[kMixPref]) {
        const int64_t r = blk * RS + lane;
        const bool have = blk < n_blocks && r < B && lane < RS;
#pragma unroll
        for (int sgi = 0; sgi < kMaxVecSeg; ++sgi) {
            v[sgi] = 0;
            if (sgi < n_vec && have) {
                const int64_t id = __ldcs(sseg[sgi].ids + r * sseg[sgi].id_stride);
                if ((uint64_t)id >= (uint64_t)sseg[sgi].rows) bad_id = true;
                else v[sgi] = (int)id;
            }
        }
#pragma unroll
        for (int m = 0; m < kMixPref; ++m) {
            c[m] = 0;
            const int sgi = n_vec + m;
            if (sgi < ga.n_seg && have) {
                const int64_t id = __ldcs(sseg[sgi].ids + r * sseg[sgi].id_stride);
                if ((uint64_t)id >= (uint64_t)sseg[sgi].rows) bad_id = true;
                else c[m] = (int)id;
            }
        }
    };
    // every byte of the 32 rows of step `blk` into ring slot `slot`, as cp.async (no registers held while in flight)
    auto issue = [&](int64_t blk, int slot, const int (&v)[kMaxVecSeg], const int (&c)[kMixPref]) {
        if (blk >= n_blocks) return;
        const int64_t base = blk * RS;
        const uint32_t t0 = wtile_s + (uint32_t)(slot * RS * TS * 4);
#pragma unroll
        for (int sgi = 0; sgi < kMaxVecSeg; ++sgi) {
            if (sgi < n_vec) {                               // aligned table rows: width / 4 chunks of 16 bytes per row
                const SmemSeg &sg = sseg[sgi];
                const int cpr = sg.width >> 2;
                const float inv_c = 1.f / (float)cpr;
                for (int t = 0; t * 32 < RS * cpr; ++t) {
                    const int idx = t * 32 + lane;
                    const int r = min((int)(((float)idx + 0.5f) * inv_c), RS - 1);
                    const int part = idx - r * cpr;
                    const int rid = __shfl_sync(kFull, v[sgi], r);
                    if (idx < RS * cpr)
                        cp_async16(t0 + (uint32_t)((r * TS + sg.col0 + 4 * part) * 4), sg.table + (int64_t)rid * sg.width + 4 * part);
                }
            }
        }
        for (int sgi = n_vec; sgi < ga.n_seg; ++sgi) {      // narrow / unaligned tables, element by element in flat order
            const SmemSeg &sg = sseg[sgi];
            int cid = 0;
            if (sgi - n_vec < kMixPref) {
                cid = sgi == n_vec ? c[0] : c[1];
            } else if (base + lane < B && lane < RS) {
                const int64_t id = __ldcs(sg.ids + (base + lane) * sg.id_stride);
                if ((uint64_t)id >= (uint64_t)sg.rows) bad_id = true;
                else cid = (int)id;
            }
            const int w = sg.width;
            const float inv_w = 1.f / (float)w;
            for (int t = 0; t * 32 < RS * w; ++t) {
                const int idx = t * 32 + lane;
                const int r = min((int)(((float)idx + 0.5f) * inv_w), RS - 1);       // exact: idx <= 32 * w, w <= 256
                const int cc = idx - r * w;
                const int rid = __shfl_sync(kFull, cid, r);
                if (idx < RS * w) cp_async4(t0 + (uint32_t)((r * TS + sg.col0 + cc) * 4), sg.table + (int64_t)rid * w + cc, 4u);
            }
        }
        if (n_num > 0) {                                     // numerics: 32 x n_num contiguous floats, zero past the batch
            const float *np = ga.num + base * n_num;
            const int64_t lim = (B - base) * n_num;
            for (int t = 0; t * 32 < RS * n_num; ++t) {
                const int idx = t * 32 + lane;
                const int r = (int)(((float)idx + 0.5f) * inv_num);
                const int cc = idx - r * n_num;
                const bool in = idx < lim;
                if (idx < RS * n_num) cp_async4(t0 + (uint32_t)((r * TS + num_c0 + cc) * 4), in ? np + idx : ga.num, in ? 4u : 0u);
            }
        }
    };

    const int64_t blk0 = (int64_t)blockIdx.x * W + warp;
    int vid[kMaxVecSeg], cpre[kMixPref];
    // prologue: S - 1 steps in flight
    load_ids(blk0, vid, cpre);
#pragma unroll
    for (int p = 0; p < S - 1; ++p) {
        issue(blk0 + p * blk_step, p, vid, cpre);
        cp_async_commit();
        load_ids(blk0 + (p + 1) * blk_step, vid, cpre);
    }
    int k = 0;
    for (int64_t blk = blk0; blk < n_blocks; blk += blk_step, ++k) {
        // keep S - 1 steps ahead in flight: step k + S - 1 goes into the slot step k - 1 just released
        issue(blk + (S - 1) * blk_step, (k + S - 1) % S, vid, cpre);
        cp_async_commit();
        load_ids(blk + (int64_t)S * blk_step, vid, cpre);        // its latency hides behind the compute below
        cp_async_wait<S - 1>();                                  // step k has landed (this thread's copies) ...
        __syncwarp();                                            // ... and every other lane's
        const float *tile = wtile + (size_t)(k % S) * RS * TS;
        const int64_t base = blk * RS;
#pragma unroll 2
        for (int it = 0; it < RS / 4; ++it) {
            const int r = it * 4 + grp;
            const int64_t row = base + r;
            const bool active = row < B;
            float x[NE];
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const float4 v4 = *reinterpret_cast<const float4 *>(tile + r * TS + 32 * j + 4 * lane8);
                x[4 * j] = v4.x; x[4 * j + 1] = v4.y; x[4 * j + 2] = v4.z; x[4 * j + 3] = v4.w;
            }
            if (x0_out != nullptr && active) {
                if (x0_vec) {
#pragma unroll
                    for (int j = 0; j < NV; ++j)
                        __stcs(reinterpret_cast<float4 *>(x0_out + row * ldx0 + 32 * j + 4 * lane8), make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]));
                } else {
#pragma unroll
                    for (int kk = 0; kk < NE; ++kk) {
                        int col = 32 * (kk >> 2) + 4 * lane8 + (kk & 3);
                        if (col < ldx0 && col < DP) x0_out[row * ldx0 + col] = x[kk];
                    }
                }
            }
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
                const float sdot = group8_sum(p);
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    float4 b4 = *reinterpret_cast<const float4 *>(bl + 32 * j + 4 * lane8);
                    x[4 * j + 0] = fmaf(x[4 * j + 0], sdot, x[4 * j + 0]) + b4.x;
                    x[4 * j + 1] = fmaf(x[4 * j + 1], sdot, x[4 * j + 1]) + b4.y;
                    x[4 * j + 2] = fmaf(x[4 * j + 2], sdot, x[4 * j + 2]) + b4.z;
                    x[4 * j + 3] = fmaf(x[4 * j + 3], sdot, x[4 * j + 3]) + b4.w;
                }
            }
            if (y_out != nullptr && active) {
                if (y_vec) {
#pragma unroll
                    for (int j = 0; j < NV; ++j)
                        __stcs(reinterpret_cast<float4 *>(y_out + row * ldy + 32 * j + 4 * lane8), make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]));
                } else {
#pragma unroll
                    for (int kk = 0; kk < NE; ++kk) {
                        int col = 32 * (kk >> 2) + 4 * lane8 + (kk & 3);
                        if (col < D && col < ldy) y_out[row * ldy + col] = x[kk];
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
                const float sdot = group8_sum(p);
                if (active && lane8 == 0) logit_part[row] = sdot;
            }
        }
        __syncwarp();       // the slot is refilled by the next iteration's issue()
    }
    cp_async_wait<0>();
    if (bad_id && err_flag != nullptr) atomicOr(err_flag, 1);
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
    // Three forms, one per shape class (each measured the fastest for its class in round 1, profiles/r01_ncu_k1_*.md):
    //   * narrow rows (in_dim_pad <= 64, the reference's default shape): k_embed_cross_fwd_pipe -- cp.async ring of two 32-row
    //     tiles per warp, four warps per CTA;
    //   * wider rows with a prefix of 16-byte aligned segments and a mixed tail of <= 96 floats: k_embed_cross_fwd_staged;
    //   * everything else (cross-only calls on a given x, odd layouts): the element-wise k_embed_cross_fwd.
    if (x_in == nullptr && ga->n_seg > 0) {
        int n_vec = 0, mix0 = 0;
        bool ok = true;
        for (int i = 0; i < ga->n_seg; ++i) {
            const GatherSeg &sg = ga->seg[i];
            ok = ok && sg.table != nullptr && sg.rows > 0 && sg.rows < (1ll << 31) && sg.width >= 1 && sg.width <= 256;
            if (i == n_vec && n_vec < kMaxVecSeg && sg.col0 == mix0 && (sg.width & 3) == 0 && (sg.col0 & 3) == 0 &&
                (reinterpret_cast<uintptr_t>(sg.table) & 15) == 0) {
                ++n_vec;
                mix0 = sg.col0 + sg.width;
            }
        }
        const int tw = dim_pad - mix0;
        ok = ok && tw >= 0 && tw <= 96 && ga->n_num <= 256 && (ga->n_num == 0 || ga->num != nullptr);
        if (ok && nv <= 2) {
            // Measured per 4 M rows (P0; box-to-box spread is ~10 %): 32-row steps with 2 / 4 / 8 warps per CTA 405-470 / 414 /
            // 510 us; 16-row steps (twice the warps per SM) 428 / 420 / 437 us; a ring of three tiles 516 us.
            constexpr int S = 2, W = kPipeWarps, RS = 32;
            const size_t smem_p = smem + (size_t)W * S * RS * (dim_pad + 4) * sizeof(float);
            const unsigned grid_p = (unsigned)std::min<int64_t>(ceil_div(B, RS * W), (int64_t)sm_count() * 64 / W);
#define DCNR_LAUNCH_P(NVV)                                                                                                 \
    do {                                                                                                                   \
        DCNR_CUDA_CHECK(cudaFuncSetAttribute(k_embed_cross_fwd_pipe<NVV, S, W, RS>,                                        \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_p));                   \
        k_embed_cross_fwd_pipe<NVV, S, W, RS><<<grid_p, 32 * W, smem_p, stream>>>(*ga, n_vec, B, ca, x0_out, ldx0, y_out,   \
                                                                                  ldy, wf_cross, logit_part, err_flag);    \
    } while (0)
            if (nv == 1) DCNR_LAUNCH_P(1);
            else DCNR_LAUNCH_P(2);
#undef DCNR_LAUNCH_P
            DCNR_LAUNCHED();
            return DCNR_OK;
        }
        if (ok) {
            const size_t smem_s = smem + (size_t)(kThreads / 32) * 32 * (tw + 4) * sizeof(float);
            const unsigned grid_s = (unsigned)std::min<int64_t>(ceil_div(B, 32 * (kThreads / 32)), (int64_t)sm_count() * 8);
            // Measured at 4 M rows (P0, us): one pass of row loads in flight at 64 registers (4 CTAs / SM) 463; 2 passes 500;
            // 4 passes (130 registers) 548; 48 / 40 registers for 5 / 6 CTAs per SM 578 / 754 (spills).
#define DCNR_CASE_S(NVV, PFF)                                                                                              \
    case NVV:                                                                                                              \
        DCNR_CUDA_CHECK(cudaFuncSetAttribute(k_embed_cross_fwd_staged<NVV, PFF, 1>,                                        \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));                   \
        k_embed_cross_fwd_staged<NVV, PFF, 1><<<grid_s, kThreads, smem_s, stream>>>(*ga, n_vec, mix0, B, ca, x0_out, ldx0, \
                                                                                  y_out, ldy, wf_cross, logit_part,        \
                                                                                  err_flag);                               \
        break;
            switch (nv) {
                DCNR_CASE_S(3, 2) DCNR_CASE_S(4, 2) DCNR_CASE_S(5, 1) DCNR_CASE_S(6, 1) DCNR_CASE_S(7, 1) DCNR_CASE_S(8, 1)
            }
#undef DCNR_CASE_S
            DCNR_LAUNCHED();
            return DCNR_OK;
        }
    }
    // element-wise form, one row per 8-lane group per iteration (measured at 4 M rows, P0: 1 row 513 us, 2 rows 647 us, 4 rows
    // 781 us -- the kernel is issue-bound)
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(B, kGroupsPerCta), (int64_t)sm_count() * 8);
#define DCNR_CASE(NVV)                                                                                                   \
    case NVV:                                                                                                            \
        k_embed_cross_fwd<NVV, 1><<<grid, kThreads, smem, stream>>>(*ga, x_in, ldx_in, B, ca, x0_out, ldx0, y_out, ldy,  \
                                                                    wf_cross, logit_part, err_flag);                     \
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
