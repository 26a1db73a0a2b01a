// K7: deterministic sorted-segment scatter-add of embedding gradients.
//
// Replaces the implicit embedding_dense_backward of loss.backward() (train.py:225) for the tables
// of train.py:136-139.  The gradient of table row r is the sum of dx0[b, col0:col0+width] over the
// batch rows b with id_b == r, added in ascending b (the order the reference's CPU kernel uses):
//   1. (id, b) pairs are radix-sorted by id (stable, so b stays ascending inside a segment);
//   2. fixed windows of 32 sorted positions are summed sequentially, one thread per
//      (window, column); segments that live inside one window are written straight to the dense
//      gradient, window-crossing segments leave a head/tail partial;
//   3. a fix-up pass walks each window-crossing chain in window order.
// No atomics, so the result is bit-reproducible.  The radix sort itself is CUB (library plumbing,
// see DESIGN.md); steps 2-3 are the kernels below.

#include "kernels.cuh"

namespace dcnr {

constexpr int kWin = 32;

// Two tables of the same width can share ONE sort: positions B..2B-1 carry table 1's ids with bit `tbit` set, so the
// sorted order is table 0's segments, then table 1's (tbit = 32: single table).
struct ScatterTables {
    float *grad[2];
    int32_t col0[2];
    int32_t tbit;          // key bit that selects the table (32 = one table)
    int64_t B;             // batch rows per table
};

__global__ void k_scatter_prep(const int64_t *__restrict__ ids0, int64_t stride0, int64_t rows0,
                               const int64_t *__restrict__ ids1, int64_t stride1, int64_t rows1, int64_t B, int tbit,
                               uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (ids1 != nullptr ? 2 * B : B)) return;
    const bool second = i >= B;
    int64_t id = second ? ids1[(i - B) * stride1] : ids0[i * stride0];
    const int64_t n_rows = second ? rows1 : rows0;
    if (id < 0) id = 0;
    if (id >= n_rows) id = n_rows - 1;     // out-of-range ids are reported by dcnr_check_ids; stay memory-safe here
    keys[i] = (uint32_t)id | (second ? (1u << tbit) : 0u);
    vals[i] = (uint32_t)i;
}

__device__ __forceinline__ float *scatter_dst(const ScatterTables &t, uint32_t key, int width, int col) {
    const uint32_t tb = t.tbit < 32 ? (key >> t.tbit) : 0u;
    const uint32_t id = t.tbit < 32 ? (key & ((1u << t.tbit) - 1u)) : key;
    return t.grad[tb] + (int64_t)id * width + col;
}
__device__ __forceinline__ float scatter_src(const ScatterTables &t, uint32_t key, uint32_t val, const float *dx0, int64_t lddx,
                                             int col) {
    const uint32_t tb = t.tbit < 32 ? (key >> t.tbit) : 0u;
    return __ldg(dx0 + ((int64_t)val - (int64_t)tb * t.B) * lddx + t.col0[tb] + col);
}

// flags[w]: bit0 = first segment continues from window w-1 (head partial in carry[w][0])
//           bit1 = that head segment also continues into window w+1
//           bit2 = last segment starts here and continues into window w+1 (tail partial in carry[w][1])
// One WARP per window of 32 sorted positions: the lanes load the window's keys / vals once (coalesced) and broadcast
// them by shuffle; lane c then owns gradient column c (lanes >= width idle; widths > 32 loop) and walks the 32 positions
// in order with the 8 source rows of each step already in flight (the one-thread-per-column version re-read keys and
// vals from memory and had one dependent load chain per position: 67 us for the user table at B = 1 M).
__global__ void __launch_bounds__(128)
k_scatter_window(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, int64_t B, int width,
                 const float *__restrict__ dx0, int64_t lddx, ScatterTables tabs,
                 float *__restrict__ carry, uint8_t *__restrict__ flags, int64_t n_windows) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= n_windows) return;                                   // warp-uniform
    const int64_t p0 = w * kWin, p1 = min(B, p0 + kWin);
    const int cnt = (int)(p1 - p0);
    const bool has_left = p0 > 0, has_right = p1 < B;
    const uint32_t my_key = p0 + lane < p1 ? keys[p0 + lane] : 0u;
    const uint32_t my_val = p0 + lane < p1 ? vals[p0 + lane] : 0u;
    uint32_t edge = 0u;
    if (lane == 0 && has_left) edge = keys[p0 - 1];
    if (lane == 1 && has_right) edge = keys[p1];
    const uint32_t left_id = __shfl_sync(0xffffffffu, edge, 0), right_id = __shfl_sync(0xffffffffu, edge, 1);
    const uint32_t first_key = __shfl_sync(0xffffffffu, my_key, 0);
    for (int c0 = 0; c0 < width; c0 += 32) {
        const int col = c0 + lane;
        const bool on = col < width;
        uint32_t cur = first_key;
        float acc = 0.f;
        bool first = true;
        uint8_t fl = 0;
        for (int q0 = 0; q0 < cnt; q0 += 8) {
            uint32_t id[8];
            float x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int pos = min(q0 + j, cnt - 1);
                id[j] = __shfl_sync(0xffffffffu, my_key, pos);
                const uint32_t v = __shfl_sync(0xffffffffu, my_val, pos);
                x[j] = on ? scatter_src(tabs, id[j], v, dx0, lddx, col) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (q0 + j < cnt) {
                    if (id[j] != cur) {
                        if (first && has_left && cur == left_id) {
                            if (on) carry[(w * 2 + 0) * width + col] = acc;
                            fl |= 1;
                        } else if (on) {
                            *scatter_dst(tabs, cur, width, col) = acc;
                        }
                        cur = id[j];
                        acc = 0.f;
                        first = false;
                    }
                    acc += x[j];
                }
            }
        }
        const bool left_open = first && has_left && cur == left_id;
        const bool right_open = has_right && cur == right_id;
        if (left_open) {
            if (on) carry[(w * 2 + 0) * width + col] = acc;
            fl |= 1;
            if (right_open) fl |= 2;
        } else if (right_open) {
            if (on) carry[(w * 2 + 1) * width + col] = acc;
            fl |= 4;
        } else if (on) {
            *scatter_dst(tabs, cur, width, col) = acc;
        }
        if (col == 0) flags[w] = fl;
    }
}

// Widths <= 16 (the user / item tables of the P0 model): two windows per warp, lanes 0-15 and 16-31, so that every lane
// carries a gradient column and a warp keeps 16 source rows in flight instead of 8 (the one-window form leaves half the
// lanes idle).  A lane of a half holds two of its window's 32 sorted (key, val) pairs; everything else is the walk above.
__global__ void __launch_bounds__(128)
k_scatter_window16(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, int64_t B, int width,
                   const float *__restrict__ dx0, int64_t lddx, ScatterTables tabs,
                   float *__restrict__ carry, uint8_t *__restrict__ flags, int64_t n_windows) {
    constexpr unsigned kFull = 0xffffffffu;
    const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
    const int64_t wpair = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wpair * 2 >= n_windows) return;                           // warp-uniform
    const int64_t w = wpair * 2 + half;
    const bool wvalid = w < n_windows;
    const int64_t p0 = w * kWin, p1 = wvalid ? min(B, p0 + kWin) : p0;
    const int cnt = (int)(p1 - p0);
    const bool has_left = wvalid && p0 > 0, has_right = wvalid && p1 < B;
    const uint32_t k_lo = p0 + hl < p1 ? keys[p0 + hl] : 0u, k_hi = p0 + hl + 16 < p1 ? keys[p0 + hl + 16] : 0u;
    const uint32_t v_lo = p0 + hl < p1 ? vals[p0 + hl] : 0u, v_hi = p0 + hl + 16 < p1 ? vals[p0 + hl + 16] : 0u;
    uint32_t edge = 0u;
    if (hl == 0 && has_left) edge = keys[p0 - 1];
    if (hl == 1 && has_right) edge = keys[p1];
    const uint32_t left_id = __shfl_sync(kFull, edge, 0, 16), right_id = __shfl_sync(kFull, edge, 1, 16);
    const uint32_t first_key = __shfl_sync(kFull, k_lo, 0, 16);
    const int col = hl;
    const bool on = wvalid && col < width;
    uint32_t cur = first_key;
    float acc = 0.f;
    bool first = true;
    uint8_t fl = 0;
#pragma unroll 1
    for (int q0 = 0; q0 < kWin; q0 += 8) {                        // fixed trip count: both halves shuffle together
        uint32_t id[8];
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int pos = min(q0 + j, max(cnt - 1, 0));
            const uint32_t ka = __shfl_sync(kFull, k_lo, pos & 15, 16), kb = __shfl_sync(kFull, k_hi, pos & 15, 16);
            const uint32_t va = __shfl_sync(kFull, v_lo, pos & 15, 16), vb = __shfl_sync(kFull, v_hi, pos & 15, 16);
            id[j] = pos < 16 ? ka : kb;
            const uint32_t v = pos < 16 ? va : vb;
            x[j] = (on && q0 + j < cnt) ? scatter_src(tabs, id[j], v, dx0, lddx, col) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (q0 + j < cnt) {
                if (id[j] != cur) {
                    if (first && has_left && cur == left_id) {
                        if (on) carry[(w * 2 + 0) * width + col] = acc;
                        fl |= 1;
                    } else if (on) {
                        *scatter_dst(tabs, cur, width, col) = acc;
                    }
                    cur = id[j];
                    acc = 0.f;
                    first = false;
                }
                acc += x[j];
            }
        }
    }
    if (!wvalid) return;
    const bool left_open = first && has_left && cur == left_id;
    const bool right_open = has_right && cur == right_id;
    if (left_open) {
        if (on) carry[(w * 2 + 0) * width + col] = acc;
        fl |= 1;
        if (right_open) fl |= 2;
    } else if (right_open) {
        if (on) carry[(w * 2 + 1) * width + col] = acc;
        fl |= 4;
    } else if (on) {
        *scatter_dst(tabs, cur, width, col) = acc;
    }
    if (col == 0) flags[w] = fl;
}

__global__ void __launch_bounds__(128)
k_scatter_fixup(const uint32_t *__restrict__ keys, int64_t B, int width, ScatterTables tabs,
                const float *__restrict__ carry, const uint8_t *__restrict__ flags, int64_t n_windows) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_windows * width) return;
    const int64_t w = e / width;
    const int col = (int)(e % width);
    if (!(flags[w] & 4)) return;                       // only chain starts do work
    const int64_t p1 = min(B, (w + 1) * kWin);
    const uint32_t id = keys[p1 - 1];
    float acc = carry[(w * 2 + 1) * width + col];
    for (int64_t w2 = w + 1; w2 < n_windows; ++w2) {
        acc += carry[(w2 * 2 + 0) * width + col];
        if (!(flags[w2] & 2)) break;
    }
    *scatter_dst(tabs, id, width, col) = acc;
}

// ---- tiny tables (city: 100 rows, hotel_type: 6 rows => thousands of duplicates per row) -----------------------
// Sorting buys nothing here and the window-crossing chains of the sorted path become thousands of windows long
// (measured: 1.1 ms for the 6-row table at B = 1 M).  Instead every thread owns one COLUMN of a private copy of the
// table in shared memory and adds its sub-chunk of kSubRows batch rows in ascending order; the CTA then folds its
// sub-chunk tables in order and writes one partial table; partial tables are summed in CTA order in double.
// No atomics, fixed association => bit-reproducible.
constexpr int kSubRows = 64;
constexpr int kSmallThreads = 128;
constexpr int kSmallSmemFloats = 20 * 1024;     // 80 KB of private tables per CTA at most
constexpr int kSmallMaxTable = 2048;            // floats in one table copy (city 100 x 11, hotel_type 6 x 3)

__global__ void __launch_bounds__(kSmallThreads)
k_scatter_small(const int64_t *__restrict__ ids, int64_t id_stride, int64_t B, int n_rows, int width,
                const float *__restrict__ dx0, int64_t lddx, int col0, int subs_per_cta, float *__restrict__ partials) {
    extern __shared__ __align__(16) float tab[];          // [subs_per_cta][n_rows][width]
    const int tsz = n_rows * width;
    for (int i = threadIdx.x; i < subs_per_cta * tsz; i += kSmallThreads) tab[i] = 0.f;
    __syncthreads();
    const int s = threadIdx.x / width, j = threadIdx.x % width;
    if (s < subs_per_cta) {
        const int64_t b0 = ((int64_t)blockIdx.x * subs_per_cta + s) * kSubRows, b1 = min(B, b0 + kSubRows);
        float *mine = tab + (size_t)s * tsz + j;
        for (int64_t bb = b0; bb < b1; bb += 8) {                    // 8 (id, value) pairs in flight, added in batch order
            int64_t id[8];
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int64_t b = min(bb + q, b1 - 1);
                id[q] = __ldg(ids + b * id_stride);
                v[q] = __ldg(dx0 + b * lddx + col0 + j);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (bb + q < b1) {
                    const int64_t r = id[q] < 0 ? 0 : (id[q] >= n_rows ? n_rows - 1 : id[q]);   // dcnr_check_ids reports bad ids
                    mine[r * width] += v[q];
                }
            }
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < tsz; e += kSmallThreads) {
        float acc = 0.f;
        for (int q = 0; q < subs_per_cta; ++q) acc += tab[(size_t)q * tsz + e];
        partials[(int64_t)blockIdx.x * tsz + e] = acc;
    }
}

static int small_subs_per_cta(int64_t n_rows, int32_t width) {
    if (width > 32 || n_rows * width > kSmallMaxTable) return 0;       // >= 4 sub-chunks per CTA => <= 32 B of partials per row
    return (int)std::min<int64_t>(kSmallThreads / width, kSmallSmemFloats / (n_rows * width));
}
static int64_t small_ctas(int64_t B, int subs) { return ceil_div(ceil_div(std::max<int64_t>(B, 1), kSubRows), subs); }

// All tiny tables of a model in ONE pass over the batch (the P0 model has two: city and hotel_type, whose dx0 columns sit
// in the same 64 bytes of every row): thread = (sub-chunk, column of one of the tables).  Same private-table scheme as
// k_scatter_small; one partial block [sum of table sizes] per CTA.
struct SmallTables {
    const int64_t *ids[DCNR_MAX_CAT];
    int32_t id_stride[DCNR_MAX_CAT], n_rows[DCNR_MAX_CAT], width[DCNR_MAX_CAT], col0[DCNR_MAX_CAT];
    int32_t tab_off[DCNR_MAX_CAT], col_begin[DCNR_MAX_CAT + 1];
    int32_t n, tsz_total, w_total;
};

__global__ void __launch_bounds__(kSmallThreads)
k_scatter_small_multi(SmallTables st, int64_t B, const float *__restrict__ dx0, int64_t lddx, int subs_per_cta,
                      float *__restrict__ partials) {
    extern __shared__ __align__(16) float tab[];          // [subs_per_cta][tsz_total]
    const int tsz = st.tsz_total;
    for (int i = threadIdx.x; i < subs_per_cta * tsz; i += kSmallThreads) tab[i] = 0.f;
    __syncthreads();
    const int s = threadIdx.x / st.w_total, jj = threadIdx.x % st.w_total;
    if (s < subs_per_cta) {
        int t = 0;
        while (t + 1 < st.n && jj >= st.col_begin[t + 1]) ++t;
        const int j = jj - st.col_begin[t], width = st.width[t], n_rows = st.n_rows[t];
        const int64_t *ids = st.ids[t];
        const int64_t id_stride = st.id_stride[t];
        const int col = st.col0[t] + j;
        const int64_t b0 = ((int64_t)blockIdx.x * subs_per_cta + s) * kSubRows, b1 = min(B, b0 + kSubRows);
        float *mine = tab + (size_t)s * tsz + st.tab_off[t] + j;
        for (int64_t bb = b0; bb < b1; bb += 8) {                    // 8 (id, value) pairs in flight, added in batch order
            int64_t id[8];
            float v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int64_t b = min(bb + q, b1 - 1);
                id[q] = __ldg(ids + b * id_stride);
                v[q] = __ldg(dx0 + b * lddx + col);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (bb + q < b1) {
                    const int64_t r = id[q] < 0 ? 0 : (id[q] >= n_rows ? n_rows - 1 : id[q]);   // dcnr_check_ids reports bad ids
                    mine[r * width] += v[q];
                }
            }
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < tsz; e += kSmallThreads) {
        float acc = 0.f;
        for (int q = 0; q < subs_per_cta; ++q) acc += tab[(size_t)q * tsz + e];
        partials[(int64_t)blockIdx.x * tsz + e] = acc;
    }
}

// true (and launched) when every categorical table with a gradient takes the tiny-table path and they fit one pass
static bool try_scatter_small_multi(const dcnr_dims *dims, const dcnr_batch *batch, const float *dx0, int64_t lddx,
                                    const dcnr_grads *grads, void *scratch, int64_t scratch_bytes, cudaStream_t stream, int *rc) {
    *rc = DCNR_OK;
    const int64_t B = batch->batch;
    if (B <= 0) return false;
    SmallTables st;
    memset(&st, 0, sizeof(st));
    int col = 2 * dims->emb_dim, n = 0;
    float *out[DCNR_MAX_CAT];
    for (int i = 0; i < dims->n_cat; ++i) {
        if (grads->cat_table[i]) {
            if (small_subs_per_cta(dims->cat_rows[i], dims->cat_width[i]) == 0) return false;
            st.ids[n] = batch->cat_features + i;
            st.id_stride[n] = dims->n_cat;
            st.n_rows[n] = (int32_t)dims->cat_rows[i];
            st.width[n] = dims->cat_width[i];
            st.col0[n] = col;
            st.tab_off[n] = st.tsz_total;
            st.col_begin[n] = st.w_total;
            st.tsz_total += st.n_rows[n] * st.width[n];
            st.w_total += st.width[n];
            out[n] = grads->cat_table[i];
            ++n;
        }
        col += dims->cat_width[i];
    }
    st.n = n;
    st.col_begin[n] = st.w_total;
    if (n < 2 || st.w_total > 64 || st.tsz_total > kSmallMaxTable) return false;
    const int subs = std::min(kSmallThreads / st.w_total, kSmallSmemFloats / st.tsz_total);
    if (subs < 2) return false;
    const int64_t ctas = small_ctas(B, subs);
    if (scratch_bytes < ctas * st.tsz_total * 4) return false;
    float *partials = reinterpret_cast<float *>(scratch);
    const size_t smem = (size_t)subs * st.tsz_total * sizeof(float);
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(k_scatter_small_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    k_scatter_small_multi<<<(unsigned)ctas, kSmallThreads, smem, stream>>>(st, B, dx0, lddx, subs, partials);
    DCNR_LAUNCHED();
    for (int t = 0; t < n && *rc == DCNR_OK; ++t)
        *rc = launch_sum_partials_2d(partials + st.tab_off[t], ctas, st.n_rows[t], st.width[t], st.width[t], out[t], st.width[t],
                                     stream, st.tsz_total);
    return true;
}

// rows b >= B (up to cap) are padding: id 0 with an all-zero gradient row, which adds nothing to row 0's sum
__global__ void k_pack_embed_grads(const int64_t *__restrict__ user_ids, const int64_t *__restrict__ item_ids,
                                   const float *__restrict__ dx0, int64_t lddx, int64_t B, int64_t cap, int w2,
                                   int64_t *__restrict__ ids_out, float *__restrict__ rows_out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= cap * w2) return;
    const int64_t b = e / w2;
    const int c = (int)(e % w2);
    const bool real = b < B;
    rows_out[e] = real ? __ldg(dx0 + b * lddx + c) : 0.f;
    if (c == 0) ids_out[2 * b] = real ? user_ids[b] : 0;
    if (c == 1) ids_out[2 * b + 1] = real ? item_ids[b] : 0;
}

int launch_pack_embed_grads(const int64_t *user_ids, const int64_t *item_ids, const float *dx0, int64_t lddx, int64_t B,
                            int64_t cap, int32_t emb_dim, int64_t *ids_out, float *rows_out, cudaStream_t stream) {
    if (cap <= 0) return DCNR_OK;
    const int w2 = 2 * emb_dim;
    k_pack_embed_grads<<<(unsigned)ceil_div(cap * w2, 256), 256, 0, stream>>>(user_ids, item_ids, dx0, lddx, B, cap, w2, ids_out,
                                                                             rows_out);
    DCNR_LAUNCHED();
    return DCNR_OK;
}

static int sort_bits(int64_t n_rows) {
    int b = 1;
    while (((int64_t)1 << b) < n_rows && b < 32) ++b;
    return b;
}

int64_t scatter_scratch_bytes(int64_t B) {
    if (B <= 0) return 256;
    const int64_t n_windows = ceil_div(B, kWin);
    int64_t bytes = 4 * round_up(B * 4, 256);                       // keys/vals double buffers
    bytes += round_up(n_windows * 2 * 256 * 4, 256);                // carry, width <= 256
    bytes += round_up(n_windows, 256);                              // flags
    bytes += radix_sort_scratch_bytes(B);
    return bytes;      // (the tiny-table path's partial tables need <= 32 B per batch row: they reuse this space)
}

// Sorted-segment path for one table, or for two tables of equal width sharing one radix sort (ids1 != NULL).
static int launch_scatter_sorted(const int64_t *ids0, int64_t stride0, int64_t rows0, float *grad0, int32_t col0,
                                 const int64_t *ids1, int64_t stride1, int64_t rows1, float *grad1, int32_t col1, int64_t B,
                                 int32_t width, const float *dx0, int64_t lddx, void *scratch, int64_t scratch_bytes,
                                 cudaStream_t stream) {
    const bool two = ids1 != nullptr;
    const int64_t n = two ? 2 * B : B;
    DCNR_REQUIRE(n < 0x7fffffffLL, "batch too large for one scatter");
    if (scratch_bytes < scatter_scratch_bytes(n)) {
        set_error("scatter scratch too small (%lld < %lld)", (long long)scratch_bytes, (long long)scatter_scratch_bytes(n));
        return DCNR_ERR_WORKSPACE;
    }
    int tbit = 32;
    if (two) {
        tbit = std::max(sort_bits(rows0), sort_bits(rows1));
        DCNR_REQUIRE(tbit < 32, "tables too large to share a sort");
    }
    Arena ar(scratch, scratch_bytes);
    const int64_t n_windows = ceil_div(n, kWin);
    uint32_t *k0 = ar.take<uint32_t>(n), *k1 = ar.take<uint32_t>(n);
    uint32_t *v0 = ar.take<uint32_t>(n), *v1 = ar.take<uint32_t>(n);
    float *carry = ar.take<float>(n_windows * 2 * 256);
    uint8_t *flags = ar.take<uint8_t>(n_windows);
    void *temp = ar.take<char>(radix_sort_scratch_bytes(n));

    k_scatter_prep<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(ids0, stride0, rows0, ids1, stride1, rows1, B, tbit, k0, v0);
    DCNR_LAUNCHED();
    // stable LSD radix sort (radix_sort.cu, hand-written: 11-bit digits, two passes for the 21-bit keys of 1 M x 100 K tables)
    uint32_t *ks = nullptr, *vs = nullptr;
    DCNR_TRY(launch_radix_sort_pairs(k0, v0, k1, v1, n, two ? tbit + 1 : sort_bits(rows0), temp, &ks, &vs, stream));
    ScatterTables tabs;
    tabs.grad[0] = grad0; tabs.grad[1] = two ? grad1 : grad0;
    tabs.col0[0] = col0; tabs.col0[1] = two ? col1 : col0;
    tabs.tbit = tbit; tabs.B = B;
    const int64_t threads = n_windows * width;
    if (width <= 16)      // two sorted windows per warp (one window would leave half the lanes idle at width 16)
        k_scatter_window16<<<(unsigned)ceil_div(ceil_div(n_windows, 2), 4), 128, 0, stream>>>(ks, vs, n, width,
                                                                                              dx0, lddx, tabs, carry, flags, n_windows);
    else
        k_scatter_window<<<(unsigned)ceil_div(n_windows, 4), 128, 0, stream>>>(ks, vs, n, width, dx0, lddx, tabs,
                                                                             carry, flags, n_windows);
    DCNR_LAUNCHED();
    k_scatter_fixup<<<(unsigned)ceil_div(threads, 128), 128, 0, stream>>>(ks, n, width, tabs, carry, flags, n_windows);
    DCNR_LAUNCHED();
    return DCNR_OK;
}

int launch_embed_scatter(const int64_t *ids, int64_t id_stride, int64_t B, int64_t n_rows, int32_t width,
                         const float *dx0, int64_t lddx, int32_t col0, float *grad_table, void *scratch,
                         int64_t scratch_bytes, cudaStream_t stream) {
    DCNR_REQUIRE(width >= 1 && width <= 256, "embedding width %d unsupported", width);
    DCNR_REQUIRE(n_rows >= 1 && n_rows <= 0xffffffffLL, "table rows out of range");
    const int subs = small_subs_per_cta(n_rows, width);
    if (subs > 0 && B > 0) {
        const int64_t ctas = small_ctas(B, subs);
        const int64_t tsz = n_rows * width;
        if (scratch_bytes < ctas * tsz * 4) {
            set_error("scatter scratch too small");
            return DCNR_ERR_WORKSPACE;
        }
        float *partials = reinterpret_cast<float *>(scratch);
        const size_t smem = (size_t)subs * tsz * sizeof(float);
        if (smem > 48 * 1024)
            DCNR_CUDA_CHECK(cudaFuncSetAttribute(k_scatter_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_scatter_small<<<(unsigned)ctas, kSmallThreads, smem, stream>>>(ids, id_stride, B, (int)n_rows, width, dx0, lddx, col0,
                                                                       subs, partials);
        DCNR_LAUNCHED();
        return launch_sum_partials_2d(partials, ctas, (int32_t)n_rows, width, width, grad_table, width, stream);
    }
    DCNR_CUDA_CHECK(cudaMemsetAsync(grad_table, 0, (size_t)n_rows * width * sizeof(float), stream));
    if (B <= 0) return DCNR_OK;
    return launch_scatter_sorted(ids, id_stride, n_rows, grad_table, col0, nullptr, 0, 0, nullptr, 0, B, width, dx0, lddx, scratch,
                                 scratch_bytes, stream);
}

// user + item tables (same width, both too large for the tiny-table path): one shared sort
int launch_embed_scatter_pair(const int64_t *ids0, int64_t stride0, int64_t rows0, float *grad0, int32_t col0,
                              const int64_t *ids1, int64_t stride1, int64_t rows1, float *grad1, int32_t col1, int64_t B,
                              int32_t width, const float *dx0, int64_t lddx, void *scratch, int64_t scratch_bytes,
                              cudaStream_t stream) {
    const bool shareable = grad0 != nullptr && grad1 != nullptr && B > 0 && small_subs_per_cta(rows0, width) == 0 &&
                           small_subs_per_cta(rows1, width) == 0 && std::max(sort_bits(rows0), sort_bits(rows1)) < 31 &&
                           2 * B < 0x7fffffffLL && scratch_bytes >= scatter_scratch_bytes(2 * B);
    if (!shareable) {
        if (grad0 != nullptr)
            DCNR_TRY(launch_embed_scatter(ids0, stride0, B, rows0, width, dx0, lddx, col0, grad0, scratch, scratch_bytes, stream));
        if (grad1 != nullptr)
            DCNR_TRY(launch_embed_scatter(ids1, stride1, B, rows1, width, dx0, lddx, col1, grad1, scratch, scratch_bytes, stream));
        return DCNR_OK;
    }
    DCNR_CUDA_CHECK(cudaMemsetAsync(grad0, 0, (size_t)rows0 * width * sizeof(float), stream));
    DCNR_CUDA_CHECK(cudaMemsetAsync(grad1, 0, (size_t)rows1 * width * sizeof(float), stream));
    return launch_scatter_sorted(ids0, stride0, rows0, grad0, col0, ids1, stride1, rows1, grad1, col1, B, width, dx0, lddx, scratch,
                                 scratch_bytes, stream);
}

}  // namespace dcnr

using namespace dcnr;

extern "C" int dcnr_embed_scatter_bwd(const dcnr_dims *dims, const dcnr_batch *batch, const float *dx0, int64_t lddx,
                                      const dcnr_grads *grads, void *scratch, int64_t scratch_bytes,
                                      dcnr_stream_t stream) {
    DCNR_REQUIRE(dims && batch && dx0 && grads && scratch, "null argument");
    DCNR_REQUIRE(lddx >= dims->in_dim, "lddx too small");
    cudaStream_t st = as_stream(stream);
    const int E = dims->emb_dim;
    DCNR_TRY(launch_embed_scatter_pair(batch->user_ids, 1, dims->n_users, grads->user_table, 0, batch->item_ids, 1, dims->n_items,
                                       grads->item_table, E, batch->batch, E, dx0, lddx, scratch, scratch_bytes, st));
    int rc = DCNR_OK;
    if (try_scatter_small_multi(dims, batch, dx0, lddx, grads, scratch, scratch_bytes, st, &rc)) return rc;
    int col = 2 * E;
    for (int i = 0; i < dims->n_cat; ++i) {
        if (grads->cat_table[i])
            DCNR_TRY(launch_embed_scatter(batch->cat_features + i, dims->n_cat, batch->batch, dims->cat_rows[i],
                                          dims->cat_width[i], dx0, lddx, col, grads->cat_table[i], scratch,
                                          scratch_bytes, st));
        col += dims->cat_width[i];
    }
    return DCNR_OK;
}

// ---- owner side of the row-sharded embedding exchange --------------------------------------------------------
namespace dcnr {
// one float4 (or one float when the width is not a multiple of 4) per thread; rows are contiguous so a warp
// reads whole 64..256-byte table rows and writes a contiguous output
template <typename T>
__global__ void k_gather_rows(const T *__restrict__ table, int64_t rows, int wv, const int64_t *__restrict__ ids, int64_t n,
                              T *__restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * wv) return;
    const int64_t i = e / wv;
    const int c = (int)(e % wv);
    int64_t id = __ldg(ids + i);
    id = id < 0 ? 0 : (id >= rows ? rows - 1 : id);
    out[e] = __ldg(table + id * wv + c);
}
}  // namespace dcnr

extern "C" int dcnr_gather_rows(const float *table, int64_t rows, int32_t width, const int64_t *ids, int64_t n, float *out,
                                dcnr_stream_t stream) {
    DCNR_REQUIRE(table && ids && out && rows >= 1 && width >= 1, "bad argument");
    if (n <= 0) return DCNR_OK;
    cudaStream_t st = as_stream(stream);
    if (width % 4 == 0 && ((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
        const int wv = width / 4;
        k_gather_rows<float4><<<(unsigned)ceil_div(n * wv, 256), 256, 0, st>>>(reinterpret_cast<const float4 *>(table), rows, wv, ids,
                                                                            n, reinterpret_cast<float4 *>(out));
    } else {
        k_gather_rows<float><<<(unsigned)ceil_div(n * width, 256), 256, 0, st>>>(table, rows, width, ids, n, out);
    }
    DCNR_LAUNCHED();
    return DCNR_OK;
}

extern "C" int dcnr_scatter_rows(const int64_t *ids, int64_t n, int64_t rows, int32_t width, const float *g, int64_t ldg,
                                 float *grad_table, void *scratch, int64_t scratch_bytes, dcnr_stream_t stream) {
    DCNR_REQUIRE(ids && g && grad_table && scratch && ldg >= width, "bad argument");
    return launch_embed_scatter(ids, 1, n, rows, width, g, ldg, 0, grad_table, scratch, scratch_bytes, as_stream(stream));
}
