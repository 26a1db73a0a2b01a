// K9/K10: cosine top-k over the item-embedding catalog and the cross-shard merge.
//
// Replaces NearestNeighbors(metric='cosine', algorithm='brute').fit / .kneighbors at
// main.py:268-269 (fit), :200 (candidate generation, k = 11) and :300 (/similar_items, k = n+1).
// Arithmetic contract (bit-exact with oracle/knn_oracle.c, see its header): sequential-fma norms
// and dot products, dist = clip(1 - sim, 0, 2), total order (dist ascending, index ascending).
//
// Scan kernel: the catalog is cut into contiguous slices, one CTA per (slice, query tile).  Each
// thread scores one catalog row per round against the tile's queries (queries broadcast from
// shared memory).  A candidate is kept only if its 64-bit key (dist bits << 32 | row) beats the
// CTA's current k-th best key for that query; survivors go to a small shared buffer that is
// bitonic-sorted and truncated to k whenever it could overflow.  Keys are unique, so the selection
// is deterministic although the append order is not.  Per-slice lists are then merged by sorting
// groups of lists in shared memory.
#include "kernels.cuh"

namespace dcnr {

constexpr int kTT = 256;        // threads per CTA
constexpr int kQT = 8;          // queries per CTA tile
constexpr int kMaxD = 128;      // embedding dim limit (floats), multiple of 4
constexpr int kMergeMaxKeys = 8192;
constexpr int kCap = 512;       // per-query candidate buffer: >= k + kTT, power of two (k <= 256)
typedef unsigned long long u64;
constexpr u64 kMaxKey = ~0ull;

__device__ __forceinline__ void bitonic_sort_u64(u64 *a, int n, int tid, int nthreads) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < (n >> 1); i += nthreads) {
                const int lo = ((i / stride) * (stride << 1)) + (i % stride);
                const int hi = lo + stride;
                const bool asc = (lo & size) == 0;
                const u64 x = a[lo], y = a[hi];
                if ((x > y) == asc) { a[lo] = y; a[hi] = x; }
            }
            __syncthreads();
        }
    }
}

__global__ void k_knn_normalize(const float *__restrict__ in, float *__restrict__ out, int64_t n, int d) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const float *v = in + r * d;
    float ss = 0.f;
    for (int j = 0; j < d; ++j) ss = __fmaf_rn(v[j], v[j], ss);
    float nrm = __fsqrt_rn(ss);
    if (nrm == 0.f) nrm = 1.f;
    for (int j = 0; j < d; ++j) out[r * d + j] = __fdiv_rn(v[j], nrm);
}

template <int DV>   // DV = d/4 float4 per row (0 = runtime d)
__global__ void __launch_bounds__(kTT)
k_knn_scan(const float *__restrict__ cat, int64_t n, int d, const float *__restrict__ queries, int nq, int k, int kp,
           int64_t rows_per_slice, u64 *__restrict__ out_keys) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int cap = kCap;
    u64 *buf = reinterpret_cast<u64 *>(smem_raw);                      // [kQT][cap]
    u64 *tau = buf + (size_t)kQT * cap;                                // [kQT]
    int *cnt = reinterpret_cast<int *>(tau + kQT);                     // [kQT]
    float *sq = reinterpret_cast<float *>(cnt + kQT);                  // [kQT][d]
    const int tid = threadIdx.x;
    const int q0 = blockIdx.y * kQT;
    const int nqt = min(kQT, nq - q0);
    for (int i = tid; i < kQT * d; i += kTT) sq[i] = (i / d) < nqt ? queries[(int64_t)q0 * d + i] : 0.f;
    if (tid < kQT) { tau[tid] = kMaxKey; cnt[tid] = 0; }
    __syncthreads();

    const int64_t s0 = (int64_t)blockIdx.x * rows_per_slice;
    const int64_t s1 = min(n, s0 + rows_per_slice);
    const int dv = DV > 0 ? DV : (d >> 2);
    for (int64_t base = s0; base < s1; base += kTT) {
        const int64_t row = base + tid;
        if (row < s1) {
            float sim[kQT];
#pragma unroll
            for (int q = 0; q < kQT; ++q) sim[q] = 0.f;
            const float4 *rp = reinterpret_cast<const float4 *>(cat + row * d);
#pragma unroll 4
            for (int j = 0; j < dv; ++j) {
                const float4 e = __ldg(rp + j);
#pragma unroll
                for (int q = 0; q < kQT; ++q) {
                    const float4 qv = *reinterpret_cast<const float4 *>(sq + q * d + 4 * j);
                    sim[q] = __fmaf_rn(qv.x, e.x, sim[q]);
                    sim[q] = __fmaf_rn(qv.y, e.y, sim[q]);
                    sim[q] = __fmaf_rn(qv.z, e.z, sim[q]);
                    sim[q] = __fmaf_rn(qv.w, e.w, sim[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < kQT; ++q) {
                if (q < nqt) {
                    float dist = __fsub_rn(1.0f, sim[q]);
                    dist = fminf(fmaxf(dist, 0.f), 2.f);
                    const u64 key = ((u64)__float_as_uint(dist) << 32) | (u64)(uint32_t)row;
                    if (key < tau[q]) {
                        const int pos = atomicAdd(&cnt[q], 1);
                        buf[(size_t)q * cap + pos] = key;      // pos < cap: cnt <= cap - kTT at round start
                    }
                }
            }
        }
        __syncthreads();
        for (int q = 0; q < nqt; ++q) {
            const int c = cnt[q];                              // uniform across the CTA
            if (c > cap - kTT) {
                u64 *b = buf + (size_t)q * cap;
                for (int i = c + tid; i < cap; i += kTT) b[i] = kMaxKey;
                __syncthreads();
                bitonic_sort_u64(b, cap, tid, kTT);
                if (tid == 0) {
                    cnt[q] = min(c, k);
                    if (c >= k) tau[q] = b[k - 1];
                }
                __syncthreads();
            }
        }
    }
    // final: sort every query's survivors and emit the slice's best kp keys
    for (int q = 0; q < nqt; ++q) {
        const int c = cnt[q];
        u64 *b = buf + (size_t)q * cap;
        for (int i = c + tid; i < cap; i += kTT) b[i] = kMaxKey;
        __syncthreads();
        bitonic_sort_u64(b, cap, tid, kTT);
        u64 *o = out_keys + ((int64_t)blockIdx.x * nq + (q0 + q)) * kp;
        for (int i = tid; i < kp; i += kTT) o[i] = (i < k) ? b[i] : kMaxKey;
        __syncthreads();
    }
}

// Sorts `group` lists of kp keys per query in shared memory and keeps the best kp.
// in: [n_lists][nq][kp]; out: [ceil(n_lists/group)][nq][kp].  When dist_out != NULL this is the
// last level and (dist, idx_base + row) are written instead.
__global__ void __launch_bounds__(kTT)
k_knn_merge_keys(const u64 *__restrict__ in, int n_lists, int nq, int kp, int k, int group, int n_sort,
                 u64 *__restrict__ out, float *__restrict__ dist_out, int64_t *__restrict__ idx_out, int64_t idx_base) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *a = reinterpret_cast<u64 *>(smem_raw);
    const int q = blockIdx.x, g = blockIdx.y, tid = threadIdx.x;
    const int l0 = g * group, l1 = min(n_lists, l0 + group);
    const int have = (l1 - l0) * kp;
    for (int i = tid; i < n_sort; i += kTT) {
        u64 v = kMaxKey;
        if (i < have) v = in[((int64_t)(l0 + i / kp) * nq + q) * kp + (i % kp)];
        a[i] = v;
    }
    __syncthreads();
    bitonic_sort_u64(a, n_sort, tid, kTT);
    if (dist_out != nullptr) {
        for (int i = tid; i < k; i += kTT) {
            const u64 key = a[i];
            if (key == kMaxKey) {
                dist_out[(int64_t)q * k + i] = __int_as_float(0x7f800000);
                idx_out[(int64_t)q * k + i] = -1;
            } else {
                dist_out[(int64_t)q * k + i] = __uint_as_float((uint32_t)(key >> 32));
                idx_out[(int64_t)q * k + i] = idx_base + (int64_t)(uint32_t)key;
            }
        }
    } else {
        for (int i = tid; i < kp; i += kTT) out[((int64_t)g * nq + q) * kp + i] = a[i];
    }
}

// Cross-shard merge on (dist, int64 idx) pairs, lexicographic order, (inf,-1) padding sorts last.
struct DI { float d; int64_t i; };
__device__ __forceinline__ bool di_greater(float xd, int64_t xi, float yd, int64_t yi) {
    const bool xs = xi < 0, ys = yi < 0;          // sentinels after everything
    if (xs != ys) return xs;
    return (xd > yd) || (xd == yd && xi > yi);
}
__global__ void __launch_bounds__(kTT)
k_knn_merge_parts(const float *__restrict__ dist_parts, const int64_t *__restrict__ idx_parts, int n_parts, int nq, int k,
                  int n_sort, float *__restrict__ dist_out, int64_t *__restrict__ idx_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int64_t *si = reinterpret_cast<int64_t *>(smem_raw);
    float *sd = reinterpret_cast<float *>(si + n_sort);
    const int q = blockIdx.x, tid = threadIdx.x;
    const int have = n_parts * k;
    for (int i = tid; i < n_sort; i += kTT) {
        float d = __int_as_float(0x7f800000);
        int64_t id = -1;
        if (i < have) {
            const int64_t off = ((int64_t)(i / k) * nq + q) * k + (i % k);
            d = dist_parts[off];
            id = idx_parts[off];
        }
        sd[i] = d;
        si[i] = id;
    }
    __syncthreads();
    for (int size = 2; size <= n_sort; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < (n_sort >> 1); i += kTT) {
                const int lo = ((i / stride) * (stride << 1)) + (i % stride), hi = lo + stride;
                const bool asc = (lo & size) == 0;
                const float xd = sd[lo], yd = sd[hi];
                const int64_t xi = si[lo], yi = si[hi];
                if (di_greater(xd, xi, yd, yi) == asc) { sd[lo] = yd; sd[hi] = xd; si[lo] = yi; si[hi] = xi; }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < k; i += kTT) {
        dist_out[(int64_t)q * k + i] = si[i] < 0 ? __int_as_float(0x7f800000) : sd[i];
        idx_out[(int64_t)q * k + i] = si[i];
    }
}

static int next_pow2(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

struct KnnPlan {
    int kp, tiles, slices, group;
    int64_t rows_per_slice;
    int levels;
    int64_t lists[8];   // list count entering each merge level
};

static int make_plan(int64_t n, int nq, int k, KnnPlan *p) {
    p->kp = std::max(next_pow2(k), 32);
    p->tiles = (int)ceil_div(nq, kQT);
    int64_t want = std::max<int64_t>(1, ceil_div(2 * (int64_t)sm_count(), p->tiles));
    int64_t rows = round_up(ceil_div(std::max<int64_t>(n, 1), want), kTT);
    p->rows_per_slice = rows;
    p->slices = (int)std::max<int64_t>(1, ceil_div(n, rows));
    p->group = std::max(2, kMergeMaxKeys / p->kp);
    p->levels = 0;
    int64_t cur = p->slices;
    do {
        p->lists[p->levels++] = cur;
        cur = ceil_div(cur, p->group);
    } while (cur > 1 && p->levels < 8);
    return DCNR_OK;
}

}  // namespace dcnr

using namespace dcnr;

extern "C" int dcnr_knn_normalize(const float *in, float *out, int64_t n, int32_t d, dcnr_stream_t stream) {
    DCNR_REQUIRE(in && out && d >= 1, "null argument");
    if (n <= 0) return DCNR_OK;
    k_knn_normalize<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(in, out, n, d);
    DCNR_LAUNCHED();
    return DCNR_OK;
}

extern "C" int64_t dcnr_knn_scratch_bytes(int64_t n, int32_t d, int32_t n_queries, int32_t k) {
    (void)d;
    if (n_queries <= 0 || k <= 0) return 256;
    KnnPlan p;
    make_plan(n, n_queries, k, &p);
    int64_t bytes = 0;
    for (int l = 0; l < p.levels; ++l) bytes += round_up(p.lists[l] * n_queries * (int64_t)p.kp * 8, 256);
    return bytes + 256;
}

extern "C" int dcnr_knn_topk(const float *catalog_hat, int64_t n, int32_t d, const float *queries_hat, int32_t n_queries,
                             int32_t k, int64_t idx_base, float *dist_out, int64_t *idx_out, void *scratch,
                             int64_t scratch_bytes, dcnr_stream_t stream) {
    DCNR_REQUIRE(catalog_hat && queries_hat && dist_out && idx_out && scratch, "null argument");
    DCNR_REQUIRE(d >= 4 && d % 4 == 0 && d <= kMaxD, "embedding dim %d unsupported (multiple of 4, <= %d)", d, kMaxD);
    DCNR_REQUIRE(k >= 1 && k <= 256, "k %d unsupported (1..256)", k);
    DCNR_REQUIRE(n >= 0 && n <= 0xffffffffLL, "catalog shard too large");
    DCNR_REQUIRE(((uintptr_t)catalog_hat & 15) == 0, "catalog must be 16-byte aligned");
    if (n_queries <= 0) return DCNR_OK;
    if (scratch_bytes < dcnr_knn_scratch_bytes(n, d, n_queries, k)) {
        set_error("knn scratch too small");
        return DCNR_ERR_WORKSPACE;
    }
    cudaStream_t st = as_stream(stream);
    KnnPlan p;
    make_plan(n, n_queries, k, &p);
    Arena ar(scratch, scratch_bytes);
    u64 *level_buf[8];
    for (int l = 0; l < p.levels; ++l) level_buf[l] = ar.take<u64>(p.lists[l] * n_queries * (int64_t)p.kp);

    const size_t smem = (size_t)kQT * kCap * 8 + kQT * 8 + kQT * 4 + (size_t)kQT * d * 4;
    dim3 grid((unsigned)p.slices, (unsigned)p.tiles);
#define DCNR_SCAN(DVV)                                                                                              \
    do {                                                                                                            \
        if (smem > 48 * 1024)                                                                                       \
            DCNR_CUDA_CHECK(cudaFuncSetAttribute(k_knn_scan<DVV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_knn_scan<DVV><<<grid, kTT, smem, st>>>(catalog_hat, n, d, queries_hat, n_queries, k, p.kp, p.rows_per_slice, \
                                                 level_buf[0]);                                                     \
    } while (0)
    if (d == 16) DCNR_SCAN(4);
    else if (d == 32) DCNR_SCAN(8);
    else if (d == 64) DCNR_SCAN(16);
    else DCNR_SCAN(0);
#undef DCNR_SCAN
    DCNR_LAUNCHED();
    for (int l = 0; l < p.levels; ++l) {
        const int n_lists = (int)p.lists[l];
        const bool last = l == p.levels - 1;
        const int group = last ? n_lists : p.group;
        const int n_sort = next_pow2(std::min(group, n_lists) * p.kp);
        DCNR_REQUIRE(n_sort <= 2 * kMergeMaxKeys, "top-k merge fan-in too large");
        const size_t msmem = (size_t)n_sort * 8;
        if (msmem > 48 * 1024)
            DCNR_CUDA_CHECK(cudaFuncSetAttribute(k_knn_merge_keys, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem));
        dim3 mg((unsigned)n_queries, (unsigned)ceil_div(n_lists, group));
        k_knn_merge_keys<<<mg, kTT, msmem, st>>>(level_buf[l], n_lists, n_queries, p.kp, k, group, n_sort,
                                                 last ? nullptr : level_buf[l + 1], last ? dist_out : nullptr,
                                                 last ? idx_out : nullptr, idx_base);
        DCNR_LAUNCHED();
    }
    return DCNR_OK;
}

extern "C" int dcnr_knn_merge(const float *dist_parts, const int64_t *idx_parts, int32_t n_parts, int32_t n_queries,
                              int32_t k, float *dist_out, int64_t *idx_out, dcnr_stream_t stream) {
    DCNR_REQUIRE(dist_parts && idx_parts && dist_out && idx_out, "null argument");
    DCNR_REQUIRE(n_parts >= 1 && k >= 1, "bad n_parts / k");
    if (n_queries <= 0) return DCNR_OK;
    const int n_sort = next_pow2(n_parts * k);
    DCNR_REQUIRE(n_sort <= 16384, "n_parts * k = %d too large for one merge (max 16384)", n_parts * k);
    const size_t smem = (size_t)n_sort * 12;
    if (smem > 48 * 1024)
        DCNR_CUDA_CHECK(cudaFuncSetAttribute(k_knn_merge_parts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_knn_merge_parts<<<(unsigned)n_queries, kTT, smem, as_stream(stream)>>>(dist_parts, idx_parts, n_parts, n_queries, k,
                                                                            n_sort, dist_out, idx_out);
    DCNR_LAUNCHED();
    return DCNR_OK;
}
