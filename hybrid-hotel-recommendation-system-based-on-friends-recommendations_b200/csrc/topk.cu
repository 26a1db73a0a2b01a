// K9/K10: cosine top-k over the item-embedding catalog and the cross-shard merge.
//
// Replaces NearestNeighbors(metric='cosine', algorithm='brute').fit / .kneighbors at
// main.py:268-269 (fit), :200 (candidate generation, k = 11) and :300 (/similar_items, k = n+1).
// Arithmetic contract (bit-exact with oracle/knn_oracle.c, see its header): sequential-fma norms
// and dot products, dist = clip(1 - sim, 0, 2), total order (dist ascending, index ascending).
//
// Scan kernel: the catalog is cut into contiguous slices, one CTA per (slice, query tile of 1 or 8).
// Every warp streams groups of 32 consecutive rows through its own cp.async ring (16-byte
// LDGSTS, fully coalesced 512-byte requests, 2-5 groups in flight per warp so ~50-100 KB per SM are
// outstanding -- the first version's one-row-per-thread __ldg loop had 16 KB and ran at 18 % of HBM);
// rows are staged with a one-float4 pad so that each lane then reads ITS row with conflict-free
// LDS.128 and scores it against the tile's queries (broadcast from shared memory) with the
// sequential-fma order of the oracle.  A candidate is kept only if its 64-bit key
// (dist bits << 32 | row) beats the CTA's current k-th best key for that query; survivors go to a
// shared buffer that is bitonic-sorted and truncated to k whenever it could overflow.  Keys are
// unique, so the selection is deterministic although the append order is not.  Per-slice lists are
// then merged by sorting groups of lists in shared memory.  (k_knn_scan below is the generic-width
// fallback for embedding widths outside {16, 24, 32, 48, 64}.)
#include <cooperative_groups.h>

#include "kernels.cuh"

namespace dcnr {

constexpr int kTT = 256;        // threads per CTA
constexpr int kQT = 8;          // queries per CTA tile
constexpr int kMaxD = 128;      // embedding dim limit (floats), multiple of 4
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

// Merge of the per-slice lists ([n_lists][nq][kp], each sorted ascending, kMaxKey padded) into the final
// top k of one query per CTA.  A threshold prunes first: with j = the probe depth and r = ceil(k / j), the
// r-th smallest of the lists' j-th entries T bounds the k-th best key from above (r lists hold j keys <= T
// each), so only list prefixes <= T can matter -- typically a few hundred keys instead of n_lists * k.  The
// prefixes are appended to a candidate buffer 16 lists at a time and the buffer is sorted / truncated to
// k (which tightens T) whenever the next chunk might not fit, so any input is handled; the usual case is
// one small sort.  tau_out != NULL: only the k-th best key is wanted (the sampled pre-pass).
constexpr int kCandCap = 8192;
constexpr int kChunkLists = 16;
__device__ void merge_select_body(const u64 *__restrict__ lists, int n_lists, int nq, int kp, int k, int n_probe_sort, int q,
                                  unsigned char *smem_raw, float *__restrict__ dist_out, int64_t *__restrict__ idx_out,
                                  int64_t idx_base, u64 *__restrict__ tau_out) {
    u64 *cand = reinterpret_cast<u64 *>(smem_raw);                    // [kCandCap]
    u64 *probe = cand + kCandCap;                                      // [n_probe_sort]
    __shared__ int s_cnt;
    __shared__ u64 s_tau;
    const int tid = threadIdx.x;
    const int klist = min(k, kp);
    int j = max(1, (2 * k + n_lists - 1) / n_lists);
    j = min(j, klist);
    int r = (k + j - 1) / j;
    if (r > n_lists) {                                                 // few lists: probe deeper
        j = min(klist, (k + n_lists - 1) / n_lists);
        r = min(n_lists, (k + j - 1) / j);
    }
    for (int i = tid; i < n_probe_sort; i += kTT)
        probe[i] = i < n_lists ? lists[((int64_t)i * nq + q) * kp + (j - 1)] : kMaxKey;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    bitonic_sort_u64(probe, n_probe_sort, tid, kTT);
    if (tid == 0) s_tau = (int64_t)r * j >= k ? probe[r - 1] : kMaxKey;  // inclusive bound (kMaxKey: no pruning)
    __syncthreads();

    auto compact = [&]() {                                             // keep the best k candidates, tighten the bound
        const int c = s_cnt;
        int n_sort = 32;
        while (n_sort < c) n_sort <<= 1;
        for (int i = c + tid; i < n_sort; i += kTT) cand[i] = kMaxKey;
        __syncthreads();
        bitonic_sort_u64(cand, n_sort, tid, kTT);
        if (tid == 0) {
            s_cnt = min(c, k);
            if (c >= k) s_tau = min(s_tau, cand[k - 1]);
        }
        __syncthreads();
    };
    // Lists are sorted, so the keys <= tau of a list are a prefix: 16 lanes read a list 16 keys at a time and stop at
    // the first block that is not entirely below the bound (with a tight bound that is the first block).  Four
    // chunks of 16 lists are in flight per pass, so the usual merge is a handful of dependent memory round trips.
    constexpr int kLanes = kTT / kChunkLists;                          // 16 lanes per list
    const int sub = tid % kLanes, grp = tid / kLanes;
    const unsigned gmask = 0xffffu << ((tid & 31) & ~(kLanes - 1));   // this list's lanes inside the warp
    for (int l0 = 0; l0 < n_lists; l0 += 4 * kChunkLists) {
        if (s_cnt > kCandCap - 4 * kChunkLists * kLanes) compact();   // uniform: s_cnt is read after a barrier
        const u64 tau = s_tau;
        u64 first[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int li = l0 + u * kChunkLists + grp;
            first[u] = (li < n_lists && sub < klist) ? lists[((int64_t)li * nq + q) * kp + sub] : kMaxKey;
        }
        bool more = false;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool take = first[u] <= tau && first[u] != kMaxKey;
            if (take) cand[atomicAdd(&s_cnt, 1)] = first[u];
            more |= __all_sync(gmask, take) && kLanes < klist;         // whole block below the bound: the list goes on
        }
        __syncthreads();
        if (__syncthreads_or(more)) {                                  // rare: loose bound or long runs in one list
            for (int u = 0; u < 4; ++u) {
                const int li = l0 + u * kChunkLists + grp;
                for (int b0 = kLanes; b0 < klist; b0 += kLanes) {      // uniform trip count; appends <= 16 lists x 16 per pass
                    if (s_cnt > kCandCap - kChunkLists * kLanes) compact();
                    const u64 t2 = s_tau;
                    const u64 key = (li < n_lists && b0 + sub < klist) ? lists[((int64_t)li * nq + q) * kp + b0 + sub] : kMaxKey;
                    if (key <= t2 && key != kMaxKey) cand[atomicAdd(&s_cnt, 1)] = key;
                    __syncthreads();
                }
            }
        }
    }
    compact();
    const int c = s_cnt;
    if (tau_out != nullptr) {
        if (tid == 0) tau_out[q] = c >= k ? cand[k - 1] : kMaxKey;
        return;
    }
    for (int i = tid; i < k; i += kTT) {
        const u64 key = i < c ? cand[i] : kMaxKey;
        if (key == kMaxKey) {
            dist_out[(int64_t)q * k + i] = __int_as_float(0x7f800000);
            idx_out[(int64_t)q * k + i] = -1;
        } else {
            dist_out[(int64_t)q * k + i] = __uint_as_float((uint32_t)(key >> 32));
            idx_out[(int64_t)q * k + i] = idx_base + (int64_t)(uint32_t)key;
        }
    }
}

__global__ void __launch_bounds__(kTT)
k_knn_merge_select(const u64 *__restrict__ lists, int n_lists, int nq, int kp, int k, int n_probe_sort,
                   float *__restrict__ dist_out, int64_t *__restrict__ idx_out, int64_t idx_base,
                   u64 *__restrict__ tau_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    merge_select_body(lists, n_lists, nq, kp, k, n_probe_sort, blockIdx.x, smem_raw, dist_out, idx_out, idx_base, tau_out);
}

constexpr int kCap2 = 1024;     // candidate buffer of the single-query streaming kernel (checked every 2nd round)
constexpr int kCap8 = 1024;     // ... of the 8-query kernel (one CTA per SM; 512 with two CTAs per SM measured 1.6x slower)

__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int DV, int QT>   // DV float4 per row (d = 4*DV), QT queries per CTA
__global__ void __launch_bounds__(kTT)
k_knn_stream(const float *__restrict__ cat, int64_t n, const float *__restrict__ queries, int nq, int k, int kp,
             int64_t groups_per_slice, int64_t gstride, int n_stages, const u64 *__restrict__ tau0,
             u64 *__restrict__ out_keys, int fused, int warm_rounds, u64 *__restrict__ probes, int n_probe_sort,
             float *__restrict__ dist_out, int64_t *__restrict__ idx_out, int64_t idx_base) {
    constexpr int d = 4 * DV;
    constexpr int ROWQ = DV + ((DV & 1) ? 0 : 1);      // staged row pitch in float4: odd => conflict-free LDS.128
    constexpr int cap = QT == 1 ? kCap2 : kCap8;
    constexpr int kCheck = 2;                          // rounds between CTA-wide check points (<= kCheck * kTT appends)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *buf = reinterpret_cast<u64 *>(smem_raw);                      // [QT][cap]
    u64 *tau = buf + (size_t)QT * cap;                                 // [8] (QT used)
    int *cnt = reinterpret_cast<int *>(tau + 8);                       // [8] (QT used)
    unsigned *flush_mask = reinterpret_cast<unsigned *>(cnt + 8);      // [1] (+3 pad) -- every section stays 16-byte aligned
    float *sq = reinterpret_cast<float *>(flush_mask + 4);             // [QT][d]
    float4 *ring = reinterpret_cast<float4 *>(sq + QT * d);            // [8 warps][n_stages][32][ROWQ]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.y * QT;
    const int nqt = min(QT, nq - q0);
    for (int i = tid; i < QT * d; i += kTT) sq[i] = (i / d) < nqt ? queries[(int64_t)q0 * d + i] : 0.f;
    if (tid < QT) {
        // a sampled pre-pass (same kernel, gstride > 1) may hand in the k-th best key of a row subset: every key
        // above it is already known not to be in the top k, so the buffers below almost never fill
        u64 t0 = kMaxKey;
        if (tau0 != nullptr && tid < nqt && tau0[q0 + tid] != kMaxKey) t0 = tau0[q0 + tid] + 1;
        tau[tid] = t0;
        cnt[tid] = 0;
    }
    if (tid == 0) *flush_mask = 0u;
    __syncthreads();

    // the slice is a range of 32-row groups; with gstride > 1 only every gstride-th group of the catalog is visited
    const int64_t n_groups = (n + 32 * gstride - 1) / (32 * gstride);
    const int64_t vg0 = (int64_t)blockIdx.x * groups_per_slice;
    const int64_t vg1 = min(n_groups, vg0 + groups_per_slice);
    const int iters = vg1 > vg0 ? (int)((vg1 - vg0 + 7) / 8) : 0;        // one group per warp per iteration
    float4 *wring = ring + (size_t)warp * n_stages * 32 * ROWQ;
    const float4 *cat4 = reinterpret_cast<const float4 *>(cat);

    auto issue = [&](int it) {                           // this warp's group of iteration `it` into ring slot it % n_stages
        const int64_t vg = vg0 + (int64_t)it * 8 + warp;
        if (it < iters && vg < vg1) {
            const int64_t g0 = vg * gstride * 32;         // first row of the group
            float4 *st = wring + (size_t)(it % n_stages) * 32 * ROWQ;
#pragma unroll
            for (int i = 0; i < DV; ++i) {
                const int f = lane + 32 * i;              // float4 index inside the group: row f / DV, part f % DV
                const int r = f / DV, part = f % DV;
                if (g0 + r < n) cp_async16(st + r * ROWQ + part, cat4 + g0 * DV + f);
            }
        }
        cp_async_commit();
    };
    auto flush = [&](int q) {                            // CTA-wide: keep the best k of buffer q, tighten tau
        const int c = min(cnt[q], cap);
        int n_sort = 32;
        while (n_sort < c) n_sort <<= 1;
        u64 *b = buf + (size_t)q * cap;
        for (int i = c + tid; i < n_sort; i += kTT) b[i] = kMaxKey;
        __syncthreads();
        bitonic_sort_u64(b, n_sort, tid, kTT);
        if (tid == 0) {
            cnt[q] = min(c, k);
            if (c >= k) tau[q] = min(tau[q], b[k - 1]);
        }
        __syncthreads();
    };

    // Single-launch mode (cooperative grid): after `warm_rounds` rounds every CTA publishes the 32 best keys it has
    // seen per query; after a grid-wide barrier every CTA derives the same upper bound on the final k-th best key
    // from those lists (r lists with j keys <= T each, r * j >= k) and continues with it as its threshold, so the rest
    // of the scan appends only a few hundred keys per CTA -- what the separate sampled pre-pass achieves with two
    // extra launches.  A second barrier at the end lets CTA x of a tile merge query x, so the whole top-k is ONE launch.
    constexpr int kProbe = 32;
    const int slices = gridDim.x;
    auto exchange_bound = [&]() {
        for (int q = 0; q < nqt; ++q) {
            flush(q);
            const u64 *b = buf + (size_t)q * cap;
            const int c = cnt[q];
            u64 *o = probes + ((int64_t)blockIdx.x * nq + (q0 + q)) * kProbe;
            for (int i = tid; i < kProbe; i += kTT) o[i] = i < c ? b[i] : kMaxKey;
        }
        cooperative_groups::this_grid().sync();
        int j = max(1, (2 * k + slices - 1) / slices);
        j = min(j, min(k, kProbe));
        int r = (k + j - 1) / j;
        if (r <= slices && (int64_t)r * j >= k) {
            for (int q = 0; q < nqt; ++q) {
                u64 *sc = buf + (size_t)q * cap + cap / 2;                  // upper half of the buffer is free (cnt <= k <= cap / 2)
                for (int i = tid; i < n_probe_sort; i += kTT)
                    sc[i] = i < slices ? probes[((int64_t)i * nq + (q0 + q)) * kProbe + (j - 1)] : kMaxKey;
                __syncthreads();
                bitonic_sort_u64(sc, n_probe_sort, tid, kTT);
                if (tid == 0 && sc[r - 1] != kMaxKey) tau[q] = min(tau[q], sc[r - 1] + 1);
                __syncthreads();
            }
        }
    };
    bool exchanged = false;
    for (int i = 0; i < n_stages - 1; ++i) issue(i);
    for (int it = 0; it < iters; ++it) {
        issue(it + n_stages - 1);
        // all but the n_stages-1 most recent groups have landed => group `it` is in shared memory
        switch (n_stages) {
            case 2: cp_async_wait<1>(); break;
            case 3: cp_async_wait<2>(); break;
            case 4: cp_async_wait<3>(); break;
            case 5: cp_async_wait<4>(); break;
            default: cp_async_wait<5>(); break;
        }
        __syncwarp();
        const int64_t vg = vg0 + (int64_t)it * 8 + warp;
        const int64_t row = vg < vg1 ? vg * gstride * 32 + lane : n;
        const float4 *rp = wring + (size_t)(it % n_stages) * 32 * ROWQ + lane * ROWQ;
        float sim[QT];
#pragma unroll
        for (int q = 0; q < QT; ++q) sim[q] = 0.f;
#pragma unroll
        for (int j = 0; j < DV; ++j) {
            const float4 e = rp[j];
#pragma unroll
            for (int q = 0; q < QT; ++q) {
                const float4 qv = *reinterpret_cast<const float4 *>(sq + q * d + 4 * j);
                sim[q] = __fmaf_rn(qv.x, e.x, sim[q]);
                sim[q] = __fmaf_rn(qv.y, e.y, sim[q]);
                sim[q] = __fmaf_rn(qv.z, e.z, sim[q]);
                sim[q] = __fmaf_rn(qv.w, e.w, sim[q]);
            }
        }
        if (row < n) {
#pragma unroll
            for (int q = 0; q < QT; ++q) {
                if (q < nqt) {
                    float dist = __fsub_rn(1.0f, sim[q]);
                    dist = fminf(fmaxf(dist, 0.f), 2.f);
                    const u64 key = ((u64)__float_as_uint(dist) << 32) | (u64)(uint32_t)row;
                    if (key < tau[q]) {
                        const int pos = atomicAdd(&cnt[q], 1);          // < cap: cnt <= cap - kCheck kTT at every check point
                        buf[(size_t)q * cap + pos] = key;
                        if (pos >= cap - kCheck * kTT) atomicOr(flush_mask, 1u << q);
                    }
                }
            }
        }
        __syncwarp();                                     // every lane of this warp is done with its ring slot
        // CTA-wide check point every kCheck rounds (<= kCheck * kTT appends per query in between)
        if ((it % kCheck) != kCheck - 1 && it + 1 < iters && !(fused && !exchanged && it + 1 == warm_rounds)) continue;
        __syncthreads();                                  // appends visible
        const unsigned mask = *reinterpret_cast<volatile unsigned *>(flush_mask);
        if (mask != 0u) {                                 // CTA-uniform
            __syncthreads();
            if (tid == 0) *flush_mask = 0u;
            for (int q = 0; q < nqt; ++q)
                if (mask & (1u << q)) flush(q);
        }
        if (fused && !exchanged && it + 1 == warm_rounds) {
            exchange_bound();
            exchanged = true;
        }
    }
    if (fused && !exchanged) exchange_bound();           // every CTA takes part in the barrier exactly once
    cp_async_wait<0>();
    // final: sort every query's survivors and emit the slice's best kp keys
    for (int q = 0; q < nqt; ++q) {
        flush(q);
        const u64 *b = buf + (size_t)q * cap;
        const int c = cnt[q];
        u64 *o = out_keys + ((int64_t)blockIdx.x * nq + (q0 + q)) * kp;
        for (int i = tid; i < kp; i += kTT) o[i] = (i < c) ? b[i] : kMaxKey;
        __syncthreads();
    }
    if (fused) {
        cooperative_groups::this_grid().sync();
        for (int ql = blockIdx.x; ql < nqt; ql += gridDim.x)
            merge_select_body(out_keys, slices, nq, kp, k, n_probe_sort, q0 + ql, smem_raw, dist_out, idx_out, idx_base, nullptr);
    }
}

static int stream_stage_bytes(int dv) { return 32 * (dv + ((dv & 1) ? 0 : 1)) * 16; }
static bool stream_supported(int d) { return d == 16 || d == 24 || d == 32 || d == 48 || d == 64; }
static int stream_stages(int d, int qt) {
    const int budget = (qt == 1 ? 84 : 136) * 1024;                   // ring bytes: two CTAs per SM (1 query) / one (8 queries)
    return std::max(2, std::min(6, budget / (8 * stream_stage_bytes(d / 4))));
}
static size_t stream_smem(int d, int qt) {
    const size_t scan = (size_t)qt * (qt == 1 ? kCap2 : kCap8) * 8 + 8 * 8 + 8 * 4 + 16 + (size_t)qt * d * 4 +
                        (size_t)8 * stream_stages(d, qt) * stream_stage_bytes(d / 4);
    return std::max(scan, (size_t)(kCandCap + 1024) * 8);            // the fused epilogue merges in the same memory
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
    int kp, qt, tiles, slices;
    int64_t groups_per_slice;      // 32-row groups per CTA of the main scan
    bool stream;                   // streaming kernel (supported width) or the generic fallback
    int warm_rounds;               // single-launch mode: CTA rounds before the bound exchange (~64 K rows over the grid)
    // sampled pre-pass (streaming kernel only, large catalogs): every pre_gstride-th group, pre_slices CTAs per tile
    int pre_slices;
    int64_t pre_gstride, pre_groups_per_slice;
};

constexpr int64_t kPrepassMinRows = 1 << 18;      // below this the local filters are good enough
constexpr int64_t kPrepassSampleRows = 1 << 16;

static int make_plan(int64_t n, int32_t d, int nq, int k, KnnPlan *p) {
    p->kp = std::max(next_pow2(k), 32);
    p->stream = stream_supported(d);
    p->qt = (p->stream && nq == 1) ? 1 : kQT;
    p->tiles = (int)ceil_div(nq, p->qt);
    // CTAs resident at once: the single-query streaming kernel fits 2 per SM, everything else 1 per SM
    const int64_t resident = (p->stream && p->qt == kQT ? 1 : 2) * (int64_t)sm_count();
    const int64_t want = std::max<int64_t>(1, ceil_div(resident, p->tiles));
    const int64_t groups = ceil_div(std::max<int64_t>(n, 1), 32);
    p->groups_per_slice = round_up(ceil_div(groups, want), 8);          // whole CTA rounds (8 warps x 1 group)
    p->slices = (int)std::max<int64_t>(1, ceil_div(groups, p->groups_per_slice));
    p->warm_rounds = (int)std::max<int64_t>(1, std::min<int64_t>(3, ceil_div(kPrepassSampleRows, (int64_t)p->slices * kTT)));
    p->pre_slices = 0;
    p->pre_gstride = 1;
    p->pre_groups_per_slice = 0;
    if (p->stream && n >= kPrepassMinRows && k <= kPrepassSampleRows / 64) {
        p->pre_gstride = std::max<int64_t>(2, n / kPrepassSampleRows);
        const int64_t pre_groups = ceil_div(n, 32 * p->pre_gstride);
        // 512-row slices (two CTA rounds, one 512-key sort each; >= 2k rows so every list is full)
        p->pre_groups_per_slice = std::max<int64_t>(16, round_up(ceil_div(2 * (int64_t)k, 32), 8));
        p->pre_slices = (int)ceil_div(pre_groups, p->pre_groups_per_slice);
    }
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
    if (n_queries <= 0 || k <= 0) return 256;
    KnnPlan p;
    make_plan(n, d, n_queries, k, &p);
    int64_t bytes = round_up((int64_t)p.slices * n_queries * (int64_t)p.kp * 8, 256);
    bytes += round_up((int64_t)std::max(p.pre_slices, 1) * n_queries * (int64_t)p.kp * 8, 256);
    bytes += round_up((int64_t)n_queries * 8, 256);
    bytes += round_up((int64_t)p.slices * n_queries * 32 * 8, 256);       // per-CTA probe lists of the single-launch mode
    return bytes + 256;
}

static int launch_merge_select(const u64 *lists, int n_lists, int nq, int kp, int k, float *dist_out, int64_t *idx_out,
                               int64_t idx_base, u64 *tau_out, cudaStream_t st) {
    const int n_probe_sort = std::max(32, next_pow2(n_lists));
    DCNR_REQUIRE(n_probe_sort <= 16384, "top-k merge fan-in too large (%d lists)", n_lists);
    const size_t smem = (size_t)(kCandCap + n_probe_sort) * 8;
    DCNR_CUDA_CHECK(cudaFuncSetAttribute(k_knn_merge_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_knn_merge_select<<<(unsigned)nq, kTT, smem, st>>>(lists, n_lists, nq, kp, k, n_probe_sort, dist_out, idx_out, idx_base,
                                                       tau_out);
    DCNR_LAUNCHED();
    return DCNR_OK;
}

struct FusedOut {          // single-launch mode: probe scratch and the final outputs (NULL probes = plain scan)
    u64 *probes;
    float *dist_out;
    int64_t *idx_out;
    int64_t idx_base;
};

// *fused_ok (when non-NULL) reports whether the cooperative single-launch form was used; if the grid cannot be
// co-resident nothing is launched and the caller takes the multi-kernel path.
template <int DV, int QT>
static int launch_stream(const float *cat, int64_t n, const float *q, int nq, int k, const KnnPlan &p, int slices,
                         int64_t groups_per_slice, int64_t gstride, const u64 *tau0, u64 *out, cudaStream_t st,
                         const FusedOut *fo, bool *fused_ok) {
    constexpr int d = 4 * DV;
    const size_t smem = stream_smem(d, QT);
    DCNR_CUDA_CHECK(cudaFuncSetAttribute(k_knn_stream<DV, QT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)slices, (unsigned)p.tiles);
    int n_stages = stream_stages(d, QT);
    int fused = 0, warm = p.warm_rounds, n_probe_sort = std::max(32, next_pow2(slices));
    u64 *probes = nullptr;
    float *dist_out = nullptr;
    int64_t *idx_out = nullptr, idx_base = 0;
    if (fo != nullptr) {
        int per_sm = 0;
        DCNR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_knn_stream<DV, QT>, kTT, smem));
        const bool fits = (int64_t)per_sm * sm_count() >= (int64_t)slices * p.tiles && n_probe_sort <= (QT == 1 ? kCap2 : kCap8) / 2 &&
                          k <= (QT == 1 ? kCap2 : kCap8) / 2;
        if (fused_ok != nullptr) *fused_ok = fits;
        if (!fits) return DCNR_OK;
        fused = 1; probes = fo->probes; dist_out = fo->dist_out; idx_out = fo->idx_out; idx_base = fo->idx_base;
        void *args[] = {&cat, &n, &q, &nq, &k, const_cast<int *>(&p.kp), &groups_per_slice, &gstride, &n_stages, &tau0, &out,
                        &fused, &warm, &probes, &n_probe_sort, &dist_out, &idx_out, &idx_base};
        DCNR_CUDA_CHECK(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(k_knn_stream<DV, QT>), grid, dim3(kTT), args, smem, st));
        DCNR_LAUNCHED();
        return DCNR_OK;
    }
    k_knn_stream<DV, QT><<<grid, kTT, smem, st>>>(cat, n, q, nq, k, p.kp, groups_per_slice, gstride, n_stages, tau0, out, fused,
                                                  warm, probes, n_probe_sort, dist_out, idx_out, idx_base);
    DCNR_LAUNCHED();
    return DCNR_OK;
}

template <int DV>
static int launch_stream_q(const float *cat, int64_t n, const float *q, int nq, int k, const KnnPlan &p, int slices,
                           int64_t groups_per_slice, int64_t gstride, const u64 *tau0, u64 *out, cudaStream_t st,
                           const FusedOut *fo, bool *fused_ok) {
    if (p.qt == 1)
        return launch_stream<DV, 1>(cat, n, q, nq, k, p, slices, groups_per_slice, gstride, tau0, out, st, fo, fused_ok);
    return launch_stream<DV, kQT>(cat, n, q, nq, k, p, slices, groups_per_slice, gstride, tau0, out, st, fo, fused_ok);
}

static int launch_stream_any(int d, const float *cat, int64_t n, const float *q, int nq, int k, const KnnPlan &p, int slices,
                             int64_t groups_per_slice, int64_t gstride, const u64 *tau0, u64 *out, cudaStream_t st,
                             const FusedOut *fo = nullptr, bool *fused_ok = nullptr) {
    switch (d) {
        case 16: return launch_stream_q<4>(cat, n, q, nq, k, p, slices, groups_per_slice, gstride, tau0, out, st, fo, fused_ok);
        case 24: return launch_stream_q<6>(cat, n, q, nq, k, p, slices, groups_per_slice, gstride, tau0, out, st, fo, fused_ok);
        case 32: return launch_stream_q<8>(cat, n, q, nq, k, p, slices, groups_per_slice, gstride, tau0, out, st, fo, fused_ok);
        case 48: return launch_stream_q<12>(cat, n, q, nq, k, p, slices, groups_per_slice, gstride, tau0, out, st, fo, fused_ok);
        default: return launch_stream_q<16>(cat, n, q, nq, k, p, slices, groups_per_slice, gstride, tau0, out, st, fo, fused_ok);
    }
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
    make_plan(n, d, n_queries, k, &p);
    Arena ar(scratch, scratch_bytes);
    u64 *lists = ar.take<u64>((int64_t)p.slices * n_queries * p.kp);
    u64 *pre_lists = ar.take<u64>((int64_t)std::max(p.pre_slices, 1) * n_queries * p.kp);
    u64 *tau0 = ar.take<u64>(n_queries);
    u64 *probes = ar.take<u64>((int64_t)p.slices * n_queries * 32);

    if (p.stream) {
        // single cooperative launch (scan + bound exchange + merge) whenever the grid can be co-resident
        if (p.qt == 1) {         // (8-query tiles: the separate sampled pre-pass measured faster, 0.53 vs 0.65 ms)
            FusedOut fo{probes, dist_out, idx_out, idx_base};
            bool fused_ok = false;
            DCNR_TRY(launch_stream_any(d, catalog_hat, n, queries_hat, n_queries, k, p, p.slices, p.groups_per_slice, 1, nullptr,
                                       lists, st, &fo, &fused_ok));
            if (fused_ok) return DCNR_OK;
        }
        const u64 *t0 = nullptr;
        if (p.pre_slices > 0) {      // sampled pre-pass: k-th best key of ~64 K rows -> initial bound of the full scan
            DCNR_TRY(launch_stream_any(d, catalog_hat, n, queries_hat, n_queries, k, p, p.pre_slices, p.pre_groups_per_slice,
                                       p.pre_gstride, nullptr, pre_lists, st));
            DCNR_TRY(launch_merge_select(pre_lists, p.pre_slices, n_queries, p.kp, k, nullptr, nullptr, 0, tau0, st));
            t0 = tau0;
        }
        DCNR_TRY(launch_stream_any(d, catalog_hat, n, queries_hat, n_queries, k, p, p.slices, p.groups_per_slice, 1, t0, lists,
                                   st));
    } else {
        const size_t smem = (size_t)kQT * kCap * 8 + kQT * 8 + kQT * 4 + (size_t)kQT * d * 4;
        if (smem > 48 * 1024)
            DCNR_CUDA_CHECK(cudaFuncSetAttribute(k_knn_scan<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid((unsigned)p.slices, (unsigned)p.tiles);
        k_knn_scan<0><<<grid, kTT, smem, st>>>(catalog_hat, n, d, queries_hat, n_queries, k, p.kp, p.groups_per_slice * 32,
                                               lists);
        DCNR_LAUNCHED();
    }
    return launch_merge_select(lists, p.slices, n_queries, p.kp, k, dist_out, idx_out, idx_base, nullptr, st);
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
