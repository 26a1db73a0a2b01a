// K9 for query BATCHES: cosine top-k with a tensor-core shortlist and an exact fp32 re-score (sm_100a).
//
// Replaces NearestNeighbors.kneighbors (main.py:200, :300) for many queries at once -- the batched candidate generation of
// SURVEY 8f-2 (one query per positive hotel, main.py:196-203) and cfg4's Q = 32 / 1024.  Results are the ones of
// oracle/knn_oracle.c bit for bit (sequential-fma fp32 scores, dist = clip(1 - sim, 0, 2), order (dist, index) ascending):
// the tensor core only decides WHICH rows get the exact treatment.
//
//   scan    k_knn_tc_scan: persistent, one CTA per SM.  A 128-row catalog tile lands by TMA (fp32, K-major, 32- / 64- / 128-byte
//           swizzle), the queries (up to 1 024, resident in shared memory as the B operand) are multiplied block by block
//           with tcgen05.mma kind::tf32 (M = 128 rows, N <= 256 queries, K = 8 per instruction) into a double-buffered
//           TMEM accumulator.  Eight warps read the scores back (tcgen05.ld, 32 queries per instruction) and compare them
//           with a per-query threshold; the rare survivors (row, query) go through a shared-memory ring to a flusher warp
//           that appends them to per-query lists in global memory.  The catalog is read once for all queries; the scan is
//           bound by the read-back of 128 x Q scores per tile, not by the MMA (64 clk per 128 x 256 x 16 block) nor by HBM.
//   select  k_knn_tc_select: one CTA per query re-scores its list exactly, sorts the keys (dist bits << 32 | row) and emits
//           either the top-k or the next threshold.
//
// Threshold and proof of completeness.  kind::tf32 reads fp32 operands with the low 13 mantissa bits ignored, so for unit
// vectors |s~ - q.e| <= 2^-9 (1 + 2^-11) |q||e|; the exact path's own rounding is <= d 2^-24.  With eps = 2.0e-3 every row whose
// exact dist is <= tau has s~ >= 1 - tau - eps.  tau is the k-th smallest EXACT dist over a subset of the rows, hence an upper
// bound of the k-th smallest over all rows: no member of the true top-k is dropped.  Levels: (0) 8 192 strided rows scored
// exactly by one CTA per query -> tau0; (1) tensor-core scan of every s-th tile (~ n / 32 rows) with tau0 -> exact top-k of
// that sample -> tau1; (2) scan of all tiles with tau1 -> ~ k s rows per query -> exact top-k.  A list that overflows its
// capacity (pathological duplicates) sets *status; the caller then uses the exact streaming kernels (dcnr_knn_topk).
#include <cuda.h>

#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace dcnr {
namespace kt {

using namespace ptx;
typedef unsigned long long u64;

constexpr int BM = 128;                 // catalog rows per tile
constexpr int NB = 256;                 // queries per MMA block (accumulator columns)
constexpr int kSelWarps = 16;           // score read-back warps: four per TMEM lane quadrant
constexpr int kThreads = 32 * (3 + kSelWarps);      // warp 0 TMA, warp 1 MMA, warps 2-17 score read-back, warp 18 flusher
constexpr int kMaxBuf = 8;              // accumulator slots in tensor memory (512 columns / slot width)
constexpr int kRing = 4096;             // shared-memory ring of (query, row) survivors
constexpr int kSampleRows = 8192;       // level 0
constexpr int kListCap = 16384;         // survivors per query and level
constexpr float kEps = 2.0e-3f;
constexpr u64 kEmpty = ~0ull;
constexpr u64 kMaxKey = ~0ull;
constexpr int kMaxQueries = 1024;       // per launch (shared memory: d 16 -> 1 024, d 32 -> 512, d 64 -> 256)

struct Params {
    int64_t n_rows, n_tiles, tile_stride;       // tile t covers rows [t * tile_stride * 128 * sub, + 128 * sub)
    int32_t d, kb_floats, nkb, nq, nqb, stages;
    int32_t slot_cols, nbuf;                    // accumulator slot width (32 .. 256 columns per sub-tile) and count (2 .. kMaxBuf)
    int32_t sub;                                // 128-row sub-tiles per tile (4 / 2 for <= 32 / <= 64 queries: one barrier round
                                                // then covers 512 / 256 rows and every read-back warp has work on every tile)
    const float *thr;                           // [nq] keep a row when its approximate score is >= thr
    uint32_t *lists;                            // [nq][kListCap]
    int32_t *counts;                            // [nq]
};

__device__ __forceinline__ float exact_dist(const float *__restrict__ row, const float *__restrict__ q, int dv) {
    float sim = 0.f;
    const float4 *rp = reinterpret_cast<const float4 *>(row);
    for (int j = 0; j < dv; ++j) {
        const float4 e = __ldg(rp + j);
        const float4 qv = *reinterpret_cast<const float4 *>(q + 4 * j);
        sim = __fmaf_rn(qv.x, e.x, sim);
        sim = __fmaf_rn(qv.y, e.y, sim);
        sim = __fmaf_rn(qv.z, e.z, sim);
        sim = __fmaf_rn(qv.w, e.w, sim);
    }
    const float dist = __fsub_rn(1.0f, sim);
    return fminf(fmaxf(dist, 0.f), 2.f);
}

__device__ __forceinline__ void sort_keys(u64 *a, int n, int tid, int nthreads) {
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


// The k smallest of n keys (dist bits << 32 | row), sorted, without sorting everything: one histogram over the RANGE of the
// dist bits (2 048 linear bins between the smallest and the largest: the wanted keys sit in the sparse lower tail), the keys of
// the bins up to the one holding the k-th are collected (k + a few) and bitonic-sorted.  Degenerate inputs (massive ties: more
// than kSmall keys collected) take the full sort.  Returns the array whose first min(n, k) entries are the answer.
constexpr int kBins = 2048, kSmall = 2048;
struct SelectScratch {
    u64 small[kSmall];
    uint32_t hist[kBins];
    uint32_t red[64];
    uint32_t ctl[4];          // lo, shift, b*, collected
};
__device__ u64 *block_smallest_k(u64 *keys, int n, int n_pad, int k, SelectScratch *ss, int tid, int nthreads) {
    // n_pad: power of two >= n with keys[n..n_pad) = kMaxKey (for the fallback sort)
    uint32_t lo = 0xffffffffu, hi = 0u;
    for (int i = tid; i < n; i += nthreads) {
        const uint32_t b = (uint32_t)(keys[i] >> 32);
        lo = min(lo, b);
        hi = max(hi, b);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((tid & 31) == 0) { ss->red[tid >> 5] = lo; ss->red[32 + (tid >> 5)] = hi; }
    for (int i = tid; i < kBins; i += nthreads) ss->hist[i] = 0;
    __syncthreads();
    if (tid == 0) {
        const int nw = (nthreads + 31) >> 5;
        uint32_t l = 0xffffffffu, h = 0u;
        for (int w = 0; w < nw; ++w) { l = min(l, ss->red[w]); h = max(h, ss->red[32 + w]); }
        const uint32_t range = n > 0 ? h - l : 0u;
        int shift = 0;
        while ((range >> shift) >= (uint32_t)kBins) ++shift;
        ss->ctl[0] = l; ss->ctl[1] = (uint32_t)shift; ss->ctl[3] = 0;
    }
    __syncthreads();
    const uint32_t base = ss->ctl[0], shift = ss->ctl[1];
    for (int i = tid; i < n; i += nthreads) atomicAdd(&ss->hist[((uint32_t)(keys[i] >> 32) - base) >> shift], 1u);
    __syncthreads();
    if (tid < 32) {                                   // bin holding the k-th smallest: 64 bins per lane, warp scan of the lane sums
        uint32_t sum = 0;
        for (int j = 0; j < kBins / 32; ++j) sum += ss->hist[tid * (kBins / 32) + j];
        uint32_t incl = sum;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (tid >= o) incl += v;
        }
        const uint32_t want = (uint32_t)min(k, n);
        const uint32_t before = incl - sum;
        if (want > 0 && before < want && incl >= want) {
            uint32_t c = before;
            for (int j = 0; j < kBins / 32; ++j) {
                c += ss->hist[tid * (kBins / 32) + j];
                if (c >= want) { ss->ctl[2] = (uint32_t)(tid * (kBins / 32) + j); break; }
            }
        }
        if (want == 0 && tid == 0) ss->ctl[2] = 0;
    }
    __syncthreads();
    const uint32_t bstar = ss->ctl[2];
    for (int i = tid; i < n; i += nthreads) {
        const u64 key = keys[i];
        if ((((uint32_t)(key >> 32) - base) >> shift) <= bstar) {
            const uint32_t pos = atomicAdd(&ss->ctl[3], 1u);
            if (pos < (uint32_t)kSmall) ss->small[pos] = key;
        }
    }
    __syncthreads();
    const int m = (int)ss->ctl[3];
    if (m > kSmall) {                                 // massive ties: sort everything
        sort_keys(keys, n_pad, tid, nthreads);
        return keys;
    }
    int m_pad = 32;
    while (m_pad < m) m_pad <<= 1;
    m_pad = max(m_pad, 256);                          // k <= 256 entries are always readable
    for (int i = m + tid; i < m_pad; i += nthreads) ss->small[i] = kMaxKey;
    __syncthreads();
    sort_keys(ss->small, m_pad, tid, nthreads);
    return ss->small;
}

// ---- level 0: exact k-th smallest dist over kSampleRows strided rows, one CTA per query ----
__global__ void __launch_bounds__(1024)
k_knn_tc_tau0(const float *__restrict__ cat, int64_t n, int d, const float *__restrict__ queries, int k, float *__restrict__ thr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *keys = reinterpret_cast<u64 *>(smem_raw);                       // [kSampleRows]
    SelectScratch *ss = reinterpret_cast<SelectScratch *>(keys + kSampleRows);
    float *sq = reinterpret_cast<float *>(ss + 1);                       // [d]
    const int q = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < d; i += blockDim.x) sq[i] = queries[(int64_t)q * d + i];
    __syncthreads();
    for (int i = tid; i < kSampleRows; i += blockDim.x) {
        const int64_t row = ((int64_t)i * n) / kSampleRows;
        keys[i] = ((u64)__float_as_uint(exact_dist(cat + row * d, sq, d >> 2)) << 32) | (u64)(uint32_t)row;
    }
    __syncthreads();
    const u64 *best = block_smallest_k(keys, kSampleRows, kSampleRows, k, ss, tid, blockDim.x);
    if (tid == 0) thr[q] = (1.0f - __uint_as_float((uint32_t)(best[k - 1] >> 32))) - kEps;
}

// ---- the tensor-core scan ----
__global__ void __launch_bounds__(kThreads, 1)
k_knn_tc_scan(const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmQ, Params p) {
    extern __shared__ __align__(1024) uint8_t smem_scan[];
    uint8_t *smem = smem_scan + ((1024u - (smem_u32(smem_scan) & 1023u)) & 1023u);
    const int kbb = p.kb_floats * 4;                                   // row bytes of a K block: 32, 64 or 128
    const int q_sub = NB * kbb, c_sub = BM * kbb;                      // one [256 x KB] query / [128 x KB] catalog sub-tile
    uint8_t *qs = smem;                                                // [nqb][nkb][NB x KB]
    uint8_t *cs = qs + (size_t)p.nqb * p.nkb * q_sub;                  // [stages][sub][nkb][BM x KB]
    // The thresholds ride in the MMA: one extra K = 8 step multiplies a constant A block (columns 0, 1 = 1) with a per-query B block
    // (columns 0, 1 = -thr_hi, -thr_lo; thr = hi + lo, both exact in tf32), so the accumulator holds s~ - thr and "keep this
    // row" is its sign bit.  Both blocks are K-major with 32-byte rows (SWIZZLE_32B: the 16-byte half of a row is XOR-ed with
    // bit 2 of the row index).
    uint8_t *tb = cs + (size_t)p.stages * p.sub * p.nkb * c_sub;       // [nqb][NB x 32 B]
    uint8_t *ta = tb + (size_t)p.nqb * NB * 32;                        // [BM x 32 B]
    u64 *ring = reinterpret_cast<u64 *>(ta + BM * 32);                 // [kRing]
    uint64_t *bars = reinterpret_cast<uint64_t *>(ring + kRing);       // full[stages], empty[stages], accfull[kMaxBuf], accfree[kMaxBuf], qfull
    uint32_t *ctl = reinterpret_cast<uint32_t *>(bars + 2 * p.stages + 2 * kMaxBuf + 1);        // head, tail, done, tmem slot
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * p.stages, accfull0 = empty0 + 8 * p.stages,
                   accfree0 = accfull0 + 8 * kMaxBuf, qfull = accfree0 + 8 * kMaxBuf;
    volatile uint32_t *head = ctl, *tail = ctl + 1, *done = ctl + 2;
    uint32_t *tmem_slot = ctl + 3;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < kRing; i += kThreads) ring[i] = kEmpty;
    for (int i = threadIdx.x; i < p.nqb * NB; i += kThreads) {         // row i of the threshold blocks (query i)
        float hi = __int_as_float(0xff800000), lo = 0.f;               // no query: -inf, the sign bit is always set
        if (i < p.nq) {
            const float t = -p.thr[i];
            hi = __uint_as_float(__float_as_uint(t) & 0xffffe000u);
            lo = t - hi;
        }
        float4 *row = reinterpret_cast<float4 *>(tb + (size_t)i * 32);
        const int sw = (i >> 2) & 1;
        row[sw] = make_float4(hi, lo, 0.f, 0.f);
        row[sw ^ 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = threadIdx.x; i < BM; i += kThreads) {
        float4 *row = reinterpret_cast<float4 *>(ta + (size_t)i * 32);
        const int sw = (i >> 2) & 1;
        row[sw] = make_float4(1.f, 1.f, 0.f, 0.f);
        row[sw ^ 1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    fence_proxy_async_smem();                                          // generic-proxy writes -> visible to the tensor core
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int b = 0; b < kMaxBuf; ++b) {
            mbar_init(accfull0 + 8 * b, 1);
            mbar_init(accfree0 + 8 * b, kSelWarps);
        }
        mbar_init(qfull, 1);
        ctl[0] = ctl[1] = ctl[2] = 0;
        mbar_init_fence();
    }
    if (warp == 1) tmem_alloc<1>(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int64_t my_tiles = p.n_tiles > (int64_t)blockIdx.x ? (p.n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    if (warp == 0) {
        // ---------------- TMA producer: the queries once, then this CTA's catalog tiles ----------------
        if (elect_one()) {
            mbar_expect_tx(qfull, (uint32_t)(p.nqb * p.nkb * q_sub));
            for (int qb = 0; qb < p.nqb; ++qb)
                for (int kb = 0; kb < p.nkb; ++kb)
                    tma_load_2d(smem_u32(qs + (size_t)(qb * p.nkb + kb) * q_sub), &tmQ, kb * p.kb_floats, qb * NB, qfull);
        }
        __syncwarp();
        int s = 0; uint32_t ph = 0;
        for (int64_t i = 0; i < my_tiles; ++i) {
            const int64_t t = blockIdx.x + i * gridDim.x;
            mbar_wait(empty0 + 8 * s, ph ^ 1);
            if (elect_one()) {
                mbar_expect_tx(full0 + 8 * s, (uint32_t)(p.sub * p.nkb * c_sub));
                for (int j = 0; j < p.sub; ++j)
                    for (int kb = 0; kb < p.nkb; ++kb)
                        tma_load_2d(smem_u32(cs + (size_t)((s * p.sub + j) * p.nkb + kb) * c_sub), &tmC, kb * p.kb_floats,
                                    (int)((t * p.tile_stride * p.sub + j) * BM), full0 + 8 * s);
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: every query block against the landed tile ----------------
        mbar_wait(qfull, 0);
        int s = 0; uint32_t ph = 0, buf = 0, aph = 0;                   // aph: one phase bit per accumulator slot
        for (int64_t i = 0; i < my_tiles; ++i) {
            mbar_wait(full0 + 8 * s, ph);
            tc_fence_after();
            for (int qb = 0; qb < p.nqb; ++qb) {
                mbar_wait(accfree0 + 8 * buf, ((aph >> buf) & 1u) ^ 1u);
                tc_fence_after();
                if (elect_one()) {
                    const int ncols = min(NB, ((p.nq - qb * NB + 15) >> 4) << 4);
                    const uint32_t idesc = idesc_tf32(BM, ncols);
                    const uint64_t dta = smem_desc_kmajor(smem_u32(ta), 32);
                    const uint64_t dtb = smem_desc_kmajor(smem_u32(tb + (size_t)qb * NB * 32), 32);
                    for (int j = 0; j < p.sub; ++j) {
                        mma_tf32_ss(tmem_base + (buf * p.sub + j) * p.slot_cols, dta, dtb, idesc, 0u);      // acc = -thr
                        for (int kb = 0; kb < p.nkb; ++kb) {
                            const uint64_t da = smem_desc_kmajor(smem_u32(cs + (size_t)((s * p.sub + j) * p.nkb + kb) * c_sub), kbb);
                            const uint64_t db = smem_desc_kmajor(smem_u32(qs + (size_t)(qb * p.nkb + kb) * q_sub), kbb);
                            for (int ks = 0; ks < p.kb_floats / 8; ++ks)
                                mma_tf32_ss(tmem_base + (buf * p.sub + j) * p.slot_cols, da + (uint64_t)(ks * 2),
                                            db + (uint64_t)(ks * 2), idesc, 1u);
                        }
                    }
                    mma_commit<1>(accfull0 + 8 * buf);
                    if (qb == p.nqb - 1) mma_commit<1>(empty0 + 8 * s);
                }
                __syncwarp();
                aph ^= 1u << buf;
                if (++buf == (uint32_t)p.nbuf) buf = 0;
            }
            if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
    } else if (warp < 2 + kSelWarps) {
        // ---------------- score read-back: lane = catalog row, 32 queries per tcgen05.ld ----------------
        // work item = (tile, query block, sub-tile, group of 32 queries); items are dealt round-robin to the four warps of a
        // quadrant
        const int quad = warp & 3, part = (warp - 2) >> 2, nparts = kSelWarps / 4;
        uint32_t buf = 0, fph = 0, item = 0;
        for (int64_t i = 0; i < my_tiles; ++i) {
            const int64_t t = blockIdx.x + i * gridDim.x;
            const int64_t row0 = t * p.tile_stride * p.sub * BM + quad * 32 + lane;
            for (int qb = 0; qb < p.nqb; ++qb) {
                mbar_wait(accfull0 + 8 * buf, (fph >> buf) & 1u);
                tc_fence_after();
                const int ncols = min(NB, p.nq - qb * NB);
                const int groups = (ncols + 31) >> 5;
                for (int w = 0; w < p.sub * groups; ++w) {
                    if ((int)((item + w) % nparts) != part) continue;
                    const int st = w / groups, g = w - st * groups;           // sub-tile, group of 32 queries
                    const int64_t row = row0 + (int64_t)st * BM;
                    const bool valid = row < p.n_rows;
                    uint32_t r[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (buf * p.sub + st) * p.slot_cols + g * 32, r);
                    // the accumulator holds s~ - thr: a row survives for query j iff the sign bit of r[j] is clear.  Fast
                    // path: AND of all 32 words (16 LOP3) -- sign bit still set means no survivor in this group
                    uint32_t all = r[0];
#pragma unroll
                    for (int j = 1; j < 32; ++j) all &= r[j];
                    uint32_t mask = 0;
                    if ((int32_t)all >= 0) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) mask |= ((~r[j]) >> 31) << j;
                    }
                    const int nvalid = ncols - g * 32;                        // queries of this group that exist (columns past the
                    if (nvalid < 32) mask &= (1u << nvalid) - 1u;             // MMA's N hold stale tensor-memory contents)
                    if (!valid) mask = 0;
                    while (mask != 0) {
                        const int j = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const uint32_t pos = atomicAdd(const_cast<uint32_t *>(head), 1u);
                        uint32_t spins = 0;
                        while (pos - *tail >= (uint32_t)kRing) {
                            __nanosleep(64);
                            if (++spins > (1u << 22)) __trap();
                        }
                        *reinterpret_cast<volatile u64 *>(ring + (pos & (kRing - 1))) =
                            ((u64)(uint32_t)(qb * NB + g * 32 + j) << 32) | (u64)(uint32_t)row;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(accfree0 + 8 * buf);
                item += (uint32_t)(p.sub * groups);
                fph ^= 1u << buf;
                if (++buf == (uint32_t)p.nbuf) buf = 0;
            }
        }
        __syncwarp();
        __threadfence_block();
        if (lane == 0) atomicAdd(const_cast<uint32_t *>(done), 1u);
    } else {
        // ---------------- flusher: shared-memory ring -> per-query lists in global memory ----------------
        uint32_t t0 = 0, idle = 0;
        for (;;) {
            const uint32_t h = *head;
            uint32_t avail = h - t0;
            if (avail == 0) {
                if (*done == (uint32_t)kSelWarps && *head == t0) break;
                __nanosleep(100);
                if (++idle > (1u << 26)) __trap();
                continue;
            }
            idle = 0;
            const uint32_t nnow = min(avail, 256u);
            u64 e[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                e[u] = kEmpty;
                const uint32_t off = u * 32 + lane;
                if (off < nnow) {
                    volatile u64 *slot = ring + ((t0 + off) & (kRing - 1));
                    uint32_t spins = 0;
                    while ((e[u] = *slot) == kEmpty)
                        if (++spins > (1u << 26)) __trap();
                    *slot = kEmpty;
                }
            }
            int32_t pos[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                pos[u] = e[u] != kEmpty ? atomicAdd(p.counts + (uint32_t)(e[u] >> 32), 1) : 0;
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (e[u] != kEmpty && pos[u] < kListCap) p.lists[(size_t)(uint32_t)(e[u] >> 32) * kListCap + pos[u]] = (uint32_t)e[u];
            t0 += nnow;
            __syncwarp();
            __threadfence_block();
            if (lane == 0) *tail = t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<1>(tmem_base, 512);
}

// ---- exact re-score of one query's list, sort, emit ----
// mode 0: thr[q] for the next level; mode 1: the top-k (dist_out / idx_out [nq, k]).  status: bit 0 = a list overflowed.
__global__ void __launch_bounds__(1024)
k_knn_tc_select(const float *__restrict__ cat, int d, const float *__restrict__ queries, int k, const uint32_t *__restrict__ lists,
                const int32_t *__restrict__ counts, int mode, float *__restrict__ thr, float *__restrict__ dist_out,
                int64_t *__restrict__ idx_out, int64_t idx_base, int32_t *__restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *keys = reinterpret_cast<u64 *>(smem_raw);                       // [kListCap]
    SelectScratch *ss = reinterpret_cast<SelectScratch *>(keys + kListCap);
    float *sq = reinterpret_cast<float *>(ss + 1);
    const int q = blockIdx.x, tid = threadIdx.x;
    const int raw = counts[q];
    const int cnt = min(raw, kListCap);
    if (tid == 0 && (raw > kListCap || cnt < k) && status != nullptr) atomicOr(status, raw > kListCap ? 1 : 2);
    for (int i = tid; i < d; i += blockDim.x) sq[i] = queries[(int64_t)q * d + i];
    __syncthreads();
    int n_sort = 32;
    while (n_sort < cnt) n_sort <<= 1;
    n_sort = max(n_sort, 256);
    for (int i = tid; i < n_sort; i += blockDim.x) {
        u64 key = kMaxKey;
        if (i < cnt) {
            const uint32_t row = lists[(size_t)q * kListCap + i];
            key = ((u64)__float_as_uint(exact_dist(cat + (int64_t)row * d, sq, d >> 2)) << 32) | (u64)row;
        }
        keys[i] = key;
    }
    __syncthreads();
    const u64 *best = block_smallest_k(keys, cnt, n_sort, k, ss, tid, blockDim.x);
    if (mode == 0) {
        if (tid == 0) thr[q] = cnt >= k ? (1.0f - __uint_as_float((uint32_t)(best[k - 1] >> 32))) - kEps : -4.0f;
    } else {
        for (int j = tid; j < k; j += blockDim.x) {
            const u64 key = j < cnt ? best[j] : kMaxKey;
            const bool ok = key != kMaxKey;
            dist_out[(int64_t)q * k + j] = ok ? __uint_as_float((uint32_t)(key >> 32)) : __int_as_float(0x7f800000);
            idx_out[(int64_t)q * k + j] = ok ? idx_base + (int64_t)(uint32_t)key : -1;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
// [rows, d] fp32 row-major, box {kb_floats, box_rows}, swizzle = the box's row bytes
static int make_map(CUtensorMap *tm, const float *base, int64_t rows, int d, int kb_floats, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return DCNR_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)d * 4};
    cuuint32_t box[2] = {(cuuint32_t)kb_floats, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE,
                    kb_floats == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : (kb_floats == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B),
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): rows %lld d %d", (int)r, (long long)rows, d);
        return DCNR_ERR_CUDA;
    }
    return DCNR_OK;
}

// queries resident in shared memory per launch (64-96 KB of operand blocks)
static int queries_per_launch(int d) { return d <= 24 ? 1024 : (d <= 48 ? 512 : 256); }
// K block = the widest of 32 / 16 / 8 floats (128- / 64- / 32-byte swizzle) that divides d: 16 -> 16, 24 -> 8, 32 -> 32, 48 -> 16
static int k_block_floats(int d) { return d % 32 == 0 ? 32 : (d % 16 == 0 ? 16 : 8); }

}  // namespace kt

bool knn_tc_supported(int64_t n, int32_t d, int32_t n_queries, int32_t k) {
    return (d == 16 || d == 24 || d == 32 || d == 48 || d == 64) && n >= (1 << 18) && n <= (1 << 24) && n_queries >= 1 && k >= 1 &&
           k <= 256;
}

int64_t knn_tc_scratch_bytes(int64_t n, int32_t d, int32_t n_queries, int32_t k) {
    if (!knn_tc_supported(n, d, n_queries, k)) return 0;
    const int64_t nq = std::min<int64_t>(n_queries, kt::queries_per_launch(d));
    return round_up(nq * kt::kListCap * 4, 256) + 2 * round_up(nq * 4, 256) + 256;
}

int launch_knn_tc(const float *cat, int64_t n, int32_t d, const float *queries, int32_t n_queries, int32_t k, int64_t idx_base,
                  float *dist_out, int64_t *idx_out, void *scratch, int64_t scratch_bytes, int32_t *status,
                  cudaStream_t st) {
    using namespace kt;
    DCNR_REQUIRE(knn_tc_supported(n, d, n_queries, k), "shape not supported by the tensor-core top-k");
    DCNR_REQUIRE(scratch_bytes >= knn_tc_scratch_bytes(n, d, n_queries, k), "knn scratch too small");
    DCNR_REQUIRE((((uintptr_t)cat | (uintptr_t)queries) & 15) == 0, "catalog and queries must be 16-byte aligned");
    const int qmax = queries_per_launch(d);
    Arena ar(scratch, scratch_bytes);
    uint32_t *lists = ar.take<uint32_t>((int64_t)std::min(n_queries, qmax) * kListCap);
    int32_t *counts = ar.take<int32_t>(std::min(n_queries, qmax));
    float *thr = ar.take<float>(std::min(n_queries, qmax));
    const int kbf = k_block_floats(d), nkb = d / kbf;
    const int64_t tiles_all = ceil_div(n, (int64_t)BM);
    const int64_t stride1 = std::max<int64_t>(32, ceil_div(n, (int64_t)(1 << 18)));
    CUtensorMap tmC;
    DCNR_TRY(make_map(&tmC, cat, n, d, kbf, BM));
    const size_t sel_smem = (size_t)kListCap * 8 + sizeof(SelectScratch) + (size_t)d * 4;
    const size_t tau_smem = (size_t)kSampleRows * 8 + sizeof(SelectScratch) + (size_t)d * 4;
    DCNR_CUDA_CHECK(cudaFuncSetAttribute(k_knn_tc_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sel_smem));
    DCNR_CUDA_CHECK(cudaFuncSetAttribute(k_knn_tc_tau0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tau_smem));
    for (int q0 = 0; q0 < n_queries; q0 += qmax) {
        const int nq = std::min(qmax, n_queries - q0);
        const float *qp = queries + (int64_t)q0 * d;
        CUtensorMap tmQ;
        DCNR_TRY(make_map(&tmQ, qp, nq, d, kbf, NB));
        Params p{};
        p.n_rows = n; p.d = d; p.kb_floats = kbf; p.nkb = nkb; p.nq = nq; p.nqb = (int)ceil_div(nq, NB);
        p.thr = thr; p.lists = lists; p.counts = counts;
        p.slot_cols = nq >= NB ? NB : (int)round_up(nq, 32);
        const size_t fixed = (size_t)p.nqb * nkb * NB * kbf * 4 + (size_t)p.nqb * NB * 32 + (size_t)BM * 32 + (size_t)kRing * 8 + 256 + 1024;
        p.sub = nq <= 32 ? 4 : (nq <= 64 ? 2 : 1);
        while (p.sub > 1 && fixed + (size_t)3 * p.sub * nkb * BM * kbf * 4 > (size_t)200 * 1024) p.sub >>= 1;      // >= 3 stages
        p.nbuf = std::min(kMaxBuf, 512 / (p.slot_cols * p.sub));
        const int64_t stage_bytes = (int64_t)p.sub * nkb * BM * kbf * 4;
        p.stages = (int)std::max<int64_t>(2, std::min<int64_t>(6, ((int64_t)200 * 1024 - (int64_t)fixed) / stage_bytes));
        const size_t smem = fixed + (size_t)p.stages * stage_bytes;
        DCNR_CUDA_CHECK(cudaFuncSetAttribute(k_knn_tc_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_knn_tc_tau0<<<(unsigned)nq, 1024, tau_smem, st>>>(cat, n, d, qp, k, thr);
        DCNR_LAUNCHED();
        for (int level = 1; level <= 2; ++level) {
            p.tile_stride = level == 1 ? stride1 : 1;
            p.n_tiles = ceil_div(ceil_div(tiles_all, (int64_t)p.sub), p.tile_stride);
            DCNR_CUDA_CHECK(cudaMemsetAsync(counts, 0, (size_t)nq * 4, st));
            const unsigned grid = (unsigned)std::min<int64_t>(p.n_tiles, sm_count());
            k_knn_tc_scan<<<grid, kThreads, smem, st>>>(tmC, tmQ, p);
            DCNR_LAUNCHED();
            k_knn_tc_select<<<(unsigned)nq, 1024, sel_smem, st>>>(cat, d, qp, k, lists, counts, level == 2 ? 1 : 0, thr,
                                                                 dist_out + (int64_t)q0 * k, idx_out + (int64_t)q0 * k, idx_base,
                                                                 status);
            DCNR_LAUNCHED();
        }
    }
    return DCNR_OK;
}

}  // namespace dcnr

using namespace dcnr;

extern "C" int dcnr_knn_tc_supported(int64_t n, int32_t d, int32_t n_queries, int32_t k) {
    return knn_tc_supported(n, d, n_queries, k) ? 1 : 0;
}

extern "C" int64_t dcnr_knn_tc_scratch_bytes(int64_t n, int32_t d, int32_t n_queries, int32_t k) {
    return knn_tc_scratch_bytes(n, d, n_queries, k);
}

extern "C" int dcnr_knn_topk_tc(const float *catalog_hat, int64_t n, int32_t d, const float *queries_hat, int32_t n_queries,
                                int32_t k, int64_t idx_base, float *dist_out, int64_t *idx_out, void *scratch,
                                int64_t scratch_bytes, int32_t *status, dcnr_stream_t stream) {
    DCNR_REQUIRE(catalog_hat && queries_hat && dist_out && idx_out && scratch, "null argument");
    if (n_queries <= 0) return DCNR_OK;
    return launch_knn_tc(catalog_hat, n, d, queries_hat, n_queries, k, idx_base, dist_out, idx_out, scratch, scratch_bytes,
                         status, as_stream(stream));
}
