// Stable LSD radix sort of (uint32 key, uint32 value) pairs -- the sort inside the embedding backward (K7): keys are
// (table bit | row id), values the batch positions, and stability is what makes the segment sums run in batch order
// (deterministic, the reference's embedding_dense_backward order; train.py:225).
//
// Hand-written for this path (round 1 called cub::DeviceRadixSort here).  Digits are up to 8 bits wide and spread evenly
// over the passes (21-bit keys of the 1 M x 100 K configuration: three 7-bit passes), each pass three kernels:
//   k_radix_hist     per 4096-key tile, a shared-memory digit histogram (warp-aggregated adds) -> counts[digit][tile]
//   k_radix_scan     exclusive scan of counts in (digit, tile) order, in place (one CTA per 2048 counts + a carry pass)
//   k_radix_scatter  per tile: every warp owns 512 consecutive keys (16 rounds x 32 lanes).  Phase 1 counts digits per warp
//                    with __match_any_sync (no atomics: a digit's leader lane adds the group size to the warp's private
//                    row of counters); a block-wide pass turns the 8 rows into per-warp cursors INSIDE the tile (digit-major);
//                    phase 2 re-walks the keys, which stayed in registers, and drops every pair at its sorted position in a
//                    shared-memory copy of the tile; the tile is then written out linearly, so consecutive threads write
//                    consecutive addresses of the same digit's run (32 keys per digit and tile on average: coalesced).
// Order inside a digit is (tile, warp, round, lane) = input order, i.e. every pass is stable.
#include <cooperative_groups.h>

#include "kernels.cuh"

namespace dcnr {
namespace rs {

constexpr int kThreads = 256, kWarps = 8, kRounds = 16, kTile = kThreads * kRounds;      // 4096 keys per CTA
constexpr int kMaxBits = 8, kMaxRadix = 1 << kMaxBits;

__global__ void __launch_bounds__(kThreads)
k_radix_hist(const uint32_t *__restrict__ keys, int64_t n, int shift, int radix, int64_t n_tiles, uint32_t *__restrict__ counts) {
    extern __shared__ uint32_t sh[];                                  // [radix]
    for (int d = threadIdx.x; d < radix; d += kThreads) sh[d] = 0u;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kTile;
    const uint32_t mask = (uint32_t)radix - 1u;
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll 4
    for (int r = 0; r < kRounds; ++r) {
        const int64_t i = base + (int64_t)(threadIdx.x >> 5) * (kRounds * 32) + r * 32 + lane;
        const bool ok = i < n;
        const uint32_t d = ok ? ((__ldg(keys + i) >> shift) & mask) : (0x80000000u | (uint32_t)lane);
        const uint32_t m = __match_any_sync(0xffffffffu, d);           // one atomic per distinct digit of the warp (hot ids)
        if (ok && (m & lt) == 0u) atomicAdd(&sh[d], (uint32_t)__popc(m));
    }
    __syncthreads();
    for (int d = threadIdx.x; d < radix; d += kThreads) counts[(int64_t)d * n_tiles + blockIdx.x] = sh[d];
}

// exclusive scan of `total` counts in place; chunk c = 2048 consecutive counts per CTA, its sum to sums[c]
__global__ void __launch_bounds__(kThreads)
k_radix_scan_local(uint32_t *__restrict__ counts, int64_t total, uint32_t *__restrict__ sums) {
    __shared__ uint32_t wsum[kWarps];
    const int64_t base = (int64_t)blockIdx.x * (kThreads * 8) + threadIdx.x * 8;
    uint32_t v[8], s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        v[j] = base + j < total ? counts[base + j] : 0u;
        s += v[j];
    }
    uint32_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t woff = 0, all = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
        if (w < (threadIdx.x >> 5)) woff += wsum[w];
        all += wsum[w];
    }
    uint32_t run = woff + incl - s;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (base + j < total) counts[base + j] = run;
        run += v[j];
    }
    if (threadIdx.x == 0) sums[blockIdx.x] = all;
}
// one CTA: exclusive scan of the chunk sums (n_chunks <= 2^20 in practice), in place
__global__ void __launch_bounds__(1024)
k_radix_scan_sums(uint32_t *__restrict__ sums, int64_t n_chunks) {
    __shared__ uint32_t part[1024];
    const int64_t per = (n_chunks + 1023) / 1024;
    const int64_t i0 = min(n_chunks, threadIdx.x * per), i1 = min(n_chunks, i0 + per);
    uint32_t s = 0;
    for (int64_t i = i0; i < i1; ++i) s += sums[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int t = 0; t < 1024; ++t) {
            const uint32_t v = part[t];
            part[t] = run;
            run += v;
        }
    }
    __syncthreads();
    uint32_t run = part[threadIdx.x];
    for (int64_t i = i0; i < i1; ++i) {
        const uint32_t v = sums[i];
        sums[i] = run;
        run += v;
    }
}

// small inputs (total <= 64 Ki counts): the whole exclusive scan in ONE CTA, and the chunk bases it makes redundant zeroed
__global__ void __launch_bounds__(1024)
k_radix_scan_single(uint32_t *__restrict__ counts, int64_t total, uint32_t *__restrict__ sums, int64_t n_chunks) {
    const int64_t per = (total + 1023) / 1024;
    const int64_t i0 = min(total, threadIdx.x * per), i1 = min(total, i0 + per);
    uint32_t s = 0;
    for (int64_t i = i0; i < i1; ++i) s += counts[i];
    // inclusive scan of the 1024 partial sums: warp scans + one pass over the 32 warp totals
    uint32_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += t;
    }
    __shared__ uint32_t wtot[32];
    if ((threadIdx.x & 31) == 31) wtot[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += wtot[w];
    uint32_t run = woff + incl - s;
    for (int64_t i = i0; i < i1; ++i) {
        const uint32_t v = counts[i];
        counts[i] = run;
        run += v;
    }
    for (int64_t c = threadIdx.x; c < n_chunks; c += 1024) sums[c] = 0u;
}

__global__ void __launch_bounds__(kThreads)
k_radix_scatter(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint32_t *__restrict__ keys_out,
                uint32_t *__restrict__ vals_out, int64_t n, int shift, int radix, int64_t n_tiles,
                const uint32_t *__restrict__ counts, const uint32_t *__restrict__ chunk_base) {
    __shared__ uint32_t cur[kWarps * kMaxRadix];                      // per-warp digit counters, then cursors inside the tile
    __shared__ uint32_t lbase[kMaxRadix], gbase[kMaxRadix];           // first sorted position of a digit in the tile / in the output
    __shared__ uint32_t skey[kTile], sval[kTile];                     // the tile in sorted order
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t mask = (uint32_t)radix - 1u, lt = (1u << lane) - 1u;
    for (int d = threadIdx.x; d < kWarps * radix; d += kThreads) cur[d] = 0u;
    __syncthreads();
    const int64_t tile0 = (int64_t)blockIdx.x * kTile;
    const int64_t base = tile0 + (int64_t)warp * (kRounds * 32) + lane;
    const int tile_n = (int)min((int64_t)kTile, n - tile0);
    uint32_t k[kRounds], v[kRounds];
    uint32_t *mine = cur + warp * radix;
    // phase 1: per-warp digit counts (invalid tail lanes get a digit no one shares)
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const int64_t i = base + r * 32;
        const bool ok = i < n;
        k[r] = ok ? __ldg(keys + i) : 0u;
        v[r] = ok ? __ldg(vals + i) : 0u;
        const uint32_t d = ok ? ((k[r] >> shift) & mask) : (0x80000000u | (uint32_t)lane);
        const uint32_t m = __match_any_sync(0xffffffffu, d);
        if (ok && (m & lt) == 0u) mine[d] += (uint32_t)__popc(m);     // the lowest lane of the group: warp-private counter
        __syncwarp();
    }
    __syncthreads();
    // digit totals of the tile -> exclusive scan (radix <= 256 = one digit per thread) -> per-warp cursors
    {
        const int d = threadIdx.x;
        uint32_t tot = 0;
        if (d < radix) {
#pragma unroll
            for (int w = 0; w < kWarps; ++w) tot += cur[w * radix + d];
        }
        uint32_t incl = tot;                                          // block-wide inclusive scan of tot over d
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        __shared__ uint32_t wsum[kWarps];
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w)
            if (w < warp) woff += wsum[w];
        if (d < radix) {
            uint32_t run = woff + incl - tot;
            lbase[d] = run;
            const int64_t ci = (int64_t)d * n_tiles + blockIdx.x;
            gbase[d] = counts[ci] + chunk_base[ci / (kThreads * 8)];
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                const uint32_t c = cur[w * radix + d];
                cur[w * radix + d] = run;
                run += c;
            }
        }
    }
    __syncthreads();
    // phase 2: stable positions inside the tile
#pragma unroll
    for (int r = 0; r < kRounds; ++r) {
        const int64_t i = base + r * 32;
        const bool ok = i < n;
        const uint32_t d = ok ? ((k[r] >> shift) & mask) : (0x80000000u | (uint32_t)lane);
        const uint32_t m = __match_any_sync(0xffffffffu, d);
        uint32_t pos = 0u;
        if (ok) pos = mine[d] + (uint32_t)__popc(m & lt);
        __syncwarp();
        if (ok && (m & lt) == 0u) mine[d] += (uint32_t)__popc(m);
        __syncwarp();
        if (ok) {
            skey[pos] = k[r];
            sval[pos] = v[r];
        }
    }
    __syncthreads();
    // linear write-out: sorted element j of the tile goes to gbase[digit] + (j - lbase[digit])
    for (int j = threadIdx.x; j < tile_n; j += kThreads) {
        const uint32_t key = skey[j];
        const uint32_t d = (key >> shift) & mask;
        const uint32_t pos = gbase[d] + ((uint32_t)j - lbase[d]);
        keys_out[pos] = key;
        vals_out[pos] = sval[j];
    }
}

// Small inputs (<= one tile per SM): ALL passes in ONE cooperative launch.  A pass is count -> publish the tile's digit totals ->
// grid.sync -> every CTA derives its own output bases from the [digit][tile] table (radix threads x n_tiles loads: no separate
// scan) -> rank -> write-out -> grid.sync.  Nine launches (hist, scan, scatter per pass) become one; the training step's
// 131 072 keys sort in ~45 us instead of ~110.  Loads of data written earlier in the same kernel bypass L1 (__ldcg).
__global__ void __launch_bounds__(kThreads)
k_radix_sort_coop(uint32_t *k0, uint32_t *v0, uint32_t *k1, uint32_t *v1, int64_t n, int end_bit, int passes, int n_tiles,
                  uint32_t *counts) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    __shared__ uint32_t cur[kWarps * kMaxRadix];
    __shared__ uint32_t lbase[kMaxRadix], gbase[kMaxRadix];
    __shared__ uint32_t skey[kTile], sval[kTile];
    __shared__ uint32_t wsum[kWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const int64_t tile0 = (int64_t)blockIdx.x * kTile;
    const int64_t base = tile0 + (int64_t)warp * (kRounds * 32) + lane;
    const int tile_n = (int)max((int64_t)0, min((int64_t)kTile, n - tile0));
    const uint32_t *kin = k0, *vin = v0;
    uint32_t *kout = k1, *vout = v1;
    int shift = 0;
    for (int p = 0; p < passes; ++p) {
        const int bits = (end_bit - shift + (passes - p) - 1) / (passes - p);
        const int radix = 1 << bits;
        const uint32_t mask = (uint32_t)radix - 1u;
        for (int d = threadIdx.x; d < kWarps * radix; d += kThreads) cur[d] = 0u;
        __syncthreads();
        uint32_t k[kRounds], v[kRounds];
        uint32_t *mine = cur + warp * radix;
#pragma unroll
        for (int r = 0; r < kRounds; ++r) {
            const int64_t i = base + r * 32;
            const bool ok = i < n;
            k[r] = ok ? __ldcg(kin + i) : 0u;
            v[r] = ok ? __ldcg(vin + i) : 0u;
            const uint32_t d = ok ? ((k[r] >> shift) & mask) : (0x80000000u | (uint32_t)lane);
            const uint32_t m = __match_any_sync(0xffffffffu, d);
            if (ok && (m & lt) == 0u) mine[d] += (uint32_t)__popc(m);
            __syncwarp();
        }
        __syncthreads();
        const int d = threadIdx.x;
        uint32_t tot = 0;
        if (d < radix) {
#pragma unroll
            for (int w = 0; w < kWarps; ++w) tot += cur[w * radix + d];
            counts[(int64_t)d * n_tiles + blockIdx.x] = tot;
        }
        // tile-local exclusive scan of tot over the digits (positions inside the sorted tile)
        uint32_t incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w)
            if (w < warp) woff += wsum[w];
        if (d < radix) {
            uint32_t run = woff + incl - tot;
            lbase[d] = run;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) {
                const uint32_t c = cur[w * radix + d];
                cur[w * radix + d] = run;
                run += c;
            }
        }
        __threadfence();
        grid.sync();                                                   // every tile's digit totals are published
        // global base of digit d for this tile = (keys with a smaller digit, all tiles) + (digit d in earlier tiles)
        uint32_t total = 0, before = 0;
        if (d < radix) {
            for (int t = 0; t < n_tiles; ++t) {
                const uint32_t c = __ldcg(counts + (int64_t)d * n_tiles + t);
                total += c;
                if (t < (int)blockIdx.x) before += c;
            }
        }
        uint32_t gincl = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, gincl, o);
            if (lane >= o) gincl += t;
        }
        __syncthreads();                                               // wsum reuse
        if (lane == 31) wsum[warp] = gincl;
        __syncthreads();
        uint32_t goff = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w)
            if (w < warp) goff += wsum[w];
        if (d < radix) gbase[d] = goff + gincl - total + before;
        __syncthreads();
#pragma unroll
        for (int r = 0; r < kRounds; ++r) {
            const int64_t i = base + r * 32;
            const bool ok = i < n;
            const uint32_t dg = ok ? ((k[r] >> shift) & mask) : (0x80000000u | (uint32_t)lane);
            const uint32_t m = __match_any_sync(0xffffffffu, dg);
            uint32_t pos = 0u;
            if (ok) pos = mine[dg] + (uint32_t)__popc(m & lt);
            __syncwarp();
            if (ok && (m & lt) == 0u) mine[dg] += (uint32_t)__popc(m);
            __syncwarp();
            if (ok) {
                skey[pos] = k[r];
                sval[pos] = v[r];
            }
        }
        __syncthreads();
        for (int j = threadIdx.x; j < tile_n; j += kThreads) {
            const uint32_t key = skey[j];
            const uint32_t dg = (key >> shift) & mask;
            const uint32_t pos = gbase[dg] + ((uint32_t)j - lbase[dg]);
            kout[pos] = key;
            vout[pos] = sval[j];
        }
        shift += bits;
        const uint32_t *tk = kin, *tv = vin;
        kin = kout; vin = vout;
        kout = const_cast<uint32_t *>(tk); vout = const_cast<uint32_t *>(tv);
        if (p + 1 < passes) {
            __threadfence();
            grid.sync();                                               // the pass's output is complete before it is read
        }
    }
}

}  // namespace rs


int64_t radix_sort_scratch_bytes(int64_t n) {
    const int64_t n_tiles = ceil_div(std::max<int64_t>(n, 1), rs::kTile);
    const int64_t total = (int64_t)rs::kMaxRadix * n_tiles;
    return round_up(total * 4, 256) + round_up(ceil_div(total, rs::kThreads * 8) * 4, 256);
}

// Sorts by key bits [0, end_bit).  k0 / v0 hold the input; k1 / v1 are the ping-pong buffers.  *k_out / *v_out point at
// whichever pair holds the result.
int launch_radix_sort_pairs(uint32_t *k0, uint32_t *v0, uint32_t *k1, uint32_t *v1, int64_t n, int end_bit, void *scratch,
                            uint32_t **k_out, uint32_t **v_out, cudaStream_t stream) {
    using namespace rs;
    *k_out = k0;
    *v_out = v0;
    if (n <= 1 || end_bit <= 0) return DCNR_OK;
    const int passes = (end_bit + kMaxBits - 1) / kMaxBits;
    const int64_t n_tiles = ceil_div(n, kTile);
    if (n_tiles <= sm_count()) {                                       // one cooperative launch for all passes
        int per_sm = 0;
        DCNR_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_radix_sort_coop, kThreads, 0));
        if ((int64_t)per_sm * sm_count() >= n_tiles) {
            int nt = (int)n_tiles, eb = end_bit, ps = passes;
            uint32_t *cnt = reinterpret_cast<uint32_t *>(scratch);
            void *args[] = {&k0, &v0, &k1, &v1, &n, &eb, &ps, &nt, &cnt};
            DCNR_CUDA_CHECK(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(k_radix_sort_coop), dim3((unsigned)n_tiles),
                                                        dim3(kThreads), args, 0, stream));
            DCNR_LAUNCHED();
            if (passes & 1) { *k_out = k1; *v_out = v1; }
            return DCNR_OK;
        }
    }
    uint32_t *counts = reinterpret_cast<uint32_t *>(scratch);
    uint32_t *sums = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(scratch) + round_up((int64_t)kMaxRadix * n_tiles * 4, 256));
    int shift = 0;
    for (int p = 0; p < passes; ++p) {
        const int bits = (end_bit - shift + (passes - p) - 1) / (passes - p);      // spread the bits evenly over the passes
        const int radix = 1 << bits;
        const int64_t total = (int64_t)radix * n_tiles, n_chunks = ceil_div(total, kThreads * 8);
        k_radix_hist<<<(unsigned)n_tiles, kThreads, radix * 4, stream>>>(*k_out, n, shift, radix, n_tiles, counts);
        DCNR_LAUNCHED();
        if (total <= 65536) {
            k_radix_scan_single<<<1, 1024, 0, stream>>>(counts, total, sums, n_chunks);
            DCNR_LAUNCHED();
        } else {
            k_radix_scan_local<<<(unsigned)n_chunks, kThreads, 0, stream>>>(counts, total, sums);
            DCNR_LAUNCHED();
            k_radix_scan_sums<<<1, 1024, 0, stream>>>(sums, n_chunks);
            DCNR_LAUNCHED();
        }
        uint32_t *ko = *k_out == k0 ? k1 : k0, *vo = *v_out == v0 ? v1 : v0;
        k_radix_scatter<<<(unsigned)n_tiles, kThreads, 0, stream>>>(*k_out, *v_out, ko, vo, n, shift, radix, n_tiles,
                                                                                   counts, sums);
        DCNR_LAUNCHED();
        *k_out = ko;
        *v_out = vo;
        shift += bits;
    }
    return DCNR_OK;
}

}  // namespace dcnr
