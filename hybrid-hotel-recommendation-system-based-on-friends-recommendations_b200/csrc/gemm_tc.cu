// tcgen05 tensor-core GEMM for the dense layers (sm_100a):  C[M,N] = A[M,K] . B[N,K]^T (+epilogue)
//
//   DCNR_PREC_TF32X3  fp32-parity path: every fp32 operand is split a = hi + lo with hi = rn_tf32(a)
//                     (lo is exact in fp32) and the product is accumulated as hi.hi + lo.hi + hi.lo
//                     in the fp32 TMEM accumulator -- three kind::tf32 MMAs per K step; the dropped
//                     lo.lo term and the truncation of lo are <= 2^-21 relative per product.
//   DCNR_PREC_TF32    single kind::tf32 MMA (hardware truncation to tf32): stated-tolerance fast path.
//
// Structure (one 128 x BLOCK_N output tile per CTA, BLOCK_K = 32 floats = one 128-byte swizzle row):
//   warp 0      TMA producer: cp.async.bulk.tensor loads of the A tile and the (pre-split) B tiles
//               into 128B-swizzled shared memory, completion on the stage's "full" mbarrier
//   warps 2-5   TF32X3 only: split the landed A tile in place into hi / lo (elementwise, so the
//               swizzle does not matter), fence.proxy.async, arrive on "ready"
//   warp 1      one elected thread issues tcgen05.mma (kind::tf32, M=128, N=BLOCK_N, K=8) with the
//               accumulator in TMEM; tcgen05.commit frees the stage ("empty") and finally signals
//               the epilogue ("accum")
//   warps 2-5   epilogue: tcgen05.ld 32 columns at a time, col_scale / bias / residual / relu,
//               128-byte row segments stored with float4
// B (the layer weight) is identical for every M tile, so its hi / lo split is computed once per call
// by k_split_tf32 and both halves are streamed by TMA (they stay L2 resident).
#include <cuda.h>

#include "kernels.cuh"

namespace dcnr {

namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;          // floats; 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 8;            // kind::tf32
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 4;   // 16 KB

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded spin: a pipeline bug traps (CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t spins = 0; !ok; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!ok && spins > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// K-major, SWIZZLE_128B canonical layout: 8-row groups 1024 B apart (SBO), LBO unused (1), version 1.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                             // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                             // layout type SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
        "%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Params {
    int64_t M;
    int32_t N_total, K, block_n, terms, stages, acc_cols, acc_stages, corr_sep, tmem_cols;
    int32_t num_m_tiles, num_n_tiles;
    float *C;                 // may be NULL when only the fused row dot is wanted
    int64_t ldc;
    GemmEpilogue epi;
    const float *dot_w;       // optional fused row dot with the epilogue output: [N_total]
    float *dot_out;           // [num_n_tiles][M] partial dots (summed by the caller)
};

constexpr int kThreadsP = 320;
constexpr int kEpiStageBytes = 4 * 32 * 36 * 4;   // per-warp 32x36 fp32 transpose tiles of the epilogue   // warp 0 TMA, warp 1 MMA, warps 2-5 operand split, warps 6-9 epilogue

// Persistent, warp-specialised: every CTA (one per SM) walks tiles t = blockIdx.x, +gridDim.x, ...
// Three rings run concurrently: shared-memory stages (TMA -> split -> MMA), TMEM accumulator stages
// (MMA -> epilogue) and the tile sequence itself, so the epilogue of tile i overlaps the loads and
// MMAs of tile i+1.
__global__ void __launch_bounds__(kThreadsP, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
          const __grid_constant__ CUtensorMap tmBlo, Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_tile_bytes = p.block_n * BLOCK_K * 4;
    const int stage_bytes = (p.terms == 3 ? 2 : 1) * (A_TILE_BYTES + b_tile_bytes);
    const int stages = p.stages, acc_stages = p.acc_stages;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)stages * stage_bytes);
    // bars: full[stages], ready[stages], empty[stages], tfull[acc_stages], tempty[acc_stages]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * stages + 2 * acc_stages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t full0 = smem_u32(bars), ready0 = smem_u32(bars + stages), empty0 = smem_u32(bars + 2 * stages),
                   tfull0 = smem_u32(bars + 3 * stages), tempty0 = smem_u32(bars + 3 * stages + acc_stages);
    const int num_kb = p.K / BLOCK_K;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(ready0 + 8 * s, 128);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int a = 0; a < acc_stages; ++a) {
            mbar_init(tfull0 + 8 * a, 1);
            mbar_init(tempty0 + 8 * a, 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"((uint32_t)p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t corr_off = p.corr_sep ? (uint32_t)p.acc_cols : 0u;       // TF32X3 lo-term accumulator
    const uint32_t acc_stride = (uint32_t)(p.corr_sep ? 2 * p.acc_cols : p.acc_cols);

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            const uint32_t tx = (uint32_t)(A_TILE_BYTES + (p.terms == 3 ? 2 : 1) * b_tile_bytes);
            uint32_t it = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int m0 = (t / p.num_n_tiles) * BLOCK_M, n0 = (t % p.num_n_tiles) * p.block_n;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % stages;
                    const uint32_t ph = (it / stages) & 1;
                    mbar_wait(empty0 + 8 * s, ph ^ 1);
                    uint8_t *st = smem + (size_t)s * stage_bytes;
                    mbar_expect_tx(full0 + 8 * s, tx);
                    tma_load_2d(smem_u32(st), &tmA, kb * BLOCK_K, m0, full0 + 8 * s);
                    if (p.terms == 3) {
                        tma_load_2d(smem_u32(st + 2 * A_TILE_BYTES), &tmBhi, kb * BLOCK_K, n0, full0 + 8 * s);
                        tma_load_2d(smem_u32(st + 2 * A_TILE_BYTES + b_tile_bytes), &tmBlo, kb * BLOCK_K, n0, full0 + 8 * s);
                    } else {
                        tma_load_2d(smem_u32(st + A_TILE_BYTES), &tmBhi, kb * BLOCK_K, n0, full0 + 8 * s);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.block_n >> 3) << 17) |
                                   ((uint32_t)(BLOCK_M >> 4) << 24);
            uint32_t it = 0, tl = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tl) {
                const int as = tl % acc_stages;
                mbar_wait(tempty0 + 8 * as, ((tl / acc_stages) & 1) ^ 1);     // epilogue drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_main = tmem_base + as * acc_stride, d_corr = d_main + corr_off;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % stages;
                    const uint32_t ph = (it / stages) & 1;
                    mbar_wait((p.terms == 3 ? ready0 : full0) + 8 * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    uint8_t *st = smem + (size_t)s * stage_bytes;
                    const uint32_t a_hi = smem_u32(st);
                    if (p.terms == 3) {
                        const uint32_t a_lo = a_hi + A_TILE_BYTES, b_hi = a_hi + 2 * A_TILE_BYTES, b_lo = b_hi + b_tile_bytes;
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                            const uint32_t off = k * UMMA_K * 4;
                            // lo terms first; with a separate accumulator (corr_sep) the long-running main sum
                            // sees a third of the additions (the tensor core's fp32 accumulate truncates)
                            umma_tf32(d_corr, make_desc(a_lo + off), make_desc(b_hi + off), idesc, (kb | k) != 0);
                            umma_tf32(d_corr, make_desc(a_hi + off), make_desc(b_lo + off), idesc, 1);
                            umma_tf32(d_main, make_desc(a_hi + off), make_desc(b_hi + off), idesc,
                                      p.corr_sep ? (uint32_t)((kb | k) != 0) : 1u);
                        }
                    } else {
                        const uint32_t b_hi = a_hi + A_TILE_BYTES;
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                            const uint32_t off = k * UMMA_K * 4;
                            umma_tf32(d_main, make_desc(a_hi + off), make_desc(b_hi + off), idesc, (kb | k) != 0);
                        }
                    }
                    umma_commit(empty0 + 8 * s);
                }
                umma_commit(tfull0 + 8 * as);
            }
        }
    } else if (warp < 6) {
        // ---------------- warps 2..5: split the landed A tile into hi / lo (TF32X3) ----------------
        if (p.terms == 3) {
            const int tt = threadIdx.x - 64;               // 0..127
            uint32_t it = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % stages;
                    const uint32_t ph = (it / stages) & 1;
                    mbar_wait(full0 + 8 * s, ph);
                    float4 *hi = reinterpret_cast<float4 *>(smem + (size_t)s * stage_bytes);
                    float4 *lo = hi + A_TILE_BYTES / 16;
#pragma unroll
                    for (int i = 0; i < A_TILE_BYTES / 16 / 128; ++i) {
                        const float4 v = hi[tt + 128 * i];
                        float4 h, l;
                        uint32_t u;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.x)); h.x = __uint_as_float(u); l.x = v.x - h.x;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.y)); h.y = __uint_as_float(u); l.y = v.y - h.y;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.z)); h.z = __uint_as_float(u); l.z = v.z - h.z;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.w)); h.w = __uint_as_float(u); l.w = v.w - h.w;
                        hi[tt + 128 * i] = h;
                        lo[tt + 128 * i] = l;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive(ready0 + 8 * s);
                }
            }
        }
    } else {
        // ---------------- warps 6..9: epilogue ----------------
        // TMEM gives each lane one accumulator ROW (32 columns per tcgen05.ld).  Stored directly that
        // is 32 different 128-byte lines per instruction (measured: the L1 wavefront replays made the
        // epilogue 3x longer than the main loop).  Each warp therefore transposes its 32x32 chunk through
        // a private padded shared-memory tile so that 8 lanes cover one 128-byte row segment and every
        // global access (residual read, C write) is a fully coalesced float4.
        const int quad = warp & 3;                         // TMEM lane quadrant this warp may access
        float *stg = reinterpret_cast<float *>(smem + (size_t)stages * stage_bytes + 512) + (warp - 6) * (32 * 36);
        const int sub_row = lane >> 3, col4 = lane & 7;    // transposed domain: 4 rows x 8 float4 per pass
        uint32_t tl = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tl) {
            const int as = tl % acc_stages;
            const int m0 = (t / p.num_n_tiles) * BLOCK_M, n_tile = t % p.num_n_tiles, n0 = n_tile * p.block_n;
            mbar_wait(tfull0 + 8 * as, (tl / acc_stages) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int64_t mw = (int64_t)m0 + quad * 32;     // first row of this warp
            const uint32_t t_main = tmem_base + as * acc_stride + ((uint32_t)(quad * 32) << 16);
            float dot[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) dot[i] = 0.f;
            // residual rows are prefetched one chunk ahead: 8 independent 16-byte loads per lane stay in
            // flight while the previous chunk is fetched from TMEM, transposed and stored (issued one by
            // one between the stores they were the epilogue's critical path: 2.0 ms -> 0.85 ms per layer)
            auto load_res = [&](int c0n, float4 (&dst)[8]) {
                const int ccn = c0n + 4 * col4;
                const bool okn = p.epi.residual != nullptr && c0n < p.block_n && ccn < p.block_n;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int64_t m = mw + 4 * i + sub_row;
                    dst[i] = (okn && m < p.M) ? ldg4(p.epi.residual + m * p.epi.ldr + n0 + ccn) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            };
            float4 res[8], res_next[8];
            load_res(0, res_next);
            for (int c0 = 0; c0 < p.block_n; c0 += 32) {
                const int cc = c0 + 4 * col4;              // column inside the tile of this lane's float4
                const bool col_ok = cc < p.block_n;        // block_n is a multiple of 16
                const int n = n0 + cc;
#pragma unroll
                for (int i = 0; i < 8; ++i) res[i] = res_next[i];
                load_res(c0 + 32, res_next);
                float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f), w4 = b4;
                if (col_ok) {
                    if (p.epi.col_scale != nullptr) s4 = ldg4(p.epi.col_scale + n);
                    if (p.epi.bias != nullptr) b4 = ldg4(p.epi.bias + n);
                    if (p.dot_w != nullptr) w4 = ldg4(p.dot_w + n);
                }
                uint32_t r[32];
                tmem_ld32(t_main + (uint32_t)c0, r);
                if (p.corr_sep) {
                    uint32_t r2[32];
                    tmem_ld32(t_main + corr_off + (uint32_t)c0, r2);
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4 *>(stg + lane * 36 + j) =
                        make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                    __uint_as_float(r[j + 3]));
                __syncwarp();
                if (col_ok) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = 4 * i + sub_row;
                        const int64_t m = mw + row;
                        float4 v = *reinterpret_cast<const float4 *>(stg + row * 36 + 4 * col4);
                        v.x = fmaf(v.x, s4.x, b4.x) + res[i].x; v.y = fmaf(v.y, s4.y, b4.y) + res[i].y;
                        v.z = fmaf(v.z, s4.z, b4.z) + res[i].z; v.w = fmaf(v.w, s4.w, b4.w) + res[i].w;
                        if (p.epi.relu) {
                            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                        }
                        if (m < p.M) {
                            if (p.C != nullptr) st4(p.C + m * p.ldc + n, v);
                            dot[i] = fmaf(v.x, w4.x, fmaf(v.y, w4.y, fmaf(v.z, w4.z, fmaf(v.w, w4.w, dot[i]))));
                        }
                    }
                }
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(tempty0 + 8 * as);                 // 128 arrivals free the accumulator stage
            if (p.dot_w != nullptr) {                      // fused row dot: out[n_tile][m] = sum_n v[m,n] * w[n]
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float d = dot[i];
                    d += __shfl_xor_sync(0xffffffffu, d, 1);
                    d += __shfl_xor_sync(0xffffffffu, d, 2);
                    d += __shfl_xor_sync(0xffffffffu, d, 4);
                    const int64_t m = mw + 4 * i + sub_row;
                    if (col4 == 0 && m < p.M) p.dot_out[(int64_t)n_tile * p.M + m] = d;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                     : "memory");
    }
}

// hi = rn_tf32(x), lo = x - hi  (exact).  One pass over the layer weight per call.
__global__ void k_split_tf32(const float *__restrict__ src, int64_t lds, float *__restrict__ hi, float *__restrict__ lo,
                             int rows, int cols) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)rows * cols) return;
    const int r = (int)(e / cols), c = (int)(e % cols);
    const float v = src[(int64_t)r * lds + c];
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
    const float h = __uint_as_float(u);
    hi[e] = h;
    lo[e] = v - h;
}
// Transposed variant for dgrad: out[c][r] from src[r][c].
__global__ void k_transpose_split_tf32(const float *__restrict__ src, int64_t lds, float *__restrict__ hi,
                                       float *__restrict__ lo, int rows, int cols) {
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(int64_t)r * lds + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (c < cols && r < rows) {
            const float v = tile[threadIdx.x][i];
            uint32_t u;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
            const float h = __uint_as_float(u);
            hi[(int64_t)c * rows + r] = h;
            if (lo != nullptr) lo[(int64_t)c * rows + r] = v - h;
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

static int make_map(CUtensorMap *tm, const float *base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return DCNR_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): rows %lld cols %lld ld %lld", (int)r, (long long)rows,
                  (long long)cols, (long long)ld);
        return DCNR_ERR_CUDA;
    }
    return DCNR_OK;
}

static int pick_block_n(int64_t n) {
    for (int bn = 256; bn >= 16; bn -= 16)
        if (n % bn == 0) return bn;
    return 0;
}

}  // namespace tc

int gemm_tc_n_tiles(int64_t n) {
    const int bn = tc::pick_block_n(n);
    return bn > 0 ? (int)(n / bn) : 1;
}

bool gemm_tc_supported(int precision, bool a_kmajor, bool b_kmajor, int64_t lda, int64_t ldb, int64_t ldc, int64_t m,
                       int64_t n, int64_t k, int split_k) {
    if (precision != DCNR_PREC_TF32X3 && precision != DCNR_PREC_TF32) return false;
    if (!a_kmajor || !b_kmajor || split_k != 1) return false;
    if (k < tc::BLOCK_K || k % tc::BLOCK_K != 0 || (lda & 3) || (ldb & 3) || (ldc & 3)) return false;
    if (n % 16 != 0 || tc::pick_block_n(n) == 0 || m <= 0 || m > 0x7fffff00LL) return false;
    return true;
}

int launch_split_tf32(const float *src, int64_t lds, float *hi, float *lo, int32_t rows, int32_t cols, bool transpose,
                      cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return DCNR_OK;
    if (transpose) {
        dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32)), block(32, 8);
        tc::k_transpose_split_tf32<<<grid, block, 0, stream>>>(src, lds, hi, lo, rows, cols);
    } else {
        const int64_t total = (int64_t)rows * cols;
        tc::k_split_tf32<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(src, lds, hi, lo, rows, cols);
    }
    DCNR_LAUNCHED();
    return DCNR_OK;
}

// B must be dense [n, k] (ldb == k is not required; any ldb % 4 == 0).  For TF32X3, B is the hi half
// and B_lo the lo half (both from launch_split_tf32); A is raw fp32 and is split inside the kernel.
int launch_gemm_tc(int precision, const float *A, int64_t lda, bool a_kmajor, const float *B, int64_t ldb, bool b_kmajor,
                   float *C, int64_t ldc, int64_t m, int64_t n, int64_t k, int split_k, const GemmEpilogue &epi,
                   cudaStream_t stream, const float *B_lo, const float *dot_w, float *dot_out) {
    using namespace tc;
    DCNR_REQUIRE(gemm_tc_supported(precision, a_kmajor, b_kmajor, lda, ldb, ldc, m, n, k, split_k),
                 "shape not supported by the tcgen05 GEMM");
    const int terms = precision == DCNR_PREC_TF32X3 ? 3 : 1;
    DCNR_REQUIRE(terms == 1 || B_lo != nullptr, "TF32X3 needs the pre-split weight (B_lo)");
    DCNR_REQUIRE((((uintptr_t)A | (uintptr_t)B | (uintptr_t)C | (uintptr_t)B_lo | (uintptr_t)dot_w) & 15) == 0,
                 "operands must be 16-byte aligned");
    DCNR_REQUIRE(C != nullptr || (dot_w != nullptr && dot_out != nullptr), "no output requested");
    DCNR_REQUIRE(epi.residual == nullptr || ((epi.ldr & 3) == 0 && ((uintptr_t)epi.residual & 15) == 0),
                 "residual must be 16-byte aligned");
    DCNR_REQUIRE((epi.bias == nullptr || ((uintptr_t)epi.bias & 15) == 0) &&
                     (epi.col_scale == nullptr || ((uintptr_t)epi.col_scale & 15) == 0),
                 "bias / col_scale must be 16-byte aligned");
    Params p;
    p.M = m; p.N_total = (int32_t)n; p.K = (int32_t)k;
    p.block_n = pick_block_n(n);
    p.terms = terms;
    p.C = C; p.ldc = ldc; p.epi = epi;
    p.dot_w = dot_w; p.dot_out = dot_out;
    p.acc_cols = 32;
    while (p.acc_cols < p.block_n) p.acc_cols <<= 1;
    p.corr_sep = (terms == 3 && 4 * p.acc_cols <= 512) ? 1 : 0;      // room for main + lo accumulators, double buffered
    const int per_stage_cols = p.corr_sep ? 2 * p.acc_cols : p.acc_cols;
    p.acc_stages = 2 * per_stage_cols <= 512 ? 2 : 1;
    p.tmem_cols = 32;
    while (p.tmem_cols < p.acc_stages * per_stage_cols) p.tmem_cols <<= 1;
    const int stage_bytes = (terms == 3 ? 2 : 1) * (A_TILE_BYTES + p.block_n * BLOCK_K * 4);
    int stages = std::max(1, std::min(6, (200 * 1024) / stage_bytes));
    p.stages = stages;
    p.num_m_tiles = (int32_t)ceil_div(m, BLOCK_M);
    p.num_n_tiles = (int32_t)(n / p.block_n);
    const size_t smem = (size_t)stages * stage_bytes + 1024 + 512 + kEpiStageBytes;
    CUtensorMap tmA, tmBhi, tmBlo;
    DCNR_TRY(make_map(&tmA, A, m, k, lda, BLOCK_M));
    DCNR_TRY(make_map(&tmBhi, B, n, k, ldb, p.block_n));
    DCNR_TRY(make_map(&tmBlo, terms == 3 ? B_lo : B, n, k, ldb, p.block_n));
    DCNR_CUDA_CHECK(cudaFuncSetAttribute(k_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t num_tiles = (int64_t)p.num_m_tiles * p.num_n_tiles;
    const unsigned grid = (unsigned)std::min<int64_t>(num_tiles, sm_count());
    k_gemm_tc<<<grid, kThreadsP, smem, stream>>>(tmA, tmBhi, tmBlo, p);
    DCNR_LAUNCHED();
    return DCNR_OK;
}

}  // namespace dcnr
