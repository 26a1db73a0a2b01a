// tcgen05 weight-gradient GEMM (sm_100a):  dW[n_out, k_in] = sum_b dY[b, n_out] * X[b, k_in]
//
// The reduction index is the BATCH row, so both operands are "MN-major" for the tensor core: dY^T
// is the A operand (M = n_out, contiguous in memory), X^T the B operand (N = k_in, contiguous).
// tcgen05.mma takes MN-major tf32 operands directly (instruction-descriptor bits 15/16), so nothing
// is transposed: a 3-D TMA box {32 floats, 32 batch rows, MN/32} lands the tile as MN/32 blocks of
// [32 rows x 128 bytes].  For 32-bit MN-major operands the only shared-memory layout the tensor core
// accepts is SWIZZLE_128B_BASE32B (32-byte chunks XOR-ed with the row index inside 4-row atoms;
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B on the TMA side): LBO = 4096 (next 32 floats of M/N),
// SBO = 512 (next 4 batch rows); one K = 8 MMA consumes two 4-row atoms of every block.
//
// Grid: (n_out / 128) x (batch slabs).  A CTA accumulates its slab's [128 x k_in] partial product in
// tensor memory over all its 32-row k-blocks and writes it once; the slabs are then added in slab
// order (launch_sum_partials_2d), so the result is deterministic.  TF32X3: both operands are fresh
// activations here, so the split warps split BOTH landed tiles in place (hi) + a lo copy, and the lo
// products get their own accumulator (512 TMEM columns: nothing needs double buffering because the
// epilogue runs once).
#include <cuda.h>

#include "kernels.cuh"

namespace dcnr {
namespace wg {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 16;                 // batch rows per k-block (32-row blocks left room for two stages only: hi + lo copies of both
                                            // tiles are 96 KB at k_in = 256; 16 rows = four stages, step 2.022 -> 2.004 ms)
constexpr int kSplitWarps = 16;              // warps that split the landed tiles (the k-block cadence was split-bound with four)
constexpr int kThreads = 64 + 32 * kSplitWarps;     // warp 0 TMA, warp 1 MMA, warps 2-9 split, warps 2-5 also the epilogue

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {      // bounded: a pipeline bug traps
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
// one lane of a converged warp (the issuing warps loop with all lanes so descriptors stay in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
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
// MN-major, SWIZZLE_128B_BASE32B: atoms of [4 k-rows x 128 B]; LBO = distance between 32-float blocks along M/N,
// SBO = distance between 4-row atoms along K.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(lbo_bytes >> 4) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;                             // descriptor version (Blackwell)
    d |= (uint64_t)1 << 61;                             // SWIZZLE_128B_BASE32B
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
    int64_t B;                 // batch rows (the reduction length)
    int32_t n_out, k_in;       // dW is [n_out, k_in]; k_in = the N tile (<= 256, multiple of 32)
    int32_t terms, stages, corr_sep, tmem_cols;
    int64_t rows_per_slab;     // multiple of BLOCK_K
    float *slabs;              // [n_slabs][n_out][k_in]
    const float *center;       // optional [k_in]: subtracted from every X row before the split (TF32X3 only; the caller
                               // adds colsum(dY) (x) center back -- see launch_linear_wgrad)
};

__global__ void __launch_bounds__(kThreads, 1)
k_wgrad_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment by OFFSETTING the __shared__ array: a round trip through uintptr_t loses the address space and
    // every shared-memory access below became a generic LD / ST (the split and the epilogue ran 3-4x slower)
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int a_bytes = BLOCK_M * BLOCK_K * 4;                 // 4 blocks of [BLOCK_K rows x 128 B]
    const int b_bytes = p.k_in * BLOCK_K * 4;                  // k_in / 32 blocks
    const int stage_bytes = (p.terms == 3 ? 2 : 1) * (a_bytes + b_bytes);      // [A hi][B hi]([A lo][B lo])
    const int stages = p.stages;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)stages * stage_bytes);
    // bars: full[stages], ready[stages], empty[stages], done
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * stages + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t full0 = smem_u32(bars), ready0 = smem_u32(bars + stages), empty0 = smem_u32(bars + 2 * stages),
                   done0 = smem_u32(bars + 3 * stages);
    const int mt = blockIdx.x, slab = blockIdx.y;
    const int64_t b0 = (int64_t)slab * p.rows_per_slab, b1 = min(p.B, b0 + p.rows_per_slab);
    const int num_kb = b1 > b0 ? (int)((b1 - b0 + BLOCK_K - 1) / BLOCK_K) : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(ready0 + 8 * s, kSplitWarps);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(done0, 1);
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
    const uint32_t d_main = tmem_base, d_corr = tmem_base + (p.corr_sep ? 256u : 0u);

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        int rs = 0; uint32_t rph = 0;                       // running stage / phase (no run-time division per k-block)
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = rs;
            const uint32_t ph = rph;
            if (++rs == stages) { rs = 0; rph ^= 1u; }
            mbar_wait(empty0 + 8 * s, ph ^ 1);
            uint8_t *st = smem + (size_t)s * stage_bytes;
            const int row = (int)(b0 + (int64_t)kb * BLOCK_K);           // rows past B are zero-filled: no contribution
            if (elect_one()) {
                mbar_expect_tx(full0 + 8 * s, (uint32_t)(a_bytes + b_bytes));
                tma_load_3d(smem_u32(st), &tmA, 0, row, mt * (BLOCK_M / 32), full0 + 8 * s);
                tma_load_3d(smem_u32(st + a_bytes), &tmB, 0, row, 0, full0 + 8 * s);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (all lanes loop, one elected lane issues) ----------------
        {
            // kind::tf32, fp32 accumulate, A and B MN-major (bits 15 / 16), N = k_in, M = 128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(p.k_in >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
            const uint32_t lbo = BLOCK_K * 128;                       // bytes between 32-float blocks
            int rs = 0; uint32_t rph = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = rs;
                const uint32_t ph = rph;
                if (++rs == stages) { rs = 0; rph ^= 1u; }
                mbar_wait((p.terms == 3 ? ready0 : full0) + 8 * s, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_hi = smem_u32(smem + (size_t)s * stage_bytes);
                if (elect_one()) {
                    const uint64_t da_hi = make_desc_mn(a_hi, lbo), db_hi = make_desc_mn(a_hi + a_bytes, lbo);
                    const uint64_t da_lo = make_desc_mn(a_hi + a_bytes + b_bytes, lbo);
                    const uint64_t db_lo = make_desc_mn(a_hi + 2 * a_bytes + b_bytes, lbo);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / 8; ++k) {
                        const uint64_t o = (uint64_t)(k * 1024 >> 4);        // next 8 batch rows (two 4-row atoms) of every block
                        const uint32_t first = (uint32_t)((kb | k) != 0);
                        if (p.terms == 3) {
                            umma_tf32(d_corr, da_lo + o, db_hi + o, idesc, first);
                            umma_tf32(d_corr, da_hi + o, db_lo + o, idesc, 1);
                            umma_tf32(d_main, da_hi + o, db_hi + o, idesc, p.corr_sep ? first : 1u);
                        } else {
                            umma_tf32(d_main, da_hi + o, db_hi + o, idesc, first);
                        }
                    }
                    umma_commit(empty0 + 8 * s);
                    if (kb == num_kb - 1) umma_commit(done0);
                }
                __syncwarp();
            }
            if (num_kb == 0 && elect_one()) umma_commit(done0);
        }
    } else {
        // ---------------- warps 2..5: split both tiles (TF32X3), then the epilogue ----------------
        const int tt = threadIdx.x - 64;                   // 0..127
        if (p.terms == 3) {
            const int n4 = (a_bytes + b_bytes) / 16;       // float4 of [A][B] (contiguous in the stage)
            const int a4 = a_bytes / 16;
            int rs = 0; uint32_t rph = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = rs;
                const uint32_t ph = rph;
                if (++rs == stages) { rs = 0; rph ^= 1u; }
                mbar_wait(full0 + 8 * s, ph);
                float4 *hi = reinterpret_cast<float4 *>(smem + (size_t)s * stage_bytes);
                float4 *lo = hi + n4;
                for (int i = tt; i < n4; i += 32 * kSplitWarps) {
                    float4 v = hi[i];
                    if (p.center != nullptr && i >= a4) {
                        // X tile, SWIZZLE_128B_BASE32B: block of 32 columns = 32 batch rows x 128 B; the 32-byte chunk of a row
                        // is XOR-ed with (row & 3).  float4 i of the tile -> its first column.
                        const int o = i - a4, row = (o >> 3) % BLOCK_K, c16 = o & 7;
                        const int col = (o / (BLOCK_K * 8)) * 32 + (((c16 >> 1) ^ (row & 3)) << 3) + ((c16 & 1) << 2);
                        const float4 mu = __ldg(reinterpret_cast<const float4 *>(p.center + col));
                        v.x -= mu.x; v.y -= mu.y; v.z -= mu.z; v.w -= mu.w;
                    }
                    float4 h, l;
                    uint32_t u;
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.x)); h.x = __uint_as_float(u); l.x = v.x - h.x;
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.y)); h.y = __uint_as_float(u); l.y = v.y - h.y;
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.z)); h.z = __uint_as_float(u); l.z = v.z - h.z;
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.w)); h.w = __uint_as_float(u); l.w = v.w - h.w;
                    hi[i] = h;
                    lo[i] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(ready0 + 8 * s);
            }
        }
        // epilogue (warps 2-5): lane = one row of the [128 x k_in] partial product; each lane writes whole 128-byte lines
        if (warp < 6) {
        mbar_wait(done0, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quad = warp & 3;
        const int m = mt * BLOCK_M + quad * 32 + lane;
        float *orow = p.slabs + ((int64_t)slab * p.n_out + m) * p.k_in;
        const uint32_t t_main = tmem_base + ((uint32_t)(quad * 32) << 16);
        for (int c0 = 0; c0 < p.k_in; c0 += 32) {
            uint32_t r[32];
            if (num_kb > 0) {
                tmem_ld32(t_main + (uint32_t)c0, r);
                if (p.corr_sep) {
                    uint32_t r2[32];
                    tmem_ld32(t_main + 256u + (uint32_t)c0, r2);
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = 0u;
            }
            if (m < p.n_out) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    st4(orow + c0 + j, make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                                   __uint_as_float(r[j + 3])));
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
// [rows, cols] row-major (ld) seen as {32 floats, rows, cols / 32}; box {32, BLOCK_K, blocks}
static int make_map_mn(CUtensorMap *tm, const float *base, int64_t rows, int64_t cols, int64_t ld, int blocks) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return DCNR_ERR_CUDA;
    }
    cuuint64_t dims[3] = {32, (cuuint64_t)rows, (cuuint64_t)(cols / 32)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 4, 128};
    cuuint32_t box[3] = {32, (cuuint32_t)BLOCK_K, (cuuint32_t)blocks};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (3d) failed (%d): rows %lld cols %lld ld %lld", (int)r, (long long)rows,
                  (long long)cols, (long long)ld);
        return DCNR_ERR_CUDA;
    }
    return DCNR_OK;
}

}  // namespace wg

bool wgrad_tc_supported(int precision, int64_t lddy, int64_t ldx, int64_t m, int32_t n, int32_t k) {
    if (precision != DCNR_PREC_TF32X3 && precision != DCNR_PREC_TF32) return false;
    if (n % wg::BLOCK_M != 0 || k % 32 != 0 || k > 256 || k < 32) return false;
    if ((lddy & 3) || (ldx & 3) || lddy < n || ldx < k) return false;
    return m >= 1 && m <= 0x7fffff00LL;
}

// Batch slabs of at most kMaxSlabRows rows.  The tensor core's fp32 accumulate truncates (profiles/r02_acc_probe.md), and tensor
// memory bounds how many accumulation chains can run side by side (148 SMs x 512 columns = 74 chains per element of a 256 x 256
// gradient, main + correction accumulators), so the only way to shorter chains is more CTAs per SM, one after the other.
// Measured on the B = 65 536 gradients against float64 (worst weight gradient, profiles/r02_parity_65536.md): 886-row slabs
// (one wave of CTAs) 1.1e-5, 256-row slabs (3.5 waves) 5.5e-6 -- the error goes with the square root of the chain length --
// at 60 / 87 us per 65 536 x 256 x 256 gradient.  448 rows = two full waves at that batch.  Slabs are summed in double, in order.
constexpr int64_t kMaxSlabRows = 448;
int wgrad_tc_slabs(int64_t m, int32_t n) {
    const int64_t m_tiles = n / wg::BLOCK_M;
    const int64_t want = std::max<int64_t>(1, (int64_t)sm_count() / m_tiles);
    const int64_t rows = std::min<int64_t>(kMaxSlabRows, round_up(ceil_div(m, want), wg::BLOCK_K));
    return (int)ceil_div(m, rows);
}

// slabs: [wgrad_tc_slabs(m, n)][n][k] floats; the caller adds them in slab order
int launch_wgrad_tc(int precision, const float *dy, int64_t lddy, const float *x, int64_t ldx, float *slabs, int64_t m,
                    int32_t n, int32_t k, cudaStream_t stream, const float *center) {
    using namespace wg;
    DCNR_REQUIRE(wgrad_tc_supported(precision, lddy, ldx, m, n, k), "shape not supported by the tcgen05 wgrad");
    DCNR_REQUIRE((((uintptr_t)dy | (uintptr_t)x | (uintptr_t)slabs) & 15) == 0, "operands must be 16-byte aligned");
    Params p;
    p.B = m; p.n_out = n; p.k_in = k;
    p.terms = precision == DCNR_PREC_TF32X3 ? 3 : 1;
    p.corr_sep = p.terms == 3 ? 1 : 0;
    p.tmem_cols = p.corr_sep ? 512 : 256;
    const int n_slabs = wgrad_tc_slabs(m, n);
    p.rows_per_slab = std::min<int64_t>(kMaxSlabRows, round_up(ceil_div(m, n_slabs), BLOCK_K));
    p.slabs = slabs;
    DCNR_REQUIRE(center == nullptr || (p.terms == 3 && ((uintptr_t)center & 15) == 0), "centred wgrad needs TF32X3 and an aligned vector");
    p.center = center;
    const int stage_bytes = (p.terms == 3 ? 2 : 1) * (BLOCK_M + k) * BLOCK_K * 4;
    p.stages = std::max(1, std::min(6, (220 * 1024) / stage_bytes));
    const size_t smem = (size_t)p.stages * stage_bytes + 1024 + 256;
    CUtensorMap tmA, tmB;
    DCNR_TRY(make_map_mn(&tmA, dy, m, n, lddy, BLOCK_M / 32));
    DCNR_TRY(make_map_mn(&tmB, x, m, k, ldx, k / 32));
    DCNR_CUDA_CHECK(cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)(n / BLOCK_M), (unsigned)n_slabs);
    k_wgrad_tc<<<grid, kThreads, smem, stream>>>(tmA, tmB, p);
    DCNR_LAUNCHED();
    return DCNR_OK;
}

}  // namespace dcnr
