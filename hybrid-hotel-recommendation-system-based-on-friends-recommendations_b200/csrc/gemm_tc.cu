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

#include <cstdio>
#include <cstdlib>

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
// mbarrier wait that synchronises with arrivals from the other CTA of a pair (remote arrive / multicast commit)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t spins = 0; !ok; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!ok && spins > (1u << 24)) __trap();
    }
}
// arrive on a barrier given by its shared::cluster address (possibly in the peer CTA)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// pair form: the completion bytes are counted on `cluster_bar`, a barrier of the pair's leader CTA
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap *tm, int c0, int c1, uint32_t cluster_bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(cluster_bar)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
// One lane of a fully converged warp.  The TMA / MMA issuing warps run their loops with ALL lanes (warp-uniform control
// flow, so addresses and descriptors live in uniform registers) and only predicate the issue itself: inside an
// `if (lane == 0)` region the compiler wraps every UTCHMMA / UTMALDG in an elect + 4x R2UR.BROADCAST loop (~90 clk per MMA).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tm, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(src)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {      // at most N most recent bulk groups still READING shared memory
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// A operand from tensor memory (rows = TMEM lanes, K = 8 consecutive 32-bit columns), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,"
        "%31,%32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 2-CTA MMA (M = 256 over the pair, each CTA feeds its 128 rows of A and half of B's rows); leader thread only
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair once the MMAs issued so far are done
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
        ::"r"(bar)
        : "memory");
}
// K-major, SWIZZLE_128B canonical layout (128-byte rows): 8-row groups 1024 B apart (SBO), LBO unused (1), version 1.
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
    int32_t epi_slots;        // epilogue slots per warp (2 or 4)
    int32_t epi_depth;        // residual boxes requested this many chunks ahead (<= epi_slots - 1); the slot being refilled was
                              // stored epi_slots - epi_depth chunks ago, so that many TMA stores may still be reading
    int32_t epi_groups;       // 1: warps 6-9 drain the accumulator; 2: warps 10-13 as well (alternate 32-column chunks)
    int32_t a_col0;           // TF32X3: the A hi / lo ring in tensor memory (64 columns per stage) starts at this column
    int32_t num_m_tiles, num_n_tiles;
    float *C;                 // may be NULL when only the fused row dot is wanted
    int64_t ldc;
    GemmEpilogue epi;
    const float *dot_w;       // optional fused row dot with the epilogue output: [N_total]
    float *dot_out;           // [num_n_tiles * epi_groups][M] partial dots (summed by the caller)
};

constexpr int kThreadsP = 448;                      // warp 0 TMA, warp 1 MMA, warps 2-5 A split, warps 6-9 epilogue, warps 10-13 second epilogue group
constexpr int kEpiSlotBytes = 32 * 32 * 4;          // one epilogue slot: 32 rows x 32 fp32 columns, 128B-swizzled
constexpr int kEpiSlotsMax = 4;                     // slots per epilogue warp: 4 (residual prefetch depth 3) or 2 (one more operand stage)
constexpr int kBarBytes = 512;                      // mbarriers + the TMEM base slot

// Persistent, warp-specialised: every CTA (CTAS = 1) or CTA pair (CTAS = 2, a 2-CTA cluster) walks
// output tiles t = id, id + n, ...  Three rings run concurrently: shared-memory stages (TMA -> split
// -> MMA), TMEM accumulator stages (MMA -> epilogue) and the tile sequence itself, so the epilogue of
// tile i overlaps the loads and MMAs of tile i+1.
//
// CTAS = 2: one tcgen05.mma.cta_group::2 covers 256 rows -- each CTA stages its own 128 rows of A and
// HALF of the weight tile's rows, so the weight stream from L2 (the measured bottleneck of the 1-CTA
// kernel: 64 KB per k-block per SM against a ~42 B/clk/SM L2 share) and the shared-memory operand
// reads per CTA are halved.  The MMAs are issued by the leader CTA (cluster rank 0) only; barriers
// that gate them (fullB, ready, tempty) live in the leader and are arrived on remotely by the peer;
// barriers the leader releases (empty, tfull) are signalled in both CTAs by a multicast commit.
// TF32X3 always runs single CTAs with the A operand in tensor memory (the 2-CTA and all-shared-memory TF32X3 forms were
// measured slower in round 1 and are gone: profiles/r01_gemm_pipeline_timing.md); CTAS = 2 serves single-pass TF32.
// HAD: Hadamard operand in the epilogue (DCN-v2 cross layer).  EXTRA: warps 10-13 exist (second epilogue group for short
// reductions).  Both are compiled OUT of the dense 256 x 256 layers: as run-time branches they cost the epilogue-bound
// tf32 form 15-40 % (0.62 -> 0.72 -> 0.94 ms) and the tf32x3 form 3 %.
template <int CTAS, bool HAD, bool EXTRA>
__global__ void __launch_bounds__(EXTRA ? kThreadsP : 320, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
          const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmR,
          const __grid_constant__ CUtensorMap tmC, Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment by OFFSETTING the __shared__ array: a round trip through uintptr_t loses the address space and
    // every shared-memory access below became a generic LD / ST (the split and the epilogue ran 3-4x slower)
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t rank = CTAS == 2 ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const int bn_cta = p.block_n / CTAS;                 // weight rows staged by this CTA
    constexpr int bk = BLOCK_K;                          // floats per k-block
    const int a_tile_bytes = BLOCK_M * bk * 4;
    const int b_tile_bytes = bn_cta * bk * 4;
    // stage: [A][B hi][B lo] (TF32X3: the raw A tile, split into tensor memory by warps 2-5), [A][B] (TF32)
    const int b_off = a_tile_bytes;
    const int stage_bytes = b_off + (p.terms == 3 ? 2 : 1) * b_tile_bytes;
    const int stages = p.stages, acc_stages = p.acc_stages;
    uint8_t *epi_slots = smem + (size_t)stages * stage_bytes;            // 1024-aligned (stage sizes are multiples of 1 KB)
    const int kEpiSlots = p.epi_slots;
    uint64_t *bars = reinterpret_cast<uint64_t *>(epi_slots + 4 * (EXTRA ? p.epi_groups : 1) * kEpiSlots * kEpiSlotBytes);
    // bars: fullA[stages], fullB[stages], ready[stages], empty[stages], tfull[acc_stages], tempty[acc_stages], rfull[8][kEpiSlots]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4 * stages + 2 * acc_stages + 8 * kEpiSlotsMax);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t fullA0 = smem_u32(bars), fullB0 = smem_u32(bars + stages), ready0 = smem_u32(bars + 2 * stages),
                   empty0 = smem_u32(bars + 3 * stages), tfull0 = smem_u32(bars + 4 * stages),
                   tempty0 = smem_u32(bars + 4 * stages + acc_stages),
                   rfull0 = smem_u32(bars + 4 * stages + 2 * acc_stages);
    // the same barriers in the leader CTA, as shared::cluster addresses (identity for CTAS == 1)
    const uint32_t L_fullB0 = CTAS == 2 ? mapa(fullB0, 0) : fullB0, L_tempty0 = CTAS == 2 ? mapa(tempty0, 0) : tempty0;
    const int num_kb = p.K / bk;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    const int tile0 = blockIdx.x / CTAS, tile_step = gridDim.x / CTAS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(fullA0 + 8 * s, 1);
            mbar_init(fullB0 + 8 * s, 1);
            mbar_init(ready0 + 8 * s, 4 + 1);         // TF32X3: one arrival per A-split warp + the producer's arrive.expect_tx (weights)
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int a = 0; a < acc_stages; ++a) {
            mbar_init(tfull0 + 8 * a, 1);
            mbar_init(tempty0 + 8 * a, 4 * CTAS * (EXTRA ? p.epi_groups : 1));   // one arrival per epilogue warp
        }
        for (int i = 0; i < 8 * kEpiSlots; ++i) mbar_init(rfull0 + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (CTAS == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"((uint32_t)p.tmem_cols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                         "r"((uint32_t)p.tmem_cols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CTAS == 2) cluster_sync();                     // the peer's barriers are initialised before anyone signals them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t corr_off = p.corr_sep ? (uint32_t)p.acc_cols : 0u;       // TF32X3 lo-term accumulator
    const uint32_t acc_stride = (uint32_t)(p.corr_sep ? 2 * p.acc_cols : p.acc_cols);

    auto wait_x = [](uint32_t bar, uint32_t parity) {    // barriers signalled from the other CTA / by multicast commits
        if (CTAS == 2) mbar_wait_cluster(bar, parity);
        else mbar_wait(bar, parity);
    };

    if (warp == 0) {
        // ---------------- TMA producer (both CTAs of a pair) ----------------
        {
            uint32_t it = 0;
            int rs = 0; uint32_t rph = 0;
            for (int t = tile0; t < num_tiles; t += tile_step) {
                const int m0 = (t / p.num_n_tiles) * (BLOCK_M * CTAS) + (int)rank * BLOCK_M;
                const int n0 = (t % p.num_n_tiles) * p.block_n + (int)rank * bn_cta;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = rs;
                    const uint32_t ph = rph;
                    if (++rs == stages) { rs = 0; rph ^= 1u; }
                    wait_x(empty0 + 8 * s, ph ^ 1);
                    uint8_t *st = smem + (size_t)s * stage_bytes;
                    const uint32_t fb = L_fullB0 + 8 * s;
                    if (elect_one()) {
                    if (p.terms == 3) {
                        // TF32X3 (single CTAs): A lands on its own barrier (the split warps wait for it); the pre-split weight
                        // boxes count on "ready" next to the split warps' arrivals, so the issuing warp has ONE wait per k-block
                        mbar_expect_tx(fullA0 + 8 * s, (uint32_t)a_tile_bytes);
                        tma_load_2d(smem_u32(st), &tmA, kb * bk, m0, fullA0 + 8 * s);
                        mbar_expect_tx(ready0 + 8 * s, (uint32_t)(2 * b_tile_bytes));
                        tma_load_2d(smem_u32(st + b_off), &tmBhi, kb * bk, n0, ready0 + 8 * s);
                        tma_load_2d(smem_u32(st + b_off + b_tile_bytes), &tmBlo, kb * bk, n0, ready0 + 8 * s);
                    } else {
                        if (leader) mbar_expect_tx(fullB0 + 8 * s, (uint32_t)(CTAS * (a_tile_bytes + b_tile_bytes)));
                        if (CTAS == 2) {
                            tma_load_2d_pair(smem_u32(st), &tmA, kb * bk, m0, fb);
                            tma_load_2d_pair(smem_u32(st + a_tile_bytes), &tmBhi, kb * bk, n0, fb);
                        } else {
                            tma_load_2d(smem_u32(st), &tmA, kb * bk, m0, fb);
                            tma_load_2d(smem_u32(st + a_tile_bytes), &tmBhi, kb * bk, n0, fb);
                        }
                    }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (leader CTA only; all lanes loop, one elected lane issues) ----------------
        if (leader) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.block_n >> 3) << 17) |
                                   ((uint32_t)((BLOCK_M * CTAS) >> 4) << 24);
            // descriptors are built once per operand per k-block; a k-step only adds 32 bytes (2 in the >>4 address field),
            // which keeps the single issuing thread at a few instructions per MMA (it was ~90 clk per MMA, i.e. issue-bound
            // for 128-column tiles and close to it for 256-column ones)
            auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
                if (CTAS == 2) umma_tf32_pair(d, a, b, idesc, acc);
                else umma_tf32(d, a, b, idesc, acc);
            };
            auto mma_ts = [&](uint32_t d, uint32_t a, uint64_t b, uint32_t acc) { umma_tf32_ts(d, a, b, idesc, acc); };
            uint32_t it = 0, tl = 0;
            int rs = 0; uint32_t rph = 0;      // running stage index / phase bit (no runtime division on the issue path)
            int ras = 0; uint32_t raph = 0;    // running accumulator stage / phase
            if (p.terms == 3) {
                // Default TF32X3 form.  The tensor pipe's queue is shallow: tcgen05.mma issue blocks on it, so whatever the issuing
                // warp does between the last MMA of a k-block and the first of the next is dead time for the pipe (stamps: ~760 clk
                // of a 1 530 clk k-block: two barrier waits of ~240 clk each although both had completed long before).  Here there
                // is ONE barrier per k-block, and it is probed (non-blocking) for the NEXT k-block in the middle of this one's
                // MMAs, when the issue is blocked on the queue anyway.
                uint32_t pre_ok = 0;
                for (int t = tile0; t < num_tiles; t += tile_step) {
                    const int as = ras;
                    const uint32_t aph = raph;
                    if (++ras == acc_stages) { ras = 0; raph ^= 1u; }
                    mbar_wait(tempty0 + 8 * as, aph ^ 1);        // epilogues drained this accumulator
                    const uint32_t d_main = tmem_base + as * acc_stride, d_corr = d_main + corr_off;
                    for (int kb = 0; kb < num_kb; ++kb) {
                        const int s = rs;
                        const uint32_t ph = rph;
                        if (++rs == stages) { rs = 0; rph ^= 1u; }
                        if (!pre_ok) mbar_wait(ready0 + 8 * s, ph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t a_hi = smem_u32(smem + (size_t)s * stage_bytes);
                        const uint32_t t_hi = tmem_base + (uint32_t)p.a_col0 + (uint32_t)(s * 2 * BLOCK_K), t_lo = t_hi + (uint32_t)BLOCK_K;
                        const uint64_t db_hi = make_desc(a_hi + b_off), db_lo = make_desc(a_hi + b_off + b_tile_bytes);
                        const uint32_t acc_main = p.corr_sep ? (uint32_t)(kb != 0) : 1u;
                        if (elect_one()) {
                            mma_ts(d_corr, t_lo, db_hi, kb != 0);
                            mma_ts(d_corr, t_hi, db_lo, 1);
                            mma_ts(d_main, t_hi, db_hi, acc_main);
                            mma_ts(d_corr, t_lo + 8, db_hi + 2, 1);
                            mma_ts(d_corr, t_hi + 8, db_lo + 2, 1);
                            mma_ts(d_main, t_hi + 8, db_hi + 2, 1);
                        }
                        __syncwarp();
                        {                                      // probe the next k-block's barrier while the queue drains
                            uint32_t ok;
                            asm volatile(
                                "{\n\t.reg .pred p;\n\t"
                                "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                                "selp.u32 %0, 1, 0, p;\n\t}"
                                : "=r"(ok)
                                : "r"(ready0 + 8 * rs), "r"(rph)
                                : "memory");
                            pre_ok = __shfl_sync(0xffffffffu, ok, 0);
                        }
                        if (elect_one()) {
                            mma_ts(d_corr, t_lo + 16, db_hi + 4, 1);
                            mma_ts(d_corr, t_hi + 16, db_lo + 4, 1);
                            mma_ts(d_main, t_hi + 16, db_hi + 4, 1);
                            mma_ts(d_corr, t_lo + 24, db_hi + 6, 1);
                            mma_ts(d_corr, t_hi + 24, db_lo + 6, 1);
                            mma_ts(d_main, t_hi + 24, db_hi + 6, 1);
                            umma_commit(empty0 + 8 * s);
                            if (kb == num_kb - 1) umma_commit(tfull0 + 8 * as);
                        }
                        __syncwarp();
                    }
                }
            } else
            for (int t = tile0; t < num_tiles; t += tile_step, ++tl) {
                // single-pass TF32: A and B from shared memory (single CTAs or 2-CTA pairs)
                const int as = ras;
                const uint32_t aph = raph;
                if (++ras == acc_stages) { ras = 0; raph ^= 1u; }
                wait_x(tempty0 + 8 * as, aph ^ 1);        // epilogues drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_main = tmem_base + as * acc_stride;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = rs;
                    const uint32_t ph = rph;
                    if (++rs == stages) { rs = 0; rph ^= 1u; }
                    wait_x(fullB0 + 8 * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = smem_u32(smem + (size_t)s * stage_bytes);
                    if (elect_one()) {
                        const uint64_t da = make_desc(a_hi), db = make_desc(a_hi + a_tile_bytes);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                            const uint64_t o = (uint64_t)(k * UMMA_K * 4 >> 4);
                            mma(d_main, da + o, db + o, (kb | k) != 0);
                        }
                        if (CTAS == 2) umma_commit_pair(empty0 + 8 * s);
                        else umma_commit(empty0 + 8 * s);
                        if (kb == num_kb - 1) {                // same elected lane: the commit covers every MMA of this tile
                            if (CTAS == 2) umma_commit_pair(tfull0 + 8 * as);
                            else umma_commit(tfull0 + 8 * as);
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp < 6) {
        // ---------------- warps 2..5: split the landed A tile into hi / lo in tensor memory (TF32X3) ----------------
        if (p.terms == 3) {
            // One thread per tile row: it reads its 128-byte row of the k-block from the swizzled TMA tile (chunk c of
            // row r sits at position c ^ (r & 7): conflict-free), splits it and writes hi / lo straight into the TMEM
            // operand ring (tcgen05.st: lane = row).  No hi / lo tiles in shared memory and no shared-memory operand
            // reads for A: the shared-memory traffic that bounds the all-shared-memory variant drops by a third.
            const int quad = warp & 3;                     // TMEM lane quadrant of this warp = rows 32*quad .. +31
            const int r = quad * 32 + lane;
            const uint32_t swz = (uint32_t)(r & 7);
            uint32_t it = 0;
            int rs = 0; uint32_t rph = 0;
            for (int t = tile0; t < num_tiles; t += tile_step) {
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = rs;
                    const uint32_t ph = rph;
                    if (++rs == stages) { rs = 0; rph ^= 1u; }
                    mbar_wait(fullA0 + 8 * s, ph);
                    const uint8_t *row = smem + (size_t)s * stage_bytes + r * (bk * 4);
                    uint32_t hi[32], lo[32];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 v = *reinterpret_cast<const float4 *>(row + (((uint32_t)c ^ swz) << 4));
                        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            uint32_t u;
                            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(e[q]));
                            hi[4 * c + q] = u;
                            lo[4 * c + q] = __float_as_uint(e[q] - __uint_as_float(u));
                        }
                    }
                    const uint32_t t_hi = tmem_base + (uint32_t)p.a_col0 + (uint32_t)(s * 2 * bk) + ((uint32_t)(quad * 32) << 16);
                    tmem_st32(t_hi, hi);
                    tmem_st32(t_hi + 32u, lo);
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(ready0 + 8 * s);
                }
            }
        }
    } else if (warp < 10 || (EXTRA && p.epi_groups == 2)) {
        // ---------------- warps 6..9 (and 10..13): epilogue (each CTA drains its own 128 accumulator rows) ----------------
        // TMEM gives each lane one accumulator ROW (32 columns per tcgen05.ld), so the arithmetic is done
        // row-per-thread; all global traffic is TMA.  A warp owns kEpiSlots shared-memory slots of
        // 32 rows x 32 columns (128B-swizzled, so a lane reads / writes its own 128-byte row without bank
        // conflicts).  Per 32-column chunk: the residual box was TMA-loaded into the slot up to three chunks
        // ahead (48 KB in flight per SM -- with register-staged loads the epilogue was latency-bound at
        // ~2.4 TB/s), the lane combines accumulator, scale, bias, residual, ReLU in place, and one lane
        // TMA-stores the slot to C (rows past M are clipped by the tensor map).
        // With epi_groups == 2 the warps 10..13 join in: warp w and warp w + 4 share a TMEM lane quadrant and take
        // alternate 32-column chunks of every tile (the K = 64 initial layer is epilogue-bound with four warps).
        const int ew = warp - 6;                           // 0..7
        const int eg = EXTRA ? (ew >> 2) : 0, ngroups = EXTRA ? p.epi_groups : 1;    // this warp's chunk phase
        const int quad = warp & 3;                         // TMEM lane quadrant this warp may access
        uint8_t *slots = epi_slots + (size_t)ew * kEpiSlots * kEpiSlotBytes;
        const uint32_t rf0 = rfull0 + 8 * ew * kEpiSlots;
        const bool has_res = p.epi.residual != nullptr, has_c = p.C != nullptr, has_had = HAD && p.epi.hadamard != nullptr;
        const int cpt = (p.block_n + 31) / 32 / ngroups;   // chunks per tile for this warp (block_n / 32 is even when ngroups == 2)
        const uint32_t swz = (uint32_t)(lane & 7);
        // residual prefetch cursor: boxes are requested in chunk order, so tile / chunk / slot advance incrementally
        // (runtime divisions per chunk on the issuing lane were a measurable part of the epilogue)
        int pl_t = tile0, pl_c = 0, pl_sl = 0;
        int pl_m0 = (pl_t / p.num_n_tiles) * (BLOCK_M * CTAS) + (int)rank * BLOCK_M + quad * 32;
        int pl_n0 = (pl_t % p.num_n_tiles) * p.block_n;
        auto issue_res_load = [&]() {                      // lane 0: the next residual box into the next slot
            if (pl_t < num_tiles) {
                mbar_expect_tx(rf0 + 8 * pl_sl, (uint32_t)kEpiSlotBytes);
                tma_load_2d(smem_u32(slots + pl_sl * kEpiSlotBytes), &tmR, pl_n0 + (pl_c * ngroups + eg) * 32, pl_m0, rf0 + 8 * pl_sl);
            }
            if (++pl_sl == kEpiSlots) pl_sl = 0;
            if (++pl_c == cpt) {
                pl_c = 0;
                pl_t += tile_step;
                pl_m0 = (pl_t / p.num_n_tiles) * (BLOCK_M * CTAS) + (int)rank * BLOCK_M + quad * 32;
                pl_n0 = (pl_t % p.num_n_tiles) * p.block_n;
            }
        };
        if (has_res && lane == 0)
            for (int q = 0; q < p.epi_depth; ++q) issue_res_load();
        uint32_t tl = 0, g = 0;
        int ras = 0; uint32_t raph = 0;        // running accumulator stage / phase
        int rsl = 0; uint32_t rsph = 0;        // running epilogue slot / phase
        for (int t = tile0; t < num_tiles; t += tile_step, ++tl) {
            const int as = ras;
            const uint32_t aph = raph;
            if (++ras == acc_stages) { ras = 0; raph ^= 1u; }
            const int m0 = (t / p.num_n_tiles) * (BLOCK_M * CTAS) + (int)rank * BLOCK_M;
            const int n_tile = t % p.num_n_tiles, n0 = n_tile * p.block_n;
            wait_x(tfull0 + 8 * as, aph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int mw = m0 + quad * 32;                  // first row of this warp
            const float *had_row = (has_had && (int64_t)mw + lane < p.M)
                                       ? p.epi.hadamard + ((int64_t)mw + lane) * p.epi.ldh + n0 : nullptr;
            const uint32_t t_main = tmem_base + as * acc_stride + ((uint32_t)(quad * 32) << 16);
            float dot = 0.f;
            for (int c0 = 32 * eg; c0 < p.block_n; c0 += 32 * ngroups, ++g) {
                const uint32_t sl = (uint32_t)rsl;
                const uint32_t slph = rsph;
                if (++rsl == kEpiSlots) { rsl = 0; rsph ^= 1u; }
                uint8_t *row = slots + sl * kEpiSlotBytes + lane * 128;
                if (has_res) {
                    mbar_wait(rf0 + 8 * sl, slph);
                } else if (has_c) {                         // the store that last used this slot has finished reading it
                    if (lane == 0) {
                        if (kEpiSlots == 4) bulk_wait_read<3>();
                        else bulk_wait_read<1>();
                    }
                    __syncwarp();
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
                for (int j = 0; j < 8; ++j) {
                    float4 *cell = reinterpret_cast<float4 *>(row + (((uint32_t)j ^ swz) << 4));
                    // per-column vectors straight from global memory: warp-uniform addresses, L1 hits after the first tile
                    const float4 s4 = p.epi.col_scale != nullptr ? ldg4(p.epi.col_scale + n0 + c0 + 4 * j) : make_float4(1.f, 1.f, 1.f, 1.f);
                    const float4 b4 = p.epi.bias != nullptr ? ldg4(p.epi.bias + n0 + c0 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 v;
                    v.x = fmaf(__uint_as_float(r[4 * j]), s4.x, b4.x); v.y = fmaf(__uint_as_float(r[4 * j + 1]), s4.y, b4.y);
                    v.z = fmaf(__uint_as_float(r[4 * j + 2]), s4.z, b4.z); v.w = fmaf(__uint_as_float(r[4 * j + 3]), s4.w, b4.w);
                    if (HAD && has_had) {                   // row-per-lane 128-byte segments straight from global memory
                        const float4 h4 = had_row != nullptr ? ldg4(had_row + c0 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
                        v.x *= h4.x; v.y *= h4.y; v.z *= h4.z; v.w *= h4.w;
                    }
                    if (has_res) {
                        const float4 q = *cell;
                        v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
                    }
                    if (p.epi.relu) {
                        v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                    }
                    if (p.dot_w != nullptr) {
                        const float4 w4 = ldg4(p.dot_w + n0 + c0 + 4 * j);
                        dot = fmaf(v.x, w4.x, fmaf(v.y, w4.y, fmaf(v.z, w4.z, fmaf(v.w, w4.w, dot))));
                    }
                    if (has_c) *cell = v;
                }
                if (has_c) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    if (has_c) {
                        tma_store_2d(&tmC, smem_u32(slots + sl * kEpiSlotBytes), n0 + c0, mw);
                        bulk_commit();
                    }
                    if (has_res) {                          // refill the slot chunk g-1 used: its store must be done reading
                        if (has_c) {
                            if (kEpiSlots - p.epi_depth >= 2) bulk_wait_read<2>();
                            else bulk_wait_read<1>();
                        }
                        issue_res_load();
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {                                // one arrival per epilogue warp frees the accumulator stage
                if (CTAS == 2) mbar_arrive_cluster(L_tempty0 + 8 * as);
                else mbar_arrive(tempty0 + 8 * as);
            }
            if (p.dot_w != nullptr && (int64_t)mw + lane < p.M)        // fused row dot: out[n_tile][m] = sum_n v[m,n] * w[n]
                p.dot_out[((int64_t)n_tile * ngroups + eg) * p.M + mw + lane] = dot;
        }
        if (lane == 0) bulk_wait_all();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CTAS == 2) cluster_sync();                     // no CTA leaves while its pair may still read its memory or barriers
    if (warp == 1) {
        if (CTAS == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                         : "memory");
        else
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
    if (lo != nullptr) lo[e] = v - h;
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

static int make_map(CUtensorMap *tm, const float *base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                    int box_cols = BLOCK_K) {       // 32-float (128-byte) box rows, SWIZZLE_128B
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return DCNR_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): rows %lld cols %lld ld %lld", (int)r, (long long)rows,
                  (long long)cols, (long long)ld);
        return DCNR_ERR_CUDA;
    }
    return DCNR_OK;
}

static int pick_block_n(int64_t n, int cap = 256) {
    for (int bn = cap; bn >= 32; bn -= 32)       // the epilogue moves 32-column boxes
        if (n % bn == 0) return bn;
    return 0;
}
// TF32X3 takes the A operand from tensor memory: 128-column tiles (2 x 128 accumulator columns + the operand ring), single
// CTAs.  Single-pass TF32 runs 256-column tiles on 2-CTA pairs.
static int block_n_for(int precision, int64_t n) { return pick_block_n(n, precision == DCNR_PREC_TF32X3 ? 128 : 256); }

}  // namespace tc

// Epilogue warp groups: two (eight warps) for short reductions (k <= 64: the initial layer is epilogue-bound with four
// warps -- 0.34 vs 0.43-0.54 ms per 1 M x 256 x 64) and for TF32X3, whose accumulators are single-buffered since the
// correction terms got their own (the epilogue is then on the critical path of every tile: 79 -> 69 us per
// 65 536 x 256 x 256 layer, same-box A/B).  One group for single-pass TF32 (0.67 vs 0.62 ms: the mainloop sets the pace there)
// and when the tile has an odd number of 32-column chunks.
static int epi_groups_for(int block_n, int64_t k, int terms) {
    if (block_n <= 0 || (block_n / 32) % 2 != 0) return 1;
    return (k <= 64 || terms == 3) ? 2 : 1;
}

int gemm_tc_n_tiles(int64_t n, int precision, int64_t k) {      // partial row dots a FusedDot produces: column tiles x epilogue groups
    const int bn = tc::block_n_for(precision, n);
    return bn > 0 ? (int)(n / bn) * epi_groups_for(bn, k, precision == DCNR_PREC_TF32X3 ? 3 : 1) : 1;
}

bool gemm_tc_supported(int precision, bool a_kmajor, bool b_kmajor, int64_t lda, int64_t ldb, int64_t ldc, int64_t m,
                       int64_t n, int64_t k, int split_k) {
    if (precision != DCNR_PREC_TF32X3 && precision != DCNR_PREC_TF32) return false;
    if (!a_kmajor || !b_kmajor || split_k != 1) return false;
    if (k < tc::BLOCK_K || k % tc::BLOCK_K != 0 || (lda & 3) || (ldb & 3) || (ldc & 3)) return false;
    if (n % 32 != 0 || tc::pick_block_n(n) == 0 || m <= 0 || m > 0x7fffff00LL) return false;
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
    DCNR_REQUIRE(terms == 1 || B_lo != nullptr, "TF32X3 needs the pre-split weight (launch_split_tf32)");
    DCNR_REQUIRE((((uintptr_t)A | (uintptr_t)B | (uintptr_t)C | (uintptr_t)B_lo | (uintptr_t)dot_w) & 15) == 0,
                 "operands must be 16-byte aligned");
    DCNR_REQUIRE(C != nullptr || (dot_w != nullptr && dot_out != nullptr), "no output requested");
    DCNR_REQUIRE(epi.residual == nullptr || ((epi.ldr & 3) == 0 && ((uintptr_t)epi.residual & 15) == 0),
                 "residual must be 16-byte aligned");
    DCNR_REQUIRE(epi.hadamard == nullptr || ((epi.ldh & 3) == 0 && ((uintptr_t)epi.hadamard & 15) == 0),
                 "hadamard operand must be 16-byte aligned");
    DCNR_REQUIRE((epi.bias == nullptr || ((uintptr_t)epi.bias & 15) == 0) &&
                     (epi.col_scale == nullptr || ((uintptr_t)epi.col_scale & 15) == 0),
                 "bias / col_scale must be 16-byte aligned");
    Params p{};
    p.M = m; p.N_total = (int32_t)n; p.K = (int32_t)k;
    p.block_n = block_n_for(precision, n);
    p.terms = terms;
    p.epi_groups = epi_groups_for(p.block_n, k, terms);
    p.C = C; p.ldc = ldc; p.epi = epi;
    p.dot_w = dot_w; p.dot_out = dot_out;
    p.acc_cols = 32;
    while (p.acc_cols < p.block_n) p.acc_cols <<= 1;
    p.num_n_tiles = (int32_t)(n / p.block_n);

    // TF32X3: single CTAs (A in tensor memory); single-pass TF32: 2-CTA pairs whenever a pair has work
    int ctas = (terms == 1 && m > BLOCK_M) ? 2 : 1;
    int max_pairs = 0;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    auto plan = [&](int c, size_t *smem_out) {
        const int b_bytes = (p.block_n / c) * BLOCK_K * 4;
        const int stage_bytes = BLOCK_M * BLOCK_K * 4 + (terms == 3 ? 2 : 1) * b_bytes;
        // four epilogue slots per warp (residual prefetch depth 3) unless two slots buy another operand stage
        const int cap = terms == 3 ? 256 / (2 * BLOCK_K) : 6;
        auto stages_for = [&](int slots) {
            const int budget = 227 * 1024 - 1024 - kBarBytes - 4 * p.epi_groups * slots * kEpiSlotBytes;
            return std::max(1, std::min(cap, budget / stage_bytes));
        };
        // eight epilogue warps take two slots each (the same 64 KB as four warps x four slots)
        p.epi_slots = (p.epi_groups == 2 || (stages_for(2) > stages_for(4) && terms == 3)) ? 2 : 4;
        p.stages = stages_for(p.epi_slots);
        p.epi_depth = p.epi_slots - 1;
        *smem_out = (size_t)p.stages * stage_bytes + 1024 + kBarBytes + 4 * p.epi_groups * p.epi_slots * kEpiSlotBytes;
        // tensor memory: accumulator stage(s) first, then (TF32X3) the operand ring, 64 columns (hi | lo) per stage
        const int ring_cols = terms == 3 ? p.stages * 2 * BLOCK_K : 0;
        // TF32X3: the lo.hi / hi.lo terms ALWAYS get their own accumulator.  The tensor core's fp32 accumulate truncates toward
        // zero (profiles/r02_acc_probe.md: -0.8 eps per K = 8 step on same-sign sums), so every addition into the long-running
        // main sum costs up to one ulp of that sum whatever the size of the addend; with the 2^-11-sized correction terms kept
        // apart the main chain is k / 8 additions instead of 3 k / 8 and the measured rms error of a K = 256 layer drops from
        // 38 eps to 13 eps (fp32 FMA chain: 6).  Room is made by single-buffering the accumulators when necessary.
        p.corr_sep = (terms == 3 && 2 * p.acc_cols + ring_cols <= 512) ? 1 : 0;
        const int per_stage_cols = p.corr_sep ? 2 * p.acc_cols : p.acc_cols;
        p.acc_stages = 2 * per_stage_cols + ring_cols <= 512 ? 2 : 1;
        p.a_col0 = p.acc_stages * per_stage_cols;
        p.tmem_cols = 32;
        while (p.tmem_cols < p.a_col0 + ring_cols) p.tmem_cols <<= 1;
    };
    // warps 10-13 exist only when they have a role (second epilogue group): with 2-CTA pairs four idle warps per CTA
    // measured 16 % slower (tf32, 1 M x 256 x 256: 0.72 vs 0.62 ms)
    const unsigned threads = p.epi_groups == 2 ? kThreadsP : 320u;
    const bool had = epi.hadamard != nullptr, extra = threads == (unsigned)kThreadsP;
    typedef void (*KernFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, Params);
    const KernFn kern1 = had ? (extra ? k_gemm_tc<1, true, true> : k_gemm_tc<1, true, false>)
                             : (extra ? k_gemm_tc<1, false, true> : k_gemm_tc<1, false, false>);
    const KernFn kern2 = had ? (extra ? k_gemm_tc<2, true, true> : k_gemm_tc<2, true, false>)
                             : (extra ? k_gemm_tc<2, false, true> : k_gemm_tc<2, false, false>);
    size_t smem = 0;
    if (ctas == 2) {
        plan(2, &smem);
        DCNR_CUDA_CHECK(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cfg.gridDim = dim3(2 * (unsigned)sm_count(), 1, 1);
        cfg.blockDim = dim3(threads, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        if (cudaOccupancyMaxActiveClusters(&max_pairs, kern2, &cfg) != cudaSuccess || max_pairs < 1) {
            cudaGetLastError();
            ctas = 1;
        }
    }
    CUtensorMap tmA, tmBhi, tmBlo, tmR, tmC;
    DCNR_TRY(make_map(&tmA, A, m, k, lda, BLOCK_M));
    // epilogue boxes: 32 rows x 32 columns of the residual / output (rows and columns past the matrix are
    // zero-filled on load and clipped on store)
    if (C != nullptr) DCNR_TRY(make_map(&tmC, C, m, n, ldc, 32, 32));
    else tmC = tmA;                                  // never dereferenced
    if (epi.residual != nullptr) DCNR_TRY(make_map(&tmR, epi.residual, m, n, epi.ldr, 32, 32));
    else tmR = tmA;
    DCNR_TRY(make_map(&tmBhi, B, n, k, ldb, p.block_n / ctas));
    DCNR_TRY(make_map(&tmBlo, terms == 3 ? B_lo : B, n, k, ldb, p.block_n / ctas));
    p.num_m_tiles = (int32_t)ceil_div(m, BLOCK_M * ctas);
    const int64_t num_tiles = (int64_t)p.num_m_tiles * p.num_n_tiles;
    gemm_timer_before(stream, 2.0 * (double)m * (double)n * (double)k);
    if (ctas == 2) {
        cfg.gridDim = dim3(2 * (unsigned)std::min<int64_t>(num_tiles, max_pairs), 1, 1);
        DCNR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern2, tmA, tmBhi, tmBlo, tmR, tmC, p));
    } else {
        plan(1, &smem);
        DCNR_CUDA_CHECK(cudaFuncSetAttribute(kern1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned grid = (unsigned)std::min<int64_t>(num_tiles, sm_count());
        kern1<<<grid, threads, smem, stream>>>(tmA, tmBhi, tmBlo, tmR, tmC, p);
    }
    gemm_timer_after(stream);
    DCNR_LAUNCHED();
    return DCNR_OK;
}

}  // namespace dcnr
