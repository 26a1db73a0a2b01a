// Fused eval-mode deep tower (sm_100a):  logits = final_dot( ResBlocks( initial_deep_layer(x0) ) ) + cross half
//
// Reference: DCN_RecSys.forward in eval() -- train.py:161-170 / main.py:120-127 with ResBlock.forward main.py:83-90 --
// i.e. h0 = x0 W0^T + b0;  per block  t = relu(BN1(h W1^T + b1)),  h = relu(BN2(t W2^T + b2) + h);  logit = wf[:H].h + ...
// (BatchNorm folded into a per-column scale / shift, dropout is the identity).
//
// Round 1 ran this as five tcgen05 GEMM launches per 2^20-row chunk, every [rows, 256] fp32 activation making a round trip
// through HBM (11.8 KB per row against 144 B of algorithmic traffic).  Here ONE persistent kernel keeps a 128-row tile on the
// SM for the whole tower: activations live in tensor memory, only x0 (256 B / row) comes in and one logit (4 B / row) goes out.
//
//   * Arithmetic: kind::f16 MMAs on an error-compensated fp16 split.  Every fp32 operand v is written v = hi + lo with
//     hi = rn_f16(v), lo = rn_f16(v - hi) (|v - hi - lo| <= 2^-23 |v| inside the fp16 range) and the product is accumulated
//     as hi.hi + lo.hi + hi.lo in the fp32 TMEM accumulator: the same three-term scheme as tf32x3, at twice the tensor-pipe
//     rate and half the operand bytes.  Range: activations are pre-scaled by `sa` (a power of two), each layer's weight by a
//     power of two that puts its largest entry in [512, 1024); both are undone exactly by the epilogue scale.  An activation
//     beyond the fp16 range sets bit 1 of *flags (the caller re-runs the batch on the tf32x3 path).  terms == 1 is the bf16
//     mode (stated tolerance, not parity): hi only, bf16 operands, no range limit.
//   * Tensor memory (512 columns) = two 256-column buffers.  Layer l accumulates into buffer l & 1 while its A operand is
//     read from buffer (l - 1) & 1: the epilogue of layer l - 1 converts the accumulator IN PLACE, 16 columns at a time,
//     into the fp16 hi / lo operand of layer l (16 fp32 columns -> 8 columns of packed hi pairs + 8 of lo pairs = one K = 16
//     MMA step), so the MMAs of layer l start on K-step s as soon as group s of layer l - 1 is converted (TS-form MMA: A
//     from tensor memory).
//   * The residual input of a block is kept as fp32 in shared memory (128 KB, [col/4][row] float4: conflict-free).
//   * Weights stream from L2 through a 3-slot x 32 KB TMA ring (pre-split fp16 hi / lo, prepared once per call by
//     k_tower_prep; one slot = the [256 x 32] hi | lo tile of a K = 32 step); the x0 tile enters the same ring as a
//     shared-memory A operand (SS-form MMA) for the initial layer.
//   * Single CTAs, one per SM.  A 2-CTA form (tcgen05.mma.cta_group::2 sharing every weight tile between two SMs) was built
//     and measured first: 636 M rows/s against 689 M rows/s for single CTAs (P0, 4 Mi rows) -- at one tile per SM the kernel is
//     not shared-memory bound, and the pair pays cluster-scope barrier latency on every operand hand-over -- so it was removed.
//
// Warp roles (640 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2-3 x0 loader, warps 4-19 epilogue (the four warps
// w, w + 4, w + 8, w + 12 share a TMEM lane quadrant and take every fourth 16-column group).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace dcnr {
namespace tw {

using namespace ptx;

constexpr int H = 256;                 // hidden width this kernel is built for (accumulator = 256 TMEM columns)
constexpr int BM = 128;                // rows per CTA tile
constexpr int SLOT_BYTES = 32768;      // one ring slot: the [256 x 32] hi | lo weight tile of a K = 32 step (or an x0 chunk)
constexpr int NSLOT = 3;
constexpr int MAXL = 9;                // 1 + 2 * 4 ResBlocks
constexpr int RES_BYTES = BM * H * 4;
constexpr int kThreads = 640;
constexpr int kNumBars = 2 * NSLOT + 2 + 8 + 1 + 8;
constexpr int kSmemBytes = NSLOT * SLOT_BYTES + RES_BYTES + 2 * H * 4 + kNumBars * 8 + 16;
constexpr float kRangeLimit = 60000.f;   // fp16 max is 65504

struct Params {
    const float *x0;          // [M, ldx0] fp32, columns K0.. are never read
    int64_t ldx0, M;
    const float *vec;         // [L][2][H] (scale, shift) then wfs[H], from k_tower_prep
    const float *logit_cross; // [M] or NULL
    const float *bf;          // 1 float or NULL
    float *out;               // [M]
    int32_t *flags;           // may be NULL
    float sa;
    int32_t K0, L, terms, bf16, num_tiles;
    uint32_t relu_mask, resin_mask, resout_mask;      // bit l: layer l applies ReLU / adds the saved residual / saves its output
};

struct Ring {                 // running slot / phase of the shared ring (every role walks the same sequence)
    int s = 0;
    uint32_t ph = 0;
    __device__ __forceinline__ void next() {
        if (++s == NSLOT) { s = 0; ph ^= 1u; }
    }
    __device__ __forceinline__ void skip(int n) {
        const int t = s + n;
        ph ^= (uint32_t)(t / NSLOT) & 1u;
        s = t % NSLOT;
    }
};

__global__ void __launch_bounds__(kThreads, 1)
k_tower_eval(const __grid_constant__ CUtensorMap tmW, Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *ring = smem;
    float *res = reinterpret_cast<float *>(smem + NSLOT * SLOT_BYTES);
    float *vecs = res + BM * H;                                  // [2][H]: scale, shift of the layer being drained
    uint64_t *bars = reinterpret_cast<uint64_t *>(vecs + 2 * H);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + kNumBars);
    const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8 * NSLOT, accfull0 = empty0 + 8 * NSLOT,
                   aready0 = accfull0 + 16, accfree0 = aready0 + 64, xgo0 = accfree0 + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if ((smem_u32(smem) & 1023u) != 0) __trap();                 // swizzled operand tiles need the 1024-byte alignment
    if (threadIdx.x == 0) {
        for (int i = 0; i < NSLOT; ++i) {
            mbar_init(full0 + 8 * i, 1);                         // producer's arrive.expect_tx / the loader's arrive
            mbar_init(empty0 + 8 * i, 1);                        // tcgen05.commit
        }
        mbar_init(accfull0, 1);
        mbar_init(accfull0 + 8, 1);
        for (int c = 0; c < 8; ++c) mbar_init(aready0 + 8 * c, 8);           // the eight warps that convert 32-column chunk c
        mbar_init(accfree0, 16);
        for (int i = 0; i < 8; ++i) mbar_init(xgo0 + 8 * i, 1);             // producer -> loader: x0 slot i of this tile is free
        mbar_init_fence();
    }
    if (warp == 1) tmem_alloc<1>(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Bounded wait.  A pipeline bug must not hang the GPU: after ~2^21 failed polls the waiter records who it is (flags[1] =
    // warp | site << 8 | parity << 16 | block << 20, flags[2] = `info`; the caller may pass flags in pinned host memory so
    // the record survives the trap) and traps.
    int32_t *const dbg = p.flags;
    auto wait_x = [dbg](uint32_t bar, uint32_t parity, int site, int info) {
        uint32_t ok = 0;
        for (uint32_t spins = 0; !ok; ++spins) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
            if (!ok && spins > (1u << 21)) {
                if (dbg != nullptr && (threadIdx.x & 31) == 0) {
                    if (atomicCAS(dbg + 1, 0, (int)((threadIdx.x >> 5) | (site << 8) | (parity << 16) | (blockIdx.x << 20))) == 0)
                        dbg[2] = info;
                    __threadfence_system();
                }
                __trap();
            }
        }
    };
    const int L = p.L, n0 = p.K0 / 32;

    if (warp == 0) {
        // ---------------- TMA producer: the weight tile of every K = 32 step, in layer / K order ----------------
        Ring r;
        const uint32_t tx_bytes = (uint32_t)((p.terms == 3 ? 2 : 1) * (SLOT_BYTES / 2));
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
            for (int l = 0; l < L; ++l) {
                const int nchunks = l == 0 ? n0 : 8;
                for (int c = 0; c < nchunks; ++c) {
                    if (l == 0) {
                        // The x0 slot of this K chunk is filled by the loader warps, but the ring bookkeeping stays here: the loader
                        // would otherwise wait for a slot a whole tile (many ring revolutions) ahead, and a parity wait is only
                        // meaningful at most one phase ahead.  The producer waits for the slot in sequence and hands it over.
                        wait_x(empty0 + 8 * r.s, r.ph ^ 1u, 8, (t << 16) | (l << 8) | c);
                        if (elect_one()) mbar_arrive(xgo0 + 8 * c);
                        __syncwarp();
                        r.next();
                    }
                    wait_x(empty0 + 8 * r.s, r.ph ^ 1u, 1, (t << 16) | (l << 8) | c);
                    if (elect_one()) {
                        const uint32_t dst = smem_u32(ring + r.s * SLOT_BYTES);
                        mbar_expect_tx(full0 + 8 * r.s, tx_bytes);
                        tma_load_2d(dst, &tmW, 32 * c, l * 2 * H, full0 + 8 * r.s);
                        if (p.terms == 3) tma_load_2d(dst + SLOT_BYTES / 2, &tmW, 32 * c, l * 2 * H + H, full0 + 8 * r.s);
                    }
                    __syncwarp();
                    r.next();
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer (all lanes loop, one elected lane issues) ----------------
        // One ring slot = one K = 32 step = two K = 16 MMA groups, and the converted operand is signalled per 32 columns, so the
        // issuing warp has two barriers per six MMAs (768 clk of tensor work).  Measured history of this loop (P0, 4 Mi rows):
        // one barrier pair per K = 16 group 6.1-6.4 ms (the loop cost ~520 clk per group against 384 clk of MMAs and bounded the
        // kernel); per K = 32 group 5.6 ms; probing the next chunk's barriers between the MMA groups (mbarrier.test_wait) 5.7 ms,
        // no gain; two N = 128 passes per layer (to overlap the first conversion with the second pass) 6.7 ms -- twice the
        // barrier traffic again.
        const uint32_t idesc = idesc_f16(p.bf16 ? 1 : 0, BM, H);
        const bool three = p.terms == 3;
        Ring r;
        uint32_t g = 0, hc = 0, it = 0;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
            for (int l = 0; l < L; ++l, ++g) {
                const uint32_t d = tmem_base + (g & 1u) * 256u, a0 = tmem_base + ((g & 1u) ^ 1u) * 256u;
                if (l == 1 && it > 0) {                      // every epilogue warp has drained the previous tile's last layer
                    wait_x(accfree0, (it - 1u) & 1u, 2, (t << 16) | (l << 8));
                    tc_fence_after();
                }
                const int nchunks = l == 0 ? n0 : 8;
                for (int c = 0; c < nchunks; ++c) {
                    int sx = 0;
                    if (l == 0) {                            // A = the x0 chunk in shared memory (hi | lo, 64-byte rows): SS form
                        sx = r.s;
                        wait_x(full0 + 8 * r.s, r.ph, 3, (t << 16) | (l << 8) | c);
                        r.next();
                    } else {                                 // A = chunk c of the previous layer's output in tensor memory: TS form
                        wait_x(aready0 + 8 * c, hc & 1u, 4, (t << 16) | (l << 8) | c);
                    }
                    const int sw = r.s;
                    wait_x(full0 + 8 * r.s, r.ph, 5, (t << 16) | (l << 8) | c);
                    r.next();
                    tc_fence_after();
                    const uint32_t wb = smem_u32(ring + sw * SLOT_BYTES), xb = smem_u32(ring + sx * SLOT_BYTES);
#pragma unroll
                    for (int sp = 0; sp < 2; ++sp) {
                        if (elect_one()) {
                            const uint64_t db_hi = smem_desc_kmajor(wb + sp * 32, 64), db_lo = smem_desc_kmajor(wb + SLOT_BYTES / 2 + sp * 32, 64);
                            const uint32_t acc = (uint32_t)((c | sp) != 0);
                            if (l == 0) {
                                const uint64_t da_hi = smem_desc_kmajor(xb + sp * 32, 64), da_lo = smem_desc_kmajor(xb + 8192 + sp * 32, 64);
                                mma_f16_ss<1>(d, da_hi, db_hi, idesc, acc);
                                if (three) {
                                    mma_f16_ss<1>(d, da_lo, db_hi, idesc, 1u);
                                    mma_f16_ss<1>(d, da_hi, db_lo, idesc, 1u);
                                }
                            } else {
                                const uint32_t ta_hi = a0 + (uint32_t)(32 * c + 16 * sp), ta_lo = ta_hi + 8u;
                                mma_f16_ts<1>(d, ta_hi, db_hi, idesc, acc);
                                if (three) {
                                    mma_f16_ts<1>(d, ta_lo, db_hi, idesc, 1u);
                                    mma_f16_ts<1>(d, ta_hi, db_lo, idesc, 1u);
                                }
                            }
                            if (sp == 1) {
                                mma_commit<1>(empty0 + 8 * sw);
                                if (l == 0) mma_commit<1>(empty0 + 8 * sx);
                                if (c == nchunks - 1) mma_commit<1>(accfull0 + 8 * (g & 1u));
                            }
                        }
                        __syncwarp();
                    }
                }
                if (l > 0) ++hc;
            }
        }
    } else if (warp < 4) {
        // ---------------- x0 loader: fp32 rows -> scaled fp16 hi / lo, K-major SWIZZLE_64B operand tile in a ring slot ----------------
        const int t2 = threadIdx.x - 64, rsub = t2 >> 3, c4 = t2 & 7;    // 8 lanes cover the 128 bytes of one row's K chunk
        Ring r;
        float mx = 0.f;
        uint32_t nt = 0;                                         // tiles done by this CTA: xgo[i] completes once per tile
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++nt) {
            const int64_t m0 = (int64_t)t * BM;
            for (int i = 0; i < n0; ++i) {
                float4 v[16];
#pragma unroll
                for (int st = 0; st < 16; ++st) {
                    const int64_t row = m0 + st * 8 + rsub;
                    v[st] = row < p.M ? ldg4(p.x0 + row * p.ldx0 + 32 * i + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                const int sx = r.s;
                wait_x(xgo0 + 8 * i, nt & 1u, 6, (t << 16) | i);
                r.skip(2);
                uint8_t *slot = ring + sx * SLOT_BYTES;
#pragma unroll
                for (int st = 0; st < 16; ++st) {
                    const int rr = st * 8 + rsub;
                    const float f0 = v[st].x * p.sa, f1 = v[st].y * p.sa, f2 = v[st].z * p.sa, f3 = v[st].w * p.sa;
                    mx = fmaxf(mx, fmaxf(fmaxf(fabsf(f0), fabsf(f1)), fmaxf(fabsf(f2), fabsf(f3))));
                    const uint32_t off = (uint32_t)(rr * 64 + ((((c4 >> 1) ^ ((rr >> 1) & 3)) << 4) | ((c4 & 1) << 3)));
                    uint2 hi, lo;
                    if (p.bf16) {
                        const __nv_bfloat162 h01 = __floats2bfloat162_rn(f0, f1), h23 = __floats2bfloat162_rn(f2, f3);
                        hi.x = *reinterpret_cast<const uint32_t *>(&h01);
                        hi.y = *reinterpret_cast<const uint32_t *>(&h23);
                        lo.x = lo.y = 0u;
                    } else {
                        const __half2 h01 = __floats2half2_rn(f0, f1), h23 = __floats2half2_rn(f2, f3);
                        const float2 b01 = __half22float2(h01), b23 = __half22float2(h23);
                        const __half2 l01 = __floats2half2_rn(f0 - b01.x, f1 - b01.y), l23 = __floats2half2_rn(f2 - b23.x, f3 - b23.y);
                        hi.x = *reinterpret_cast<const uint32_t *>(&h01);
                        hi.y = *reinterpret_cast<const uint32_t *>(&h23);
                        lo.x = *reinterpret_cast<const uint32_t *>(&l01);
                        lo.y = *reinterpret_cast<const uint32_t *>(&l23);
                    }
                    *reinterpret_cast<uint2 *>(slot + off) = hi;
                    *reinterpret_cast<uint2 *>(slot + 8192 + off) = lo;
                }
                fence_proxy_async_smem();
                asm volatile("bar.sync 2, 64;" ::: "memory");
                if (t2 == 0) mbar_arrive(full0 + 8 * sx);
            }
            r.skip((L - 1) * 8);
        }
        if (!p.bf16 && mx > kRangeLimit && p.flags != nullptr) atomicOr(p.flags, 2);
    } else {
        // ---------------- epilogue warps 4..19: thread = one tile row, 16 accumulator columns (one K = 16 step) at a time ----------------
        const int ew = warp - 4, grp = ew >> 2, quad = warp & 3;
        const int row = quad * 32 + lane, et = threadIdx.x - 128;
        const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
        float4 *res4 = reinterpret_cast<float4 *>(res);
        float wreg[4];                                           // wfs of this warp's groups: column 16 (grp + 4 cc) + (lane & 15)
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) wreg[cc] = __ldg(p.vec + (int64_t)L * 2 * H + 16 * (grp + 4 * cc) + (lane & 15));
        float pre = __ldg(p.vec + et);                           // [scale | shift] of the next layer to drain, one float per thread
        const float bias_f = p.bf != nullptr ? __ldg(p.bf) : 0.f;
        uint32_t g = 0;
        float mx = 0.f;
        for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
            const int64_t m = (int64_t)t * BM + row;
            for (int l = 0; l < L; ++l, ++g) {
                wait_x(accfull0 + 8 * (g & 1u), (g >> 1) & 1u, 7, (t << 16) | (l << 8));
                tc_fence_after();
                // layer switch: everyone is done with the previous layer's vectors; publish this layer's, prefetch the next
                asm volatile("bar.sync 1, 512;" ::: "memory");
                vecs[et] = pre;
                asm volatile("bar.sync 1, 512;" ::: "memory");
                pre = __ldg(p.vec + (int64_t)(l + 1 == L ? 0 : l + 1) * 2 * H + et);
                const bool relu = (p.relu_mask >> l) & 1u, res_in = (p.resin_mask >> l) & 1u, res_out = (p.resout_mask >> l) & 1u;
                const bool last = l + 1 == L;
                const uint32_t tbuf = tmem_base + (g & 1u) * 256u + lane_addr;
                float dot = 0.f;
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    const int ks = grp + 4 * cc;                 // 16-column group = K step of the next layer

                    uint32_t v[16];
                    tmem_ld16(tbuf + 16u * ks, v);
                    float y[16];
                    const float4 *sv = reinterpret_cast<const float4 *>(vecs + 16 * ks), *hv = reinterpret_cast<const float4 *>(vecs + H + 16 * ks);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 s4 = sv[q], h4 = hv[q];
                        y[4 * q] = fmaf(__uint_as_float(v[4 * q]), s4.x, h4.x);
                        y[4 * q + 1] = fmaf(__uint_as_float(v[4 * q + 1]), s4.y, h4.y);
                        y[4 * q + 2] = fmaf(__uint_as_float(v[4 * q + 2]), s4.z, h4.z);
                        y[4 * q + 3] = fmaf(__uint_as_float(v[4 * q + 3]), s4.w, h4.w);
                    }
                    if (res_in) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 r4 = res4[(4 * ks + q) * BM + row];
                            y[4 * q] += r4.x; y[4 * q + 1] += r4.y; y[4 * q + 2] += r4.z; y[4 * q + 3] += r4.w;
                        }
                    }
                    if (relu) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) y[j] = fmaxf(y[j], 0.f);
                    }
                    if (res_out) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            res4[(4 * ks + q) * BM + row] = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
                    }
                    if (!last) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) mx = fmaxf(mx, fabsf(y[j]));
                        uint32_t o[16];                          // columns 0..7: packed hi pairs, 8..15: packed lo pairs
                        if (p.bf16) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const __nv_bfloat162 h2 = __floats2bfloat162_rn(y[2 * j], y[2 * j + 1]);
                                o[j] = *reinterpret_cast<const uint32_t *>(&h2);
                                o[8 + j] = 0u;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const __half2 h2 = __floats2half2_rn(y[2 * j], y[2 * j + 1]);
                                const float2 b2 = __half22float2(h2);
                                const __half2 l2 = __floats2half2_rn(y[2 * j] - b2.x, y[2 * j + 1] - b2.y);
                                o[j] = *reinterpret_cast<const uint32_t *>(&h2);
                                o[8 + j] = *reinterpret_cast<const uint32_t *>(&l2);
                            }
                        }
                        tmem_st16(tbuf + 16u * ks, o);
                        tmem_st_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(aready0 + 8 * (ks >> 1));
                    } else {
                        const float wv = cc == 0 ? wreg[0] : (cc == 1 ? wreg[1] : (cc == 2 ? wreg[2] : wreg[3]));
#pragma unroll
                        for (int j = 0; j < 16; ++j) dot = fmaf(y[j], __shfl_sync(0xffffffffu, wv, j), dot);
                    }
                }
                if (last) {
                    // groups 1..3 hand their part of the row dot to group 0 through a residual cell they have already consumed
                    // (the first float4 of their first 16-column group)
                    if (grp != 0) res[(4 * grp * BM + row) * 4] = dot;
                    asm volatile("bar.sync 3, 512;" ::: "memory");
                    if (grp == 0 && m < p.M) {
                        const float cross = p.logit_cross != nullptr ? __ldg(p.logit_cross + m) : 0.f;
                        p.out[m] = ((dot + res[(4 * BM + row) * 4]) + (res[(8 * BM + row) * 4] + res[(12 * BM + row) * 4])) + cross + bias_f;
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(accfree0);
                }
            }
        }
        if (!p.bf16 && mx > kRangeLimit && p.flags != nullptr) atomicOr(p.flags, 2);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<1>(tmem_base, 512);
}

// ---- weight / vector preparation: one CTA per layer ------------------------------------------------------------------------
struct PrepArgs {
    const float *w[MAXL];         // layer weight [H, K_l], row stride ldw[l]
    const float *lin_b[MAXL];     // nn.Linear bias [H]
    const float *gamma[MAXL], *beta[MAXL], *rm[MAXL], *rv[MAXL];      // BatchNorm of the layer (NULL for the initial layer)
    int32_t ldw[MAXL], kcols[MAXL];
    const float *wf;              // final_linear.weight[0:H]
    uint16_t *pack;               // [L][2 (hi, lo)][H][H] sixteen-bit
    float *vec;                   // [L][2][H] + wfs[H]
    float eps, sa;
    int32_t L, bf16;
};

__global__ void __launch_bounds__(1024)
k_tower_prep(PrepArgs a) {
    __shared__ float red[32];
    __shared__ int sh_e;
    const int l = blockIdx.x, t = threadIdx.x;
    const float *W = a.w[l];
    const int ldw = a.ldw[l], K = a.kcols[l];
    float mx = 0.f;
    for (int idx = t; idx < H * K; idx += 1024) mx = fmaxf(mx, fabsf(__ldg(W + (int64_t)(idx / K) * ldw + idx % K)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((t & 31) == 0) red[t >> 5] = mx;
    __syncthreads();
    if (t == 0) {
        float m = 0.f;
        for (int i = 0; i < 32; ++i) m = fmaxf(m, red[i]);
        int e = 0;
        if (!a.bf16 && m > 0.f && m < 1e30f) {      // largest |w| * 2^e in [512, 1024): hi <= 1024 and lo stays a normal fp16 for
            int x;                                  // every entry down to 2^-12 of the largest one
            frexpf(m, &x);
            e = 10 - x;
        }
        sh_e = e;
    }
    __syncthreads();
    const int e = sh_e;
    uint16_t *hi = a.pack + (int64_t)(l * 2) * H * H, *lo = hi + (int64_t)H * H;
    for (int idx = t; idx < H * H; idx += 1024) {
        const int n = idx >> 8, k = idx & 255;
        const float w = k < K ? ldexpf(__ldg(W + (int64_t)n * ldw + k), e) : 0.f;
        if (a.bf16) {
            const __nv_bfloat16 h = __float2bfloat16_rn(w);
            hi[idx] = *reinterpret_cast<const uint16_t *>(&h);
            lo[idx] = 0;
        } else {
            const __half h = __float2half_rn(w);
            const __half q = __float2half_rn(w - __half2float(h));
            hi[idx] = *reinterpret_cast<const uint16_t *>(&h);
            lo[idx] = *reinterpret_cast<const uint16_t *>(&q);
        }
    }
    if (t < H) {
        double s = 1.0, sh = a.lin_b[l] != nullptr ? (double)a.lin_b[l][t] : 0.0;
        if (a.gamma[l] != nullptr) {                 // eval BatchNorm folded in double (6.4e-7 vs the reference)
            s = (double)a.gamma[l][t] / sqrt((double)a.rv[l][t] + (double)a.eps);
            sh = (double)a.beta[l][t] + (sh - (double)a.rm[l][t]) * s;
        }
        a.vec[(int64_t)l * 2 * H + t] = (float)ldexp(s, -e);
        a.vec[(int64_t)l * 2 * H + H + t] = (float)(sh * (double)a.sa);
        if (l == a.L - 1) a.vec[(int64_t)a.L * 2 * H + t] = a.wf[t] / a.sa;
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

}  // namespace tw

bool tower_eval_supported(const dcnr_dims *d) {
    return d->hidden == tw::H && d->in_dim_pad % 32 == 0 && d->in_dim_pad >= 32 && d->in_dim_pad <= 256 &&
           d->n_res >= 1 && 1 + 2 * d->n_res <= tw::MAXL;
}

int64_t tower_pack_bytes(const dcnr_dims *d) {
    const int64_t L = 1 + 2 * d->n_res;
    return round_up(L * 2 * tw::H * tw::H * 2, 256) + round_up((L * 2 * tw::H + tw::H) * 4, 256);
}

// pack: tower_pack_bytes(); the fp16 (bf16) hi / lo weights first, the per-layer scale / shift vectors + wf after them
int launch_tower_prepare(const dcnr_dims *d, const dcnr_params *p, void *pack, int precision, cudaStream_t stream) {
    using namespace tw;
    DCNR_REQUIRE(tower_eval_supported(d), "model shape not supported by the fused tower");
    const int L = 1 + 2 * d->n_res;
    PrepArgs a;
    memset(&a, 0, sizeof(a));
    a.w[0] = p->w0; a.lin_b[0] = p->b0; a.ldw[0] = d->in_dim; a.kcols[0] = d->in_dim;
    for (int r = 0; r < d->n_res; ++r) {
        const int l1 = 1 + 2 * r, l2 = 2 + 2 * r;
        a.w[l1] = p->res_w1[r]; a.lin_b[l1] = p->res_b1[r]; a.gamma[l1] = p->res_g1[r]; a.beta[l1] = p->res_be1[r];
        a.rm[l1] = p->res_rm1[r]; a.rv[l1] = p->res_rv1[r];
        a.w[l2] = p->res_w2[r]; a.lin_b[l2] = p->res_b2[r]; a.gamma[l2] = p->res_g2[r]; a.beta[l2] = p->res_be2[r];
        a.rm[l2] = p->res_rm2[r]; a.rv[l2] = p->res_rv2[r];
        a.ldw[l1] = a.ldw[l2] = H; a.kcols[l1] = a.kcols[l2] = H;
    }
    a.wf = p->wf;
    a.pack = reinterpret_cast<uint16_t *>(pack);
    a.vec = reinterpret_cast<float *>(reinterpret_cast<char *>(pack) + round_up((int64_t)L * 2 * H * H * 2, 256));
    a.eps = d->bn_eps;
    a.sa = 16.f;
    a.L = L;
    a.bf16 = precision == DCNR_PREC_BF16 ? 1 : 0;
    k_tower_prep<<<L, 1024, 0, stream>>>(a);
    DCNR_LAUNCHED();
    return DCNR_OK;
}

int launch_tower_eval(const dcnr_dims *d, const float *x0, int64_t ldx0, const float *logit_cross, const float *bf,
                      const void *pack, float *out, int64_t M, int32_t *flags, int precision, int options,
                      cudaStream_t stream) {
    using namespace tw;
    DCNR_REQUIRE(tower_eval_supported(d), "model shape not supported by the fused tower");
    DCNR_REQUIRE(precision == DCNR_PREC_FP16X3 || precision == DCNR_PREC_BF16, "fused tower runs fp16x3 or bf16");
    DCNR_REQUIRE(ldx0 >= d->in_dim_pad && (ldx0 & 3) == 0 && ((uintptr_t)x0 & 15) == 0, "x0 must be 16-byte aligned, ld %% 4 == 0");
    if (M <= 0) return DCNR_OK;
    const int L = 1 + 2 * d->n_res;
    Params p;
    memset(&p, 0, sizeof(p));
    p.x0 = x0; p.ldx0 = ldx0; p.M = M;
    p.vec = reinterpret_cast<const float *>(reinterpret_cast<const char *>(pack) + round_up((int64_t)L * 2 * H * H * 2, 256));
    p.logit_cross = logit_cross; p.bf = bf; p.out = out; p.flags = flags;
    p.sa = 16.f;
    p.K0 = d->in_dim_pad; p.L = L;
    p.bf16 = precision == DCNR_PREC_BF16 ? 1 : 0;
    p.terms = p.bf16 ? 1 : 3;
    p.num_tiles = (int32_t)ceil_div(M, BM);
    for (int l = 1; l < L; ++l) {
        p.relu_mask |= 1u << l;
        if ((l & 1) == 0) p.resin_mask |= 1u << l;               // second layer of a block adds the block input
    }
    for (int l = 0; l + 1 < L; l += 2) p.resout_mask |= 1u << l;  // outputs that are the input of a following block
    const int max_ctas = options >> 8;           // options >> 8 caps the grid (tests: many tiles per CTA)
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return DCNR_ERR_CUDA;
    }
    CUtensorMap tmW;
    {
        cuuint64_t dims[2] = {(cuuint64_t)H, (cuuint64_t)L * 2 * H};
        cuuint64_t strides[1] = {(cuuint64_t)H * 2};
        cuuint32_t box[2] = {32, (cuuint32_t)H};             // one K = 32 step of all 256 weight rows: 64-byte rows, SWIZZLE_64B
        cuuint32_t estr[2] = {1, 1};
        CUresult r = fn(&tmW, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void *>(pack), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled failed (%d) for the tower weights", (int)r);
            return DCNR_ERR_CUDA;
        }
    }
    const int sms = max_ctas > 0 ? std::min(max_ctas, sm_count()) : sm_count();
    // algorithmic flops of the launch: initial layer on the UNPADDED input width, 2R hidden layers, deep half of the final dot
    gemm_timer_before(stream, (double)M * (2.0 * d->in_dim * H + (double)(L - 1) * 2.0 * H * H + 2.0 * H));
    DCNR_CUDA_CHECK(cudaFuncSetAttribute(k_tower_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    const unsigned grid = (unsigned)std::min<int64_t>(p.num_tiles, sms);
    k_tower_eval<<<grid, kThreads, kSmemBytes, stream>>>(tmW, p);
    gemm_timer_after(stream);
    DCNR_LAUNCHED();
    return DCNR_OK;
}

}  // namespace dcnr
