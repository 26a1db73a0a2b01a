// Train-mode BatchNorm1d statistics / normalise+ReLU(+dropout, +residual) forward and backward,
// BN folding for eval, column sums, row padding.  All are HBM-bound passes over [m, n] fp32.
//
// Reference lines: ResBlock.forward train.py:112-122 (main.py:83-90); BatchNorm1d defaults
// (eps 1e-5, momentum 0.1, biased variance for normalisation, unbiased into running_var).
//
// Thread mapping for every column-wise kernel: the CTA owns a fixed chunk of kChunkRows rows;
// threads are (ty, tx) with tx walking float4 column quads (coalesced 16-byte accesses) and ty
// walking rows.  Every reduction is chunk-local in a fixed order, written as one partial row per
// chunk and finalised in chunk order, so results are bit-reproducible run to run.
#include "kernels.cuh"
#include "p2p.cuh"

namespace dcnr {

constexpr int kT = 1024;     // threads per CTA: at the training batch (256 chunks on 148 SMs) 256 threads left the SMs at 12-25 % occupancy

struct ColMap {
    int tx_n, ty_n;   // threads along columns (power of two) and rows
};
static inline ColMap col_map(int n) {
    int cq = n / 4, tx = 8;
    while (tx < cq && tx < kT) tx <<= 1;
    return ColMap{tx, kT / tx};
}

// ------------------------------------------------------------------------------------ statistics
// Per thread: shifted sums around the first value it sees (no catastrophic cancellation), turned
// into (count, mean, M2); threads of a chunk and then chunks are merged with Chan's formula in
// double, always in the same order.
__global__ void __launch_bounds__(kT)
k_bn_stats_partial(const float *__restrict__ z, int64_t ldz, int64_t m, int n, int tx_n, int ty_n,
                   double *__restrict__ pmean, double *__restrict__ pm2) {
    extern __shared__ __align__(16) float sm[];   // [ty_n][n] mean, [ty_n][n] m2, [ty_n] counts
    float *s_mean = sm, *s_m2 = sm + (size_t)ty_n * n;
    int *s_cnt = reinterpret_cast<int *>(s_m2 + (size_t)ty_n * n);
    const int tx = threadIdx.x % tx_n, ty = threadIdx.x / tx_n;
    const int64_t r0 = (int64_t)blockIdx.x * kChunkRows;
    const int64_t r1 = min(r0 + kChunkRows, m);
    const int cq = n >> 2;
    for (int q = tx; q < cq; q += tx_n) {
        float4 piv = make_float4(0.f, 0.f, 0.f, 0.f), s1 = piv, s2 = piv;
        int cnt = 0;
        for (int64_t r = r0 + ty; r < r1; r += ty_n) {
            float4 v = ldg4(z + r * ldz + 4 * q);
            if (cnt == 0) piv = v;
            float dx = v.x - piv.x, dy = v.y - piv.y, dz = v.z - piv.z, dw = v.w - piv.w;
            s1.x += dx; s1.y += dy; s1.z += dz; s1.w += dw;
            s2.x = fmaf(dx, dx, s2.x); s2.y = fmaf(dy, dy, s2.y); s2.z = fmaf(dz, dz, s2.z); s2.w = fmaf(dw, dw, s2.w);
            ++cnt;
        }
        const float inv = cnt > 0 ? 1.f / (float)cnt : 0.f;
        float *pm = s_mean + (size_t)ty * n + 4 * q, *pv = s_m2 + (size_t)ty * n + 4 * q;
        pm[0] = piv.x + s1.x * inv; pm[1] = piv.y + s1.y * inv; pm[2] = piv.z + s1.z * inv; pm[3] = piv.w + s1.w * inv;
        pv[0] = s2.x - s1.x * s1.x * inv; pv[1] = s2.y - s1.y * s1.y * inv;
        pv[2] = s2.z - s1.z * s1.z * inv; pv[3] = s2.w - s1.w * s1.w * inv;
        if (q == tx) s_cnt[ty] = cnt;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < n; c += kT) {
        double cn = 0.0, mean = 0.0, m2 = 0.0;
        for (int t = 0; t < ty_n; ++t) {
            const double nb = (double)s_cnt[t];
            if (nb == 0.0) continue;
            const double mb = (double)s_mean[(size_t)t * n + c], vb = (double)s_m2[(size_t)t * n + c];
            const double tot = cn + nb, delta = mb - mean;
            mean += delta * nb / tot;
            m2 += vb + delta * delta * cn * nb / tot;
            cn = tot;
        }
        pmean[(int64_t)blockIdx.x * n + c] = mean;
        pm2[(int64_t)blockIdx.x * n + c] = m2;
    }
}

// Chan merge of (count, mean, M2) triples in double; `a` absorbs `b`.
struct Moments {
    double n, mean, m2;
};
__device__ __forceinline__ void merge_moments(Moments &a, const Moments &b) {
    if (b.n == 0.0) return;
    const double tot = a.n + b.n, delta = b.mean - a.mean;
    a.mean += delta * b.n / tot;
    a.m2 += b.m2 + delta * delta * a.n * b.n / tot;
    a.n = tot;
}

// Level 1 of the chunk merge: CTA g folds chunks [g*kMergeFan, (g+1)*kMergeFan) in order, one thread per column
// (the first version folded ALL chunks in one thread per column: 1.3 ms for 4096 chunks, 90 % of the BN statistics).
constexpr int kMergeFan = 16;
__global__ void k_bn_stats_merge1(const double *__restrict__ pmean, const double *__restrict__ pm2, int64_t chunks,
                                  int64_t m, int n, double *__restrict__ gmean, double *__restrict__ gm2,
                                  double *__restrict__ gcnt) {
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int64_t k0 = (int64_t)blockIdx.x * kMergeFan, k1 = min(chunks, k0 + kMergeFan);
    double vm[kMergeFan], v2[kMergeFan];                              // all loads first: the merge chain is long enough
#pragma unroll
    for (int j = 0; j < kMergeFan; ++j) {
        vm[j] = k0 + j < k1 ? pmean[(k0 + j) * n + c] : 0.0;
        v2[j] = k0 + j < k1 ? pm2[(k0 + j) * n + c] : 0.0;
    }
    Moments acc{0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < kMergeFan; ++j)
        if (k0 + j < k1)
            merge_moments(acc, Moments{(double)min((int64_t)kChunkRows, m - (k0 + j) * kChunkRows), vm[j], v2[j]});
    gmean[(int64_t)blockIdx.x * n + c] = acc.mean;
    gm2[(int64_t)blockIdx.x * n + c] = acc.m2;
    if (c == 0) gcnt[blockIdx.x] = acc.n;
}

// Level 2: fold the groups in order (<= chunks / 16 of them) and emit mean / rstd / running statistics.
// Group k's (mean, M2, count) are gmean[k*stride + c], gm2[k*stride + c], gcnt[k*cnt_stride].  With pack_out != NULL
// the folded (mean[n] | M2[n] | count) is written as doubles instead: this rank's contribution to the SyncBN
// all-gather, whose result is folded by the same kernel with one "group" per rank (stride 2n + 2).
__global__ void k_bn_stats_finalize(const double *__restrict__ gmean, const double *__restrict__ gm2,
                                    const double *__restrict__ gcnt, int64_t groups, int64_t stride, int64_t cnt_stride,
                                    int n, float eps, float momentum, float *__restrict__ mean_out,
                                    float *__restrict__ rstd_out, float *running_mean, float *running_var, int64_t *nbt,
                                    double *__restrict__ pack_out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && nbt != nullptr && pack_out == nullptr) *nbt += 1;
    if (c >= n) return;
    Moments acc{0.0, 0.0, 0.0};
    for (int64_t k = 0; k < groups; ++k)
        merge_moments(acc, Moments{gcnt[k * cnt_stride], gmean[k * stride + c], gm2[k * stride + c]});
    if (pack_out != nullptr) {
        pack_out[c] = acc.mean;
        pack_out[n + c] = acc.m2;
        if (c == 0) pack_out[2 * n] = acc.n;
        return;
    }
    const double m = acc.n;
    const double var_b = acc.m2 / m;
    mean_out[c] = (float)acc.mean;
    rstd_out[c] = (float)(1.0 / sqrt(var_b + (double)eps));
    if (running_mean != nullptr) {
        const double var_u = m > 1.0 ? acc.m2 / (m - 1.0) : var_b;
        running_mean[c] = (float)((1.0 - (double)momentum) * (double)running_mean[c] + (double)momentum * acc.mean);
        running_var[c] = (float)((1.0 - (double)momentum) * (double)running_var[c] + (double)momentum * var_u);
    }
}

// Levels 1 and 2 in ONE launch for batches of up to 32 groups (B <= 131 072) on a single rank: thread (column, group) folds the
// group's 16 chunks exactly like k_bn_stats_merge1, the column's thread of group 0 then folds the groups in order exactly like
// k_bn_stats_finalize -- same association, bit-identical statistics, one launch instead of two per BatchNorm.
__global__ void __launch_bounds__(1024)
k_bn_stats_merge_finalize(const double *__restrict__ pmean, const double *__restrict__ pm2, int64_t chunks, int64_t m, int n,
                          int groups, float eps, float momentum, float *__restrict__ mean_out, float *__restrict__ rstd_out,
                          float *running_mean, float *running_var, int64_t *nbt) {
    __shared__ double smean[32][33], sm2[32][33], scnt[32];
    const int tx = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    if (g < groups && c < n) {
        const int64_t k0 = (int64_t)g * kMergeFan, k1 = min(chunks, k0 + kMergeFan);
        double vm[kMergeFan], v2[kMergeFan];
#pragma unroll
        for (int j = 0; j < kMergeFan; ++j) {
            vm[j] = k0 + j < k1 ? pmean[(k0 + j) * n + c] : 0.0;
            v2[j] = k0 + j < k1 ? pm2[(k0 + j) * n + c] : 0.0;
        }
        Moments acc{0.0, 0.0, 0.0};
#pragma unroll
        for (int j = 0; j < kMergeFan; ++j)
            if (k0 + j < k1)
                merge_moments(acc, Moments{(double)min((int64_t)kChunkRows, m - (k0 + j) * kChunkRows), vm[j], v2[j]});
        smean[g][tx] = acc.mean;
        sm2[g][tx] = acc.m2;
        if (tx == 0) scnt[g] = acc.n;
    }
    __syncthreads();
    if (g != 0 || c >= n) return;
    if (c == 0 && nbt != nullptr) *nbt += 1;
    Moments acc{0.0, 0.0, 0.0};
    for (int k = 0; k < groups; ++k) merge_moments(acc, Moments{scnt[k], smean[k][tx], sm2[k][tx]});
    const double cnt = acc.n;
    const double var_b = acc.m2 / cnt;
    mean_out[c] = (float)acc.mean;
    rstd_out[c] = (float)(1.0 / sqrt(var_b + (double)eps));
    if (running_mean != nullptr) {
        const double var_u = cnt > 1.0 ? acc.m2 / (cnt - 1.0) : var_b;
        running_mean[c] = (float)((1.0 - (double)momentum) * (double)running_mean[c] + (double)momentum * acc.mean);
        running_var[c] = (float)((1.0 - (double)momentum) * (double)running_var[c] + (double)momentum * var_u);
    }
}

// SyncBN forward over peer memory: ONE CTA folds this rank's groups, exchanges (mean | M2 | count) with every rank through the
// NVLink-mapped staging slots, folds the ranks in rank order (identical on every rank) and emits mean / rstd / running
// statistics -- the pack kernel, the NCCL all-gather and the fold kernel of the NCCL path in a single launch.
__global__ void __launch_bounds__(256)
k_bn_stats_finalize_sync(const P2pView *__restrict__ views, int rank, int world, const double *__restrict__ gmean,
                         const double *__restrict__ gm2, const double *__restrict__ gcnt, int64_t groups, int n, float eps,
                         float momentum, float *__restrict__ mean_out, float *__restrict__ rstd_out, float *running_mean,
                         float *running_var, int64_t *nbt) {
    const P2pExchange x(views, rank, world);
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        Moments acc{0.0, 0.0, 0.0};
        for (int64_t k = 0; k < groups; ++k) merge_moments(acc, Moments{gcnt[k], gmean[k * n + c], gm2[k * n + c]});
        for (int r = 0; r < world; ++r) {
            double *dst = x.send_slot(r);
            dst[c] = acc.mean;
            dst[n + c] = acc.m2;
            if (c == 0) dst[2 * n] = acc.n;
        }
    }
    x.publish_and_wait();
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        Moments acc{0.0, 0.0, 0.0};
        for (int r = 0; r < world; ++r) {
            const volatile double *src = x.recv_slot(r);
            merge_moments(acc, Moments{src[2 * n], src[c], src[n + c]});
        }
        const double m = acc.n;
        const double var_b = acc.m2 / m;
        mean_out[c] = (float)acc.mean;
        rstd_out[c] = (float)(1.0 / sqrt(var_b + (double)eps));
        if (running_mean != nullptr) {
            const double var_u = m > 1.0 ? acc.m2 / (m - 1.0) : var_b;
            running_mean[c] = (float)((1.0 - (double)momentum) * (double)running_mean[c] + (double)momentum * acc.mean);
            running_var[c] = (float)((1.0 - (double)momentum) * (double)running_var[c] + (double)momentum * var_u);
        }
    }
    if (threadIdx.x == 0 && nbt != nullptr) *nbt += 1;
    x.finish();
}

// SyncBN backward over peer memory: exchange (sum dy | sum dy*xhat | rows), add the ranks in rank order, leave the global sums
// and 1 / global rows in `sums` for k_bn_bwd_apply -- pack + all-gather + unpack in one launch.
__global__ void __launch_bounds__(256)
k_bn_bwd_sync(const P2pView *__restrict__ views, int rank, int world, float *__restrict__ sums, int n2, double rows) {
    const P2pExchange x(views, rank, world);
    for (int c = threadIdx.x; c <= n2; c += blockDim.x) {
        const double v = c < n2 ? (double)sums[c] : rows;
        for (int r = 0; r < world; ++r) x.send_slot(r)[c] = v;
    }
    x.publish_and_wait();
    for (int c = threadIdx.x; c <= n2; c += blockDim.x) {
        double acc = 0.0;
        for (int r = 0; r < world; ++r) acc += x.recv_slot(r)[c];
        sums[c] = c < n2 ? (float)acc : (float)(1.0 / acc);      // sums[n2] = 1 / global batch rows
    }
    x.finish();
}

// SyncBN backward: pack this rank's (sum dy | sum dy*xhat | rows) as doubles; after the all-gather, add the ranks in
// rank order and leave the global sums and 1 / global rows for k_bn_bwd_apply.
__global__ void k_bn_bwd_pack(const float *__restrict__ sums, int n2, double rows, double *__restrict__ pack) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n2) pack[c] = (double)sums[c];
    if (c == 0) pack[n2] = rows;
}
__global__ void k_bn_bwd_unpack(const double *__restrict__ all, int world, int n2, float *__restrict__ sums) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > n2) return;
    double acc = 0.0;
    for (int r = 0; r < world; ++r) acc += all[(int64_t)r * (n2 + 2) + c];
    sums[c] = c < n2 ? (float)acc : (float)(1.0 / acc);      // sums[n2] = 1 / global batch rows
}

int64_t bn_scratch_floats(int64_t m, int32_t n) {
    const int64_t chunks = std::max<int64_t>(ceil_div(m, kChunkRows), 1);
    const int64_t groups = ceil_div(chunks, kMergeFan);
    // two double rows per chunk, two double rows + a count per merge group, 2 float rows of sums (+ 1 / rows), and the
    // SyncBN exchange buffers: (2n + 2) doubles for this rank and for each of up to kMaxWorld ranks
    return chunks * 4 * (int64_t)n + groups * (4 * (int64_t)n + 2) + 4 * (int64_t)n + 64 +
           2 * (int64_t)(kMaxWorld + 1) * (2 * (int64_t)n + 2);
}

int launch_bn_stats(const float *z, int64_t ldz, int64_t m, int32_t n, float eps, float momentum, float *mean,
                    float *rstd, float *running_mean, float *running_var, int64_t *nbt, float *scratch,
                    cudaStream_t stream, const void *comm) {
    DCNR_REQUIRE(n % 4 == 0 && n >= 4 && (ldz & 3) == 0, "bn: n=%d / ld must be multiples of 4", n);
    DCNR_REQUIRE(m >= 1, "bn: empty batch");
    const int64_t chunks = ceil_div(m, kChunkRows);
    double *pmean = reinterpret_cast<double *>(scratch);
    double *pm2 = pmean + chunks * n;
    const ColMap cm = col_map(n);
    const size_t smem = (size_t)cm.ty_n * n * 2 * sizeof(float) + cm.ty_n * sizeof(int);
    if (smem > 48 * 1024)
        DCNR_CUDA_CHECK(cudaFuncSetAttribute(k_bn_stats_partial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_bn_stats_partial<<<(unsigned)chunks, kT, smem, stream>>>(z, ldz, m, n, cm.tx_n, cm.ty_n, pmean, pm2);
    DCNR_LAUNCHED();
    const int64_t groups = ceil_div(chunks, kMergeFan);
    if (comm_world(comm) <= 1 && groups <= 32) {
        k_bn_stats_merge_finalize<<<(unsigned)ceil_div(n, 32), (unsigned)(32 * groups), 0, stream>>>(
            pmean, pm2, chunks, m, n, (int)groups, eps, momentum, mean, rstd, running_mean, running_var, nbt);
        DCNR_LAUNCHED();
        return DCNR_OK;
    }
    double *gmean = pm2 + chunks * n, *gm2 = gmean + groups * n, *gcnt = gm2 + groups * n;
    dim3 g1((unsigned)groups, (unsigned)ceil_div(n, 128));
    k_bn_stats_merge1<<<g1, 128, 0, stream>>>(pmean, pm2, chunks, m, n, gmean, gm2, gcnt);
    DCNR_LAUNCHED();
    const int world = comm_world(comm);
    if (world > 1) {
        DCNR_REQUIRE(world <= kMaxWorld, "data-parallel group larger than %d ranks", kMaxWorld);
        if (const P2pView *views = comm_p2p_views(comm); views != nullptr && 2 * n + 2 <= kP2pCap) {
            k_bn_stats_finalize_sync<<<1, 256, 0, stream>>>(views, comm_rank(comm), world, gmean, gm2, gcnt, groups, n, eps, momentum,
                                                           mean, rstd, running_mean, running_var, nbt);
            DCNR_LAUNCHED();
            return DCNR_OK;
        }
        double *mine = gcnt + groups, *all = mine + (2 * n + 2);              // exchange buffers (see bn_scratch_floats)
        k_bn_stats_finalize<<<(unsigned)ceil_div(n, 128), 128, 0, stream>>>(gmean, gm2, gcnt, groups, n, 1, n, eps, momentum,
                                                                          nullptr, nullptr, nullptr, nullptr, nullptr, mine);
        DCNR_LAUNCHED();
        DCNR_TRY(comm_allgather_f64(comm, mine, all, 2 * n + 2, stream));
        k_bn_stats_finalize<<<(unsigned)ceil_div(n, 128), 128, 0, stream>>>(all, all + n, all + 2 * n, world, 2 * n + 2,
                                                                          2 * n + 2, n, eps, momentum, mean, rstd,
                                                                          running_mean, running_var, nbt, nullptr);
        DCNR_LAUNCHED();
        return DCNR_OK;
    }
    k_bn_stats_finalize<<<(unsigned)ceil_div(n, 128), 128, 0, stream>>>(gmean, gm2, gcnt, groups, n, 1, n, eps, momentum, mean,
                                                                      rstd, running_mean, running_var, nbt, nullptr);
    DCNR_LAUNCHED();
    return DCNR_OK;
}

// ------------------------------------------------------------------------------------ Philox4x32-10
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }

// ------------------------------------------------------------------------------------ forward apply
// Purely elementwise, so its CTAs need not follow the 256-row reduction chunks: 64 rows per CTA put four times as many loads
// in flight at the training batch (B = 65 536: 1 024 CTAs instead of 256 on 148 SMs).
constexpr int kActRows = 64;
__global__ void __launch_bounds__(kT)
k_bn_act_fwd(const float *__restrict__ z, int64_t ldz, const float *__restrict__ mean,
             const float *__restrict__ rstd, const float *__restrict__ gamma, const float *__restrict__ beta,
             const float *__restrict__ residual, int64_t ldr, const uint8_t *__restrict__ keep, float drop_p,
             uint64_t seed, uint32_t layer_tag, float *__restrict__ out, int64_t ldo, int64_t m, int n, int tx_n,
             int ty_n, const uint64_t *__restrict__ seed_step) {
    const int tx = threadIdx.x % tx_n, ty = threadIdx.x / tx_n;
    const int64_t r0 = (int64_t)blockIdx.x * kActRows;
    const int64_t r1 = min(r0 + kActRows, m);
    const int cq = n >> 2;
    const bool philox = keep == nullptr && drop_p > 0.f;
    const float post = (keep != nullptr || philox) ? 1.f / (1.f - drop_p) : 1.f;
    if (seed_step != nullptr) seed += __ldg(seed_step);           // device-side step counter (CUDA-graph replays)
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    for (int q = tx; q < cq; q += tx_n) {
        const float4 mu = ldg4(mean + 4 * q), rs = ldg4(rstd + 4 * q), ga = ldg4(gamma + 4 * q), be = ldg4(beta + 4 * q);
        for (int64_t r = r0 + ty; r < r1; r += ty_n) {
            float4 v = ldg4(z + r * ldz + 4 * q);
            v.x = fmaf((v.x - mu.x) * rs.x, ga.x, be.x);
            v.y = fmaf((v.y - mu.y) * rs.y, ga.y, be.y);
            v.z = fmaf((v.z - mu.z) * rs.z, ga.z, be.z);
            v.w = fmaf((v.w - mu.w) * rs.w, ga.w, be.w);
            if (residual != nullptr) {
                const float4 rr = ldg4(residual + r * ldr + 4 * q);
                v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
            }
            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
            if (keep != nullptr) {
                const uchar4 k4 = *reinterpret_cast<const uchar4 *>(keep + r * (int64_t)n + 4 * q);
                v.x = k4.x ? v.x * post : 0.f; v.y = k4.y ? v.y * post : 0.f;
                v.z = k4.z ? v.z * post : 0.f; v.w = k4.w ? v.w * post : 0.f;
            } else if (philox) {
                const uint4 rnd = philox4x32_10(make_uint4((uint32_t)r, (uint32_t)((uint64_t)r >> 32), (uint32_t)q, layer_tag), key);
                v.x = u01(rnd.x) >= drop_p ? v.x * post : 0.f; v.y = u01(rnd.y) >= drop_p ? v.y * post : 0.f;
                v.z = u01(rnd.z) >= drop_p ? v.z * post : 0.f; v.w = u01(rnd.w) >= drop_p ? v.w * post : 0.f;
            }
            st4(out + r * ldo + 4 * q, v);
        }
    }
}

int launch_bn_act_fwd(const float *z, int64_t ldz, const float *mean, const float *rstd, const float *gamma,
                      const float *beta, const float *residual, int64_t ldr, const uint8_t *keep, float drop_p,
                      uint64_t seed, uint32_t layer_tag, float *out, int64_t ldo, int64_t m, int32_t n,
                      cudaStream_t stream, const uint64_t *seed_step) {
    DCNR_REQUIRE(n % 4 == 0 && (ldz & 3) == 0 && (ldo & 3) == 0 && (residual == nullptr || (ldr & 3) == 0),
                 "bn_act: n / ld must be multiples of 4");
    DCNR_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "dropout p must be in [0,1)");
    if (m <= 0) return DCNR_OK;
    const ColMap cm = col_map(n);
    k_bn_act_fwd<<<(unsigned)ceil_div(m, kActRows), kT, 0, stream>>>(z, ldz, mean, rstd, gamma, beta, residual, ldr,
                                                                     keep, drop_p, seed, layer_tag, out, ldo, m, n,
                                                                     cm.tx_n, cm.ty_n, seed_step);
    DCNR_LAUNCHED();
    return DCNR_OK;
}

// ------------------------------------------------------------------------------------ backward
// pass 1: partial[chunk][0][c] = sum dy, partial[chunk][1][c] = sum dy*xhat over the chunk's rows
__global__ void __launch_bounds__(kT)
k_bn_bwd_reduce(const float *__restrict__ g, int64_t ldg, const float *__restrict__ out, int64_t ldo,
                const float *__restrict__ z, int64_t ldz, const float *__restrict__ mean,
                const float *__restrict__ rstd, float post_scale, int64_t m, int n, int tx_n, int ty_n,
                float *__restrict__ partials) {
    extern __shared__ __align__(16) float sm[];   // [ty_n][2][n]
    const int tx = threadIdx.x % tx_n, ty = threadIdx.x / tx_n;
    const int64_t r0 = (int64_t)blockIdx.x * kChunkRows;
    const int64_t r1 = min(r0 + kChunkRows, m);
    const int cq = n >> 2;
    for (int q = tx; q < cq; q += tx_n) {
        const float4 mu = ldg4(mean + 4 * q), rs = ldg4(rstd + 4 * q);
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        for (int64_t r = r0 + ty; r < r1; r += ty_n) {
            const float4 gv = ldg4(g + r * ldg + 4 * q), ov = ldg4(out + r * ldo + 4 * q), zv = ldg4(z + r * ldz + 4 * q);
            const float dx = ov.x > 0.f ? gv.x * post_scale : 0.f, dy = ov.y > 0.f ? gv.y * post_scale : 0.f;
            const float dzv = ov.z > 0.f ? gv.z * post_scale : 0.f, dw = ov.w > 0.f ? gv.w * post_scale : 0.f;
            a.x += dx; a.y += dy; a.z += dzv; a.w += dw;
            b.x = fmaf(dx, (zv.x - mu.x) * rs.x, b.x); b.y = fmaf(dy, (zv.y - mu.y) * rs.y, b.y);
            b.z = fmaf(dzv, (zv.z - mu.z) * rs.z, b.z); b.w = fmaf(dw, (zv.w - mu.w) * rs.w, b.w);
        }
        st4(sm + ((size_t)ty * 2 + 0) * n + 4 * q, a);
        st4(sm + ((size_t)ty * 2 + 1) * n + 4 * q, b);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * n; c += kT) {
        float s = 0.f;
        for (int t = 0; t < ty_n; ++t) s += sm[(size_t)t * 2 * n + c];
        partials[(int64_t)blockIdx.x * 2 * n + c] = s;
    }
}

// pass 2: dz = gamma*rstd*(dy - sum_dy/m - xhat*sum_dyx/m); optional dy_out; per-chunk column sums of dz
__global__ void __launch_bounds__(kT)
k_bn_bwd_apply(const float *g, int64_t ldg, const float *__restrict__ out, int64_t ldo,
               const float *__restrict__ z, int64_t ldz, const float *__restrict__ mean,
               const float *__restrict__ rstd, const float *__restrict__ gamma, const float *__restrict__ sums,
               float post_scale, float *dz, int64_t lddz, float *dy_out, int64_t lddy, int64_t m, int n, int tx_n,
               int ty_n, float *__restrict__ partials, int global_rows) {
    extern __shared__ __align__(16) float sm[];   // [ty_n][n]
    const int tx = threadIdx.x % tx_n, ty = threadIdx.x / tx_n;
    // chunks in DESCENDING order: the reduce pass just streamed g / out / z front to back, so the tail of the batch is what
    // the L2 still holds (3 x 67 MB against 126 MB at B = 65 536)
    const int64_t chunk = (int64_t)gridDim.x - 1 - blockIdx.x;
    const int64_t r0 = chunk * kChunkRows;
    const int64_t r1 = min(r0 + kChunkRows, m);
    const int cq = n >> 2;
    const float inv_m = global_rows ? __ldg(sums + 2 * n) : 1.f / (float)m;     // SyncBN: sums and rows are global
    for (int q = tx; q < cq; q += tx_n) {
        const float4 mu = ldg4(mean + 4 * q), rs = ldg4(rstd + 4 * q), ga = ldg4(gamma + 4 * q);
        const float4 sa = ldg4(sums + 4 * q), sb = ldg4(sums + n + 4 * q);
        const float4 k = make_float4(ga.x * rs.x, ga.y * rs.y, ga.z * rs.z, ga.w * rs.w);
        const float4 ma = make_float4(sa.x * inv_m, sa.y * inv_m, sa.z * inv_m, sa.w * inv_m);
        const float4 mb = make_float4(sb.x * inv_m, sb.y * inv_m, sb.z * inv_m, sb.w * inv_m);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int64_t r = r0 + ty; r < r1; r += ty_n) {
            const float4 gv = *reinterpret_cast<const float4 *>(g + r * ldg + 4 * q);
            const float4 ov = ldg4(out + r * ldo + 4 * q), zv = ldg4(z + r * ldz + 4 * q);
            float4 dy;
            dy.x = ov.x > 0.f ? gv.x * post_scale : 0.f; dy.y = ov.y > 0.f ? gv.y * post_scale : 0.f;
            dy.z = ov.z > 0.f ? gv.z * post_scale : 0.f; dy.w = ov.w > 0.f ? gv.w * post_scale : 0.f;
            float4 d;
            d.x = k.x * (dy.x - ma.x - (zv.x - mu.x) * rs.x * mb.x);
            d.y = k.y * (dy.y - ma.y - (zv.y - mu.y) * rs.y * mb.y);
            d.z = k.z * (dy.z - ma.z - (zv.z - mu.z) * rs.z * mb.z);
            d.w = k.w * (dy.w - ma.w - (zv.w - mu.w) * rs.w * mb.w);
            if (dy_out != nullptr) st4(dy_out + r * lddy + 4 * q, dy);
            st4(dz + r * lddz + 4 * q, d);
            acc.x += d.x; acc.y += d.y; acc.z += d.z; acc.w += d.w;
        }
        st4(sm + (size_t)ty * n + 4 * q, acc);
    }
    if (partials == nullptr) return;
    __syncthreads();
    for (int c = threadIdx.x; c < n; c += kT) {
        float s = 0.f;
        for (int t = 0; t < ty_n; ++t) s += sm[(size_t)t * n + c];
        partials[chunk * n + c] = s;
    }
}

int launch_bn_act_bwd(const float *g, int64_t ldg, const float *out, int64_t ldo, const float *z, int64_t ldz,
                      const float *mean, const float *rstd, const float *gamma, float post_scale, float *dz,
                      int64_t lddz, float *dy_out, int64_t lddy, float *dgamma, float *dbeta, float *dbias,
                      int64_t m, int32_t n, float *scratch, cudaStream_t stream, const void *comm) {
    DCNR_REQUIRE(n % 4 == 0 && (ldg & 3) == 0 && (ldo & 3) == 0 && (ldz & 3) == 0 && (lddz & 3) == 0 &&
                     (dy_out == nullptr || (lddy & 3) == 0),
                 "bn_act_bwd: n / ld must be multiples of 4");
    if (m <= 0) return DCNR_OK;
    const int64_t chunks = ceil_div(m, kChunkRows);
    const ColMap cm = col_map(n);
    float *partials = scratch;                          // [chunks][2n]
    float *sums = scratch + chunks * 2 * (int64_t)n;    // [2][n]
    size_t smem = (size_t)cm.ty_n * 2 * n * sizeof(float);
    if (smem > 48 * 1024)
        DCNR_CUDA_CHECK(cudaFuncSetAttribute(k_bn_bwd_reduce, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_bn_bwd_reduce<<<(unsigned)chunks, kT, smem, stream>>>(g, ldg, out, ldo, z, ldz, mean, rstd, post_scale, m, n,
                                                          cm.tx_n, cm.ty_n, partials);
    DCNR_LAUNCHED();
    SegPtrs seg;
    memset(&seg, 0, sizeof(seg));
    seg.n = 4;
    seg.out[0] = sums;       seg.offset[0] = 0; seg.len[0] = n;
    seg.out[1] = sums + n;   seg.offset[1] = n; seg.len[1] = n;
    seg.out[2] = dbeta;      seg.offset[2] = 0; seg.len[2] = n;
    seg.out[3] = dgamma;     seg.offset[3] = n; seg.len[3] = n;
    DCNR_TRY(launch_sum_partials(partials, chunks, 2 * (int64_t)n, seg, stream));
    const int world = comm_world(comm);
    if (world > 1) {            // dgamma / dbeta stay this rank's sums (they are all-reduced with the other gradients)
        DCNR_REQUIRE(world <= kMaxWorld, "data-parallel group larger than %d ranks", kMaxWorld);
        const P2pView *views = comm_p2p_views(comm);
        double *mine = reinterpret_cast<double *>(scratch + bn_scratch_floats(m, n) - 2 * (int64_t)(kMaxWorld + 1) * (2 * (int64_t)n + 2));
        double *all = mine + (2 * n + 2);
        if (views != nullptr && 2 * n + 2 <= kP2pCap) {
            k_bn_bwd_sync<<<1, 256, 0, stream>>>(views, comm_rank(comm), world, sums, 2 * n, (double)m);
            DCNR_LAUNCHED();
        } else {
            k_bn_bwd_pack<<<(unsigned)ceil_div(2 * n + 1, 128), 128, 0, stream>>>(sums, 2 * n, (double)m, mine);
            DCNR_LAUNCHED();
            DCNR_TRY(comm_allgather_f64(comm, mine, all, 2 * n + 2, stream));
            k_bn_bwd_unpack<<<(unsigned)ceil_div(2 * n + 1, 128), 128, 0, stream>>>(all, world, 2 * n, sums);
            DCNR_LAUNCHED();
        }
    }
    smem = (size_t)cm.ty_n * n * sizeof(float);
    if (smem > 48 * 1024)
        DCNR_CUDA_CHECK(cudaFuncSetAttribute(k_bn_bwd_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    float *p2 = dbias != nullptr ? partials : nullptr;   // pass-1 partials are dead once sums exist
    k_bn_bwd_apply<<<(unsigned)chunks, kT, smem, stream>>>(g, ldg, out, ldo, z, ldz, mean, rstd, gamma, sums, post_scale,
                                                         dz, lddz, dy_out, lddy, m, n, cm.tx_n, cm.ty_n, p2, world > 1 ? 1 : 0);
    DCNR_LAUNCHED();
    if (dbias != nullptr) {
        memset(&seg, 0, sizeof(seg));
        seg.n = 1;
        seg.out[0] = dbias; seg.offset[0] = 0; seg.len[0] = n;
        DCNR_TRY(launch_sum_partials(partials, chunks, n, seg, stream));
    }
    return DCNR_OK;
}

// ------------------------------------------------------------------------------------ column sums
__global__ void __launch_bounds__(kT)
k_colsum_partial(const float *__restrict__ a, int64_t lda, int64_t m, int n, int tx_n, int ty_n,
                 float *__restrict__ partials) {
    extern __shared__ __align__(16) float sm[];   // [ty_n][n]
    const int tx = threadIdx.x % tx_n, ty = threadIdx.x / tx_n;
    const int64_t r0 = (int64_t)blockIdx.x * kChunkRows;
    const int64_t r1 = min(r0 + kChunkRows, m);
    const int cq = n >> 2;
    for (int q = tx; q < cq; q += tx_n) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int64_t r = r0 + ty; r < r1; r += ty_n) {
            const float4 v = ldg4(a + r * lda + 4 * q);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        st4(sm + (size_t)ty * n + 4 * q, acc);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < n; c += kT) {
        float s = 0.f;
        for (int t = 0; t < ty_n; ++t) s += sm[(size_t)t * n + c];
        partials[(int64_t)blockIdx.x * n + c] = s;
    }
}

int launch_colsum(const float *a, int64_t lda, int64_t m, int32_t n, float *out, float *scratch, cudaStream_t stream) {
    DCNR_REQUIRE(n % 4 == 0 && (lda & 3) == 0, "colsum: n / ld must be multiples of 4");
    if (m <= 0) return DCNR_OK;
    const int64_t chunks = ceil_div(m, kChunkRows);
    const ColMap cm = col_map(n);
    const size_t smem = (size_t)cm.ty_n * n * sizeof(float);
    if (smem > 48 * 1024)
        DCNR_CUDA_CHECK(cudaFuncSetAttribute(k_colsum_partial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_colsum_partial<<<(unsigned)chunks, kT, smem, stream>>>(a, lda, m, n, cm.tx_n, cm.ty_n, scratch);
    DCNR_LAUNCHED();
    SegPtrs seg;
    memset(&seg, 0, sizeof(seg));
    seg.n = 1;
    seg.out[0] = out; seg.offset[0] = 0; seg.len[0] = n;
    return launch_sum_partials(scratch, chunks, n, seg, stream);
}

// ------------------------------------------------------------------------------------ eval fold, padding
__global__ void k_bn_fold(const float *__restrict__ gamma, const float *__restrict__ beta,
                          const float *__restrict__ rm, const float *__restrict__ rv,
                          const float *__restrict__ lin_bias, float eps, float *__restrict__ scale,
                          float *__restrict__ shift, int n) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const double s = (double)gamma[c] / sqrt((double)rv[c] + (double)eps);
    const double b = lin_bias != nullptr ? (double)lin_bias[c] : 0.0;
    scale[c] = (float)s;
    shift[c] = (float)((double)beta[c] + (b - (double)rm[c]) * s);
}

int launch_bn_fold(const float *gamma, const float *beta, const float *rm, const float *rv, const float *lin_bias,
                   float eps, float *scale, float *shift, int32_t n, cudaStream_t stream) {
    k_bn_fold<<<(unsigned)ceil_div(n, 128), 128, 0, stream>>>(gamma, beta, rm, rv, lin_bias, eps, scale, shift, n);
    DCNR_LAUNCHED();
    return DCNR_OK;
}

__global__ void k_pad_rows(const float *__restrict__ src, int64_t lds, float *__restrict__ dst, int64_t ldd, int rows,
                           int cols, int cols_pad) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)rows * cols_pad) return;
    const int r = (int)(e / cols_pad), c = (int)(e % cols_pad);
    dst[(int64_t)r * ldd + c] = c < cols ? src[(int64_t)r * lds + c] : 0.f;
}

__global__ void k_inc_u64(uint64_t *p) { *p += 1; }
int launch_inc_u64(uint64_t *p, cudaStream_t stream) {
    k_inc_u64<<<1, 1, 0, stream>>>(p);
    DCNR_LAUNCHED();
    return DCNR_OK;
}

int launch_pad_rows(const float *src, int64_t lds, float *dst, int64_t ldd, int32_t rows, int32_t cols,
                    int32_t cols_pad, cudaStream_t stream) {
    const int64_t total = (int64_t)rows * cols_pad;
    if (total <= 0) return DCNR_OK;
    k_pad_rows<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(src, lds, dst, ldd, rows, cols, cols_pad);
    DCNR_LAUNCHED();
    return DCNR_OK;
}

}  // namespace dcnr

// ------------------------------------------------------------------------------------------------
using namespace dcnr;

extern "C" int64_t dcnr_bn_scratch_bytes(int64_t m, int32_t n) { return round_up(bn_scratch_floats(m, n) * 4, 256); }

extern "C" int dcnr_bn_stats(const float *z, int64_t ldz, int64_t m, int32_t n, float eps, float momentum, float *mean,
                             float *rstd, float *running_mean, float *running_var, int64_t *num_batches_tracked,
                             void *scratch, int64_t scratch_bytes, dcnr_stream_t stream) {
    DCNR_REQUIRE(z && mean && rstd && scratch, "null argument");
    if (scratch_bytes < dcnr_bn_scratch_bytes(m, n)) {
        set_error("bn scratch too small");
        return DCNR_ERR_WORKSPACE;
    }
    return launch_bn_stats(z, ldz, m, n, eps, momentum, mean, rstd, running_mean, running_var, num_batches_tracked,
                           reinterpret_cast<float *>(scratch), as_stream(stream));
}

extern "C" int dcnr_bn_act_fwd(const float *z, int64_t ldz, const float *mean, const float *rstd, const float *gamma,
                               const float *beta, const float *residual, int64_t ldr, const uint8_t *keep, float drop_p,
                               uint64_t seed, uint32_t layer_tag, float *out, int64_t ldo, int64_t m, int32_t n,
                               dcnr_stream_t stream) {
    DCNR_REQUIRE(z && mean && rstd && gamma && beta && out, "null argument");
    return launch_bn_act_fwd(z, ldz, mean, rstd, gamma, beta, residual, ldr, keep, drop_p, seed, layer_tag, out, ldo, m,
                             n, as_stream(stream), nullptr);
}

extern "C" int dcnr_bn_act_bwd(const float *g, int64_t ldg, const float *out, int64_t ldo, const float *z, int64_t ldz,
                               const float *mean, const float *rstd, const float *gamma, float post_scale, float *dz,
                               int64_t lddz, float *dy_out, int64_t lddy, float *dgamma, float *dbeta, float *dbias,
                               int64_t m, int32_t n, void *scratch, int64_t scratch_bytes, dcnr_stream_t stream) {
    DCNR_REQUIRE(g && out && z && mean && rstd && gamma && dz && scratch, "null argument");
    if (scratch_bytes < dcnr_bn_scratch_bytes(m, n)) {
        set_error("bn scratch too small");
        return DCNR_ERR_WORKSPACE;
    }
    return launch_bn_act_bwd(g, ldg, out, ldo, z, ldz, mean, rstd, gamma, post_scale, dz, lddz, dy_out, lddy, dgamma,
                             dbeta, dbias, m, n, reinterpret_cast<float *>(scratch), as_stream(stream));
}
