// Final dot product (deep half), its backward, BCE-with-logits, id range check.
//
// Reference lines: final_linear over cat([deep, cross]) + squeeze train.py:169-170 (main.py:126-127);
// nn.BCEWithLogitsLoss() train.py:206,224.  The concat is never materialised: the cross half of
// the dot comes out of the gather+cross kernel, the deep half is computed here.
#include "kernels.cuh"

namespace dcnr {

constexpr int kT = 256;

__global__ void __launch_bounds__(kT)
k_rowdot_fwd(const float *__restrict__ a, int64_t lda, const float *__restrict__ w, const float *__restrict__ extra,
             const float *__restrict__ bf, float *__restrict__ out, int64_t m, int n) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * kT + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * kT) >> 5;
    const int cq = n >> 2;
    const float bias = bf != nullptr ? __ldg(bf) : 0.f;
    for (int64_t r = warp; r < m; r += n_warps) {
        float p = 0.f;
        for (int q = lane; q < cq; q += 32) {
            const float4 v = ldg4(a + r * lda + 4 * q), wv = ldg4(w + 4 * q);
            p = fmaf(v.x, wv.x, p); p = fmaf(v.y, wv.y, p); p = fmaf(v.z, wv.z, p); p = fmaf(v.w, wv.w, p);
        }
        p = warp_sum(p);
        if (lane == 0) out[r] = p + (extra != nullptr ? extra[r] : 0.f) + bias;
    }
}

int launch_rowdot_fwd(const float *a, int64_t lda, const float *w, const float *extra, const float *bf, float *out,
                      int64_t m, int32_t n, cudaStream_t stream) {
    DCNR_REQUIRE(n % 4 == 0 && (lda & 3) == 0, "rowdot: n / ld must be multiples of 4");
    if (m <= 0) return DCNR_OK;
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(m, kT / 32), (int64_t)sm_count() * 8);
    k_rowdot_fwd<<<grid, kT, 0, stream>>>(a, lda, w, extra, bf, out, m, n);
    DCNR_LAUNCHED();
    return DCNR_OK;
}

// out[m] = sum_t parts[t][m] + extra[m] + bf   (the partial row dots of the fused GEMM epilogue)
__global__ void k_combine_logits(const float *__restrict__ parts, int n_parts, int64_t m, const float *__restrict__ extra,
                                 const float *__restrict__ bf, float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    float acc = 0.f;
    for (int t = 0; t < n_parts; ++t) acc += parts[(int64_t)t * m + i];
    out[i] = acc + (extra != nullptr ? extra[i] : 0.f) + (bf != nullptr ? __ldg(bf) : 0.f);
}

int launch_combine_logits(const float *parts, int n_parts, int64_t m, const float *extra, const float *bf, float *out,
                          cudaStream_t stream) {
    if (m <= 0) return DCNR_OK;
    k_combine_logits<<<(unsigned)ceil_div(m, 256), 256, 0, stream>>>(parts, n_parts, m, extra, bf, out);
    DCNR_LAUNCHED();
    return DCNR_OK;
}

constexpr int kTB = 1024;     // threads per 256-row chunk of the backward pass (occupancy at the training batch, like bn.cu)
// dh[b,:] = dlogit[b] * w ;  partial[chunk][c] = sum_b dlogit[b]*a[b,c] ;  partial[chunk][n] = sum_b dlogit[b]
__global__ void __launch_bounds__(kTB)
k_rowdot_bwd(const float *__restrict__ dlogit, const float *__restrict__ a, int64_t lda, const float *__restrict__ w,
             float *__restrict__ dh, int64_t lddh, int64_t m, int n, int tx_n, int ty_n, float *__restrict__ partials) {
    extern __shared__ __align__(16) float sm[];   // [ty_n][n + 4]
    const int tx = threadIdx.x % tx_n, ty = threadIdx.x / tx_n;
    const int64_t r0 = (int64_t)blockIdx.x * kChunkRows;
    const int64_t r1 = min(r0 + kChunkRows, m);
    const int cq = n >> 2, np = n + 4;
    for (int q = tx; q < cq; q += tx_n) {
        const float4 wv = ldg4(w + 4 * q);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float sd = 0.f;
        for (int64_t r = r0 + ty; r < r1; r += ty_n) {
            const float d = __ldg(dlogit + r);
            const float4 v = ldg4(a + r * lda + 4 * q);
            acc.x = fmaf(d, v.x, acc.x); acc.y = fmaf(d, v.y, acc.y); acc.z = fmaf(d, v.z, acc.z); acc.w = fmaf(d, v.w, acc.w);
            sd += d;
            if (dh != nullptr) st4(dh + r * lddh + 4 * q, make_float4(d * wv.x, d * wv.y, d * wv.z, d * wv.w));
        }
        st4(sm + (size_t)ty * np + 4 * q, acc);
        if (q == 0) sm[(size_t)ty * np + n] = sd;
    }
    __syncthreads();
    for (int c = threadIdx.x; c <= n; c += kTB) {
        float s = 0.f;
        for (int t = 0; t < ty_n; ++t) s += sm[(size_t)t * np + c];
        partials[(int64_t)blockIdx.x * np + c] = s;
    }
}

int launch_rowdot_bwd(const float *dlogit, const float *a, int64_t lda, const float *w, float *dh, int64_t lddh,
                      float *dw, float *dbf, int64_t m, int32_t n, float *scratch, cudaStream_t stream) {
    DCNR_REQUIRE(n % 4 == 0 && (lda & 3) == 0 && (lddh & 3) == 0, "rowdot_bwd: n / ld must be multiples of 4");
    if (m <= 0) return DCNR_OK;
    const int64_t chunks = ceil_div(m, kChunkRows);
    int tx_n = 8;
    while (tx_n < n / 4 && tx_n < kTB) tx_n <<= 1;
    const int ty_n = kTB / tx_n;
    const size_t smem = (size_t)ty_n * (n + 4) * sizeof(float);
    if (smem > 48 * 1024)
        DCNR_CUDA_CHECK(cudaFuncSetAttribute(k_rowdot_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_rowdot_bwd<<<(unsigned)chunks, kTB, smem, stream>>>(dlogit, a, lda, w, dh, lddh, m, n, tx_n, ty_n, scratch);
    DCNR_LAUNCHED();
    SegPtrs seg;
    memset(&seg, 0, sizeof(seg));
    seg.n = 2;
    seg.out[0] = dw;  seg.offset[0] = 0; seg.len[0] = n;
    seg.out[1] = dbf; seg.offset[1] = n; seg.len[1] = 1;
    return launch_sum_partials(scratch, chunks, n + 4, seg, stream);
}

// loss_i = max(z,0) - y z + log1p(exp(-|z|)) ; dz_i = (sigmoid(z) - y)/B
__global__ void __launch_bounds__(kT)
k_bce(const float *__restrict__ logits, const float *__restrict__ labels, int64_t m, float *__restrict__ partials,
      float *__restrict__ grad) {
    __shared__ float red[kT / 32];
    const float inv = 1.f / (float)m;
    float acc = 0.f;
    // fixed assignment of rows to threads: the partial of a CTA depends only on (m, gridDim)
    for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < m; i += (int64_t)gridDim.x * kT) {
        const float z = logits[i], y = labels[i];
        const float e = expf(-fabsf(z));
        acc += (fmaxf(z, 0.f) - y * z + log1pf(e)) * inv;
        if (grad != nullptr) {
            const float sig = z >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
            grad[i] = (sig - y) * inv;
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < kT / 32; ++w) s += red[w];
        partials[blockIdx.x] = s;
    }
}

__global__ void k_check_ids(GatherArgs ga, int64_t B, int32_t *flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    bool bad = false;
    for (int s = 0; s < ga.n_seg; ++s) {
        const int64_t id = ga.seg[s].ids[i * ga.seg[s].id_stride];
        bad |= (uint64_t)id >= (uint64_t)ga.seg[s].rows;
    }
    if (bad) atomicExch(flag, 1);
}

}  // namespace dcnr

using namespace dcnr;

extern "C" int dcnr_rowdot_fwd(const float *a, int64_t lda, const float *w, const float *extra, const float *bf,
                               float *out, int64_t m, int32_t n, dcnr_stream_t stream) {
    DCNR_REQUIRE(a && w && out, "null argument");
    return launch_rowdot_fwd(a, lda, w, extra, bf, out, m, n, as_stream(stream));
}

extern "C" int dcnr_bce_with_logits(const float *logits, const float *labels, int64_t batch, float *loss,
                                    float *grad_logits, float *scratch, dcnr_stream_t stream) {
    DCNR_REQUIRE(logits && labels && loss && scratch && batch > 0, "null argument / empty batch");
    const int grid = (int)std::min<int64_t>(ceil_div(batch, kT), 1024);
    k_bce<<<grid, kT, 0, as_stream(stream)>>>(logits, labels, batch, scratch, grad_logits);
    DCNR_LAUNCHED();
    SegPtrs seg;
    memset(&seg, 0, sizeof(seg));
    seg.n = 1;
    seg.out[0] = loss; seg.offset[0] = 0; seg.len[0] = 1;
    return launch_sum_partials(scratch, grid, 1, seg, as_stream(stream));
}

extern "C" int dcnr_check_ids(const dcnr_dims *dims, const dcnr_batch *batch, int32_t *err_flag, dcnr_stream_t stream) {
    DCNR_REQUIRE(dims && batch && err_flag, "null argument");
    GatherArgs ga;
    DCNR_TRY(make_gather_args(dims, nullptr, batch, &ga));
    cudaStream_t st = as_stream(stream);
    DCNR_CUDA_CHECK(cudaMemsetAsync(err_flag, 0, sizeof(int32_t), st));
    if (batch->batch > 0) {
        k_check_ids<<<(unsigned)ceil_div(batch->batch, 256), 256, 0, st>>>(ga, batch->batch, err_flag);
        DCNR_LAUNCHED();
    }
    int32_t h = 0;
    DCNR_CUDA_CHECK(cudaMemcpyAsync(&h, err_flag, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    DCNR_CUDA_CHECK(cudaStreamSynchronize(st));
    if (h != 0) {
        set_error("index out of range in self");   // torch's IndexError message for nn.Embedding
        return DCNR_ERR_INDEX;
    }
    return DCNR_OK;
}

// out[j] = base[j] + sum_n v[n] * W[n * ldw + j]   (double accumulation, n ascending: deterministic).
// Used for the initial layer's bias gradient: db0 = colsum(dz1 W1 + dy2) = colsum(dz1) W1 + colsum(dy2) by linearity, so the
// batch sum never runs over GEMM outputs (whose tensor-core rounding errors do not average out over a cancelling sum).
namespace dcnr {
// out[j] = base[j] + sum_n v[n] * W[n, j].  32 columns per CTA; the rows are dealt to 32 lanes-of-rows (thread ty takes rows
// ty, ty + 32, ...), partial sums in double, folded in ty order: a fixed association, ~3 us for 256 x 256 (one thread per
// column walking all rows was a 256-long dependent chain of L2 loads: 56 us).
__global__ void __launch_bounds__(1024)
k_vecmat_add(const float *__restrict__ v, const float *__restrict__ W, int64_t ldw, int32_t n_rows, int32_t n_cols,
             const float *__restrict__ base, float *__restrict__ out) {
    __shared__ double sh[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + tx;
    double acc = 0.0;
    if (j < n_cols)
        for (int n = ty; n < n_rows; n += 32) acc += (double)__ldg(v + n) * (double)__ldg(W + (int64_t)n * ldw + j);
    sh[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && j < n_cols) {
        double t = base != nullptr ? (double)base[j] : 0.0;
#pragma unroll
        for (int y = 0; y < 32; ++y) t += sh[y][tx];
        out[j] = (float)t;
    }
}

int launch_vecmat_add(const float *v, const float *W, int64_t ldw, int32_t n_rows, int32_t n_cols, const float *base,
                      float *out, cudaStream_t stream) {
    k_vecmat_add<<<(unsigned)ceil_div(n_cols, 32), 1024, 0, stream>>>(v, W, ldw, n_rows, n_cols, base, out);
    DCNR_LAUNCHED();
    return DCNR_OK;
}

// out[c] = mean over the first min(m, 512) rows of a[:, c] (an ESTIMATE of the column mean for the centred weight gradient:
// any vector works there, a close one removes the cancellation).  Same shape as k_vecmat_add.
__global__ void __launch_bounds__(1024)
k_col_mean_sample(const float *__restrict__ a, int64_t lda, int32_t rows, int32_t n, float *__restrict__ out) {
    __shared__ double sh[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + tx;
    double acc = 0.0;
    if (j < n)
        for (int r = ty; r < rows; r += 32) acc += (double)__ldg(a + (int64_t)r * lda + j);
    sh[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && j < n) {
        double t = 0.0;
#pragma unroll
        for (int y = 0; y < 32; ++y) t += sh[y][tx];
        out[j] = (float)(t / (double)rows);
    }
}

int launch_col_mean_sample(const float *a, int64_t lda, int64_t m, int32_t n, float *out, cudaStream_t stream) {
    const int32_t rows = (int32_t)std::min<int64_t>(m, 512);
    k_col_mean_sample<<<(unsigned)ceil_div(n, 32), 1024, 0, stream>>>(a, lda, rows, n, out);
    DCNR_LAUNCHED();
    return DCNR_OK;
}
}  // namespace dcnr

