// MMR re-rank on the device: replaces the O(20 * C) scikit-learn cosine_similarity calls of
// rerank_with_mmr (main.py:133-169) by one CTA per request.
//
// Contract (bit-exact with oracle/mmr_oracle.c): candidate vectors are normalised with sequential-fma norms, the
// similarity to the most recently selected item is a sequential-fma dot product folded into a running maximum
// (max is exact, so this equals np.max over all selected), mmr = lambda*score - (1-lambda)*max_sim in fp32 with two
// multiplications and one subtraction, and the FIRST maximum in ranked order wins (the `>` at main.py:162).
// Candidates without an embedding row (emb_idx < 0) are skipped like main.py:150; the best-scored candidate is always
// taken first (main.py:142-143).
#include "kernels.cuh"

namespace dcnr {

constexpr int kMT = 256;

__global__ void __launch_bounds__(kMT)
k_mmr_rerank(const float *__restrict__ emb, int d, const float *__restrict__ scores, const int64_t *__restrict__ emb_idx,
             const int32_t *__restrict__ offsets, float lambda, float one_minus, int top_k, int32_t *__restrict__ order_out,
             int32_t *__restrict__ count_out) {
    extern __shared__ __align__(16) float sm[];
    const int r = blockIdx.x, tid = threadIdx.x;
    const int c0 = offsets[r], C = offsets[r + 1] - c0;
    float *vhat = sm;                                   // [C][d]
    float *maxsim = vhat + (size_t)C * d;               // [C]
    int *state = reinterpret_cast<int *>(maxsim + C);   // [C]: 0 free, 1 taken, 2 unmapped
    __shared__ float red_v[kMT / 32];
    __shared__ int red_c[kMT / 32];
    __shared__ int s_best;
    int32_t *out = order_out + (int64_t)r * top_k;
    for (int i = tid; i < top_k; i += kMT) out[i] = -1;
    if (C <= 0) {
        if (tid == 0) count_out[r] = 0;
        return;
    }
    for (int c = tid; c < C; c += kMT) {
        const int64_t id = emb_idx[c0 + c];
        maxsim[c] = -INFINITY;
        state[c] = id < 0 ? 2 : 0;
        if (id < 0) continue;
        const float *v = emb + id * (int64_t)d;
        float ss = 0.f;
        for (int j = 0; j < d; ++j) ss = __fmaf_rn(v[j], v[j], ss);
        float nrm = __fsqrt_rn(ss);
        if (nrm == 0.f) nrm = 1.f;
        for (int j = 0; j < d; ++j) vhat[(size_t)c * d + j] = __fdiv_rn(v[j], nrm);
    }
    __syncthreads();
    const int want = min(top_k, C);
    int n = 1, last = 0;
    bool have_sel = false;
    if (tid == 0) {
        out[0] = 0;
        state[0] = state[0] == 2 ? 3 : 1;               // 3 = taken and unmapped
    }
    __syncthreads();
    while (n < want) {
        const bool last_mapped = state[last] == 1;
        have_sel = have_sel || last_mapped;
        float bv = -INFINITY;
        int bc = -1;
        for (int c = tid; c < C; c += kMT) {
            if (state[c] != 0) continue;
            if (last_mapped) {
                float s = 0.f;
                for (int j = 0; j < d; ++j) s = __fmaf_rn(vhat[(size_t)c * d + j], vhat[(size_t)last * d + j], s);
                if (s > maxsim[c]) maxsim[c] = s;
            }
            const float ms = have_sel ? maxsim[c] : 0.f;
            const float v = __fsub_rn(__fmul_rn(lambda, scores[c0 + c]), __fmul_rn(one_minus, ms));
            if (v > bv) { bv = v; bc = c; }             // ascending c: keeps the first maximum of this thread's candidates
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {              // (value desc, position asc)
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
            if (oc >= 0 && (bc < 0 || ov > bv || (ov == bv && oc < bc))) { bv = ov; bc = oc; }
        }
        if ((tid & 31) == 0) { red_v[tid >> 5] = bv; red_c[tid >> 5] = bc; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < kMT / 32; ++w) {
                const float ov = red_v[w];
                const int oc = red_c[w];
                if (oc >= 0 && (bc < 0 || ov > bv || (ov == bv && oc < bc))) { bv = ov; bc = oc; }
            }
            s_best = bc;
            if (bc >= 0) { out[n] = bc; state[bc] = 1; }
        }
        __syncthreads();
        if (s_best < 0) break;
        last = s_best;
        ++n;
    }
    if (tid == 0) count_out[r] = n;
}

}  // namespace dcnr

using namespace dcnr;

extern "C" int dcnr_mmr_rerank(const float *item_emb, int64_t n_items, int32_t d, const float *scores, const int64_t *emb_idx,
                               const int32_t *offsets, int32_t n_requests, float lambda, int32_t top_k, int32_t max_candidates,
                               int32_t *order_out, int32_t *count_out, dcnr_stream_t stream) {
    DCNR_REQUIRE(item_emb && scores && emb_idx && offsets && order_out && count_out, "null argument");
    DCNR_REQUIRE(n_items >= 1 && d >= 1 && top_k >= 1 && max_candidates >= 0, "bad argument");
    if (n_requests <= 0) return DCNR_OK;
    const size_t smem = (size_t)max_candidates * (d + 2) * sizeof(float) + 16;
    DCNR_REQUIRE(smem <= 220 * 1024, "a request has too many candidates for one CTA (%d x %d floats)", max_candidates, d);
    if (smem > 48 * 1024)
        DCNR_CUDA_CHECK(cudaFuncSetAttribute(k_mmr_rerank, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const float one_minus = (float)(1.0 - (double)lambda);
    k_mmr_rerank<<<(unsigned)n_requests, kMT, smem, as_stream(stream)>>>(item_emb, d, scores, emb_idx, offsets, lambda, one_minus,
                                                                        top_k, order_out, count_out);
    DCNR_LAUNCHED();
    return DCNR_OK;
}
