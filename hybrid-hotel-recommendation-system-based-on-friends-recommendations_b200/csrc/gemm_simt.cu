// IEEE-fp32 CUDA-core GEMM (the DCNR_PREC_FP32 parity path) for the three contractions of an
// nn.Linear:  forward  y = x W^T (+epilogue)            train.py:161,114,118
//             dgrad    dx = dy W                        autograd at train.py:225
//             wgrad    dW = dy^T x   (split over batch) autograd at train.py:225
// 128x128x8 tiles, 256 threads, 8x8 register micro-tiles split 4+4 so shared-memory reads are
// conflict-free 128-bit loads, double-buffered shared memory with register prefetch.
// The tensor-core paths (tcgen05, gemm_tc.cu) share the same launcher signature.
#include "kernels.cuh"

namespace dcnr {

constexpr int BM = 128, BN = 128, BK = 8, LDS_ = BM + 4;   // padded smem row: conflict-free transposed stores

template <bool KMAJOR>
struct TileLoader {
    // KMAJOR : element(row, k) = p[row*ld + k]  -> thread loads 4 consecutive k of one row
    // !KMAJOR: element(row, k) = p[k*ld + row]  -> thread loads 4 consecutive rows of one k
    __device__ static __forceinline__ float4 load(const float *__restrict__ p, int64_t ld, int64_t row0, int64_t rows,
                                                  int64_t k0, int64_t k_end, bool vec, int t) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (KMAJOR) {
            const int64_t r = row0 + (t >> 1), k = k0 + (t & 1) * 4;
            if (r < rows) {
                const float *q = p + r * ld + k;
                if (vec && k + 3 < k_end) {
                    v = ldg4(q);
                } else {
                    if (k + 0 < k_end) v.x = __ldg(q + 0);
                    if (k + 1 < k_end) v.y = __ldg(q + 1);
                    if (k + 2 < k_end) v.z = __ldg(q + 2);
                    if (k + 3 < k_end) v.w = __ldg(q + 3);
                }
            }
        } else {
            const int64_t k = k0 + (t >> 5), r = row0 + (t & 31) * 4;
            if (k < k_end) {
                const float *q = p + k * ld + r;
                if (vec && r + 3 < rows) {
                    v = ldg4(q);
                } else {
                    if (r + 0 < rows) v.x = __ldg(q + 0);
                    if (r + 1 < rows) v.y = __ldg(q + 1);
                    if (r + 2 < rows) v.z = __ldg(q + 2);
                    if (r + 3 < rows) v.w = __ldg(q + 3);
                }
            }
        }
        return v;
    }
    __device__ static __forceinline__ void store(float (*s)[LDS_], float4 v, int t) {
        if (KMAJOR) {
            const int r = t >> 1, k = (t & 1) * 4;
            s[k + 0][r] = v.x; s[k + 1][r] = v.y; s[k + 2][r] = v.z; s[k + 3][r] = v.w;
        } else {
            const int k = t >> 5, r = (t & 31) * 4;
            *reinterpret_cast<float4 *>(&s[k][r]) = v;
        }
    }
};

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(256, 2)
k_sgemm(const float *__restrict__ A, int64_t lda, const float *__restrict__ B, int64_t ldb, float *__restrict__ C,
        int64_t ldc, int64_t M, int64_t N, int64_t K, int64_t k_per_split, GemmEpilogue epi) {
    __shared__ __align__(16) float As[2][BK][LDS_];
    __shared__ __align__(16) float Bs[2][BK][LDS_];
    const int t = threadIdx.x;
    const int ty = t >> 4, tx = t & 15;
    const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
    const int64_t k_begin = (int64_t)blockIdx.z * k_per_split;
    const int64_t k_end = min(K, k_begin + k_per_split);
    const bool vecA = (lda & 3) == 0 && ((uintptr_t)A & 15) == 0;
    const bool vecB = (ldb & 3) == 0 && ((uintptr_t)B & 15) == 0;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int64_t nk = (k_end - k_begin + BK - 1) / BK;
    float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;
    if (nk > 0) {
        ra = TileLoader<A_KMAJOR>::load(A, lda, m0, M, k_begin, k_end, vecA, t);
        rb = TileLoader<B_KMAJOR>::load(B, ldb, n0, N, k_begin, k_end, vecB, t);
        TileLoader<A_KMAJOR>::store(As[0], ra, t);
        TileLoader<B_KMAJOR>::store(Bs[0], rb, t);
    }
    __syncthreads();
    for (int64_t kt = 0; kt < nk; ++kt) {
        const int cur = (int)(kt & 1);
        if (kt + 1 < nk) {
            const int64_t k0 = k_begin + (kt + 1) * BK;
            ra = TileLoader<A_KMAJOR>::load(A, lda, m0, M, k0, k_end, vecA, t);
            rb = TileLoader<B_KMAJOR>::load(B, ldb, n0, N, k0, k_end, vecB, t);
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[cur][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[cur][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[cur][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[cur][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            TileLoader<A_KMAJOR>::store(As[cur ^ 1], ra, t);
            TileLoader<B_KMAJOR>::store(Bs[cur ^ 1], rb, t);
        }
        __syncthreads();
    }

    // epilogue: v = (acc*col_scale + bias) [* hadamard] + residual ; relu
    float *Cz = C + (int64_t)blockIdx.z * M * ldc;
    const bool vecC = (ldc & 3) == 0 && ((uintptr_t)Cz & 15) == 0;
    const bool vecR = epi.residual != nullptr && (epi.ldr & 3) == 0 && ((uintptr_t)epi.residual & 15) == 0;
#pragma unroll
    for (int ih = 0; ih < 2; ++ih) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t m = m0 + ih * 64 + ty * 4 + i;
            if (m >= M) continue;
#pragma unroll
            for (int jh = 0; jh < 2; ++jh) {
                const int64_t n = n0 + jh * 64 + tx * 4;
                if (n >= N) continue;
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = acc[ih * 4 + i][jh * 4 + j];
                const bool full = n + 3 < N;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (n + j < N) {
                        if (epi.col_scale != nullptr) v[j] *= __ldg(epi.col_scale + n + j);
                        if (epi.bias != nullptr) v[j] += __ldg(epi.bias + n + j);
                    }
                }
                if (epi.hadamard != nullptr) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (n + j < N) v[j] *= __ldg(epi.hadamard + m * epi.ldh + n + j);
                }
                if (epi.residual != nullptr) {
                    if (full && vecR) {
                        const float4 r4 = ldg4(epi.residual + m * epi.ldr + n);
                        v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (n + j < N) v[j] += __ldg(epi.residual + m * epi.ldr + n + j);
                    }
                }
                if (epi.relu) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
                }
                if (full && vecC) {
                    st4(Cz + m * ldc + n, make_float4(v[0], v[1], v[2], v[3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (n + j < N) Cz[m * ldc + n + j] = v[j];
                }
            }
        }
    }
}

int launch_gemm_simt(const float *A, int64_t lda, bool a_kmajor, const float *B, int64_t ldb, bool b_kmajor, float *C,
                     int64_t ldc, int64_t m, int64_t n, int64_t k, int split_k, const GemmEpilogue &epi,
                     cudaStream_t stream) {
    if (m <= 0 || n <= 0) return DCNR_OK;
    DCNR_REQUIRE(split_k >= 1, "split_k must be >= 1");
    const int64_t k_per_split = round_up(ceil_div(std::max<int64_t>(k, 1), split_k), BK);
    dim3 grid((unsigned)ceil_div(m, BM), (unsigned)ceil_div(n, BN), (unsigned)split_k);
    DCNR_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "gemm grid too large");
    gemm_timer_before(stream, 2.0 * (double)m * (double)n * (double)k);
    if (a_kmajor && b_kmajor)
        k_sgemm<true, true><<<grid, 256, 0, stream>>>(A, lda, B, ldb, C, ldc, m, n, k, k_per_split, epi);
    else if (a_kmajor && !b_kmajor)
        k_sgemm<true, false><<<grid, 256, 0, stream>>>(A, lda, B, ldb, C, ldc, m, n, k, k_per_split, epi);
    else if (!a_kmajor && !b_kmajor)
        k_sgemm<false, false><<<grid, 256, 0, stream>>>(A, lda, B, ldb, C, ldc, m, n, k, k_per_split, epi);
    else
        k_sgemm<false, true><<<grid, 256, 0, stream>>>(A, lda, B, ldb, C, ldc, m, n, k, k_per_split, epi);
    gemm_timer_after(stream);
    DCNR_LAUNCHED();
    return DCNR_OK;
}

}  // namespace dcnr
