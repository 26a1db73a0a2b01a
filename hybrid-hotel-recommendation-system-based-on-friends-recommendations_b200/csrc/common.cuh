// Shared host/device helpers for libdcnr_sm100a.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dcnr.h"

namespace dcnr {

// ---- host side -----------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);
void count_launch(int n = 1);
int sm_count();

#define DCNR_CUDA_CHECK(expr)                                                             \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) return ::dcnr::cuda_fail(_e, #expr, __FILE__, __LINE__);   \
    } while (0)

// after every kernel launch: counts it and turns a launch failure into a status code
#define DCNR_LAUNCHED()                                                                   \
    do {                                                                                  \
        ::dcnr::count_launch();                                                           \
        DCNR_CUDA_CHECK(cudaGetLastError());                                              \
    } while (0)

#define DCNR_REQUIRE(cond, ...)                                                           \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            ::dcnr::set_error(__VA_ARGS__);                                               \
            return DCNR_ERR_INVALID;                                                      \
        }                                                                                 \
    } while (0)

#define DCNR_TRY(expr)                                                                    \
    do {                                                                                  \
        int _s = (expr);                                                                  \
        if (_s != DCNR_OK) return _s;                                                     \
    } while (0)

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }
static inline cudaStream_t as_stream(dcnr_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Bump allocator over a caller-provided workspace (256-byte aligned slices).
struct Arena {
    char *base;
    int64_t size, used;
    Arena(void *p, int64_t n) : base(reinterpret_cast<char *>(p)), size(n), used(0) {}
    template <typename T>
    T *take(int64_t count) {
        int64_t bytes = round_up(count * (int64_t)sizeof(T), 256);
        char *p = base ? base + used : nullptr;
        used += bytes;
        return reinterpret_cast<T *>(p);
    }
    bool ok() const { return base == nullptr || used <= size; }
};

// rows per fixed reduction chunk for every deterministic column reduction
constexpr int kChunkRows = 256;

// ---- device side ---------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// sum over an aligned group of 8 lanes (xor pattern: every lane gets the same value)
__device__ __forceinline__ float group8_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
#endif

// ---- internal launchers shared between translation units ------------------------------------
constexpr int kMaxSeg = 20;
struct SegPtrs {   // output vectors of one partial-sum finalize (NULL entries are skipped)
    float *out[kMaxSeg];
    int32_t offset[kMaxSeg];   // column offset inside a partial row
    int32_t len[kMaxSeg];
    int32_t n;
};
// out[i][c] = sum_p partials[p*ld + offset[i] + c], p ascending, accumulated in double.
// event pair around a dense-layer GEMM launch when dcnr_gemm_timing_begin() is active (no-ops otherwise)
void gemm_timer_before(cudaStream_t stream, double flops);
void gemm_timer_after(cudaStream_t stream);
int launch_sum_partials(const float *partials, int64_t n_partials, int64_t ld, const SegPtrs &seg,
                        cudaStream_t stream);
// out[r*ldo + c] = sum_p partials[(p*rows + r)*cols_pad + c] (+ u[r] * v[c] when both are given) for c < cols
int launch_sum_partials_2d(const float *partials, int64_t n_partials, int32_t rows, int32_t cols_pad,
                           int32_t cols, float *out, int64_t ldo, cudaStream_t stream, int64_t pitch = 0,   // pitch 0: dense
                           const float *u = nullptr, const float *v = nullptr);

}  // namespace dcnr
