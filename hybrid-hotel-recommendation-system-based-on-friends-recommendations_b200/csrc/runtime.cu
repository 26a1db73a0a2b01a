// Error reporting, launch accounting and the generic deterministic partial-sum finalizers.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace dcnr {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};   // process-wide: autograd runs backward on its own thread

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return DCNR_ERR_CUDA;
}

void count_launch(int n) { g_launches += n; }

int sm_count() {
    static thread_local int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
        cached = n;
    }
    return cached;
}

// One thread per output column; partial rows are added in ascending order in double so the
// result does not depend on how the producing grid was scheduled.
__global__ void k_sum_partials(const float *__restrict__ partials, int64_t n_partials, int64_t ld, SegPtrs seg) {
    int s = blockIdx.y;
    if (s >= seg.n || seg.out[s] == nullptr) return;
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= seg.len[s]) return;
    const float *p = partials + seg.offset[s] + c;
    double acc = 0.0;
    int64_t i = 0;
    for (; i + 8 <= n_partials; i += 8) {      // 8 loads in flight, added in ascending order (same result as the plain loop)
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldg(p + (i + j) * ld);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += (double)v[j];
    }
    for (; i < n_partials; ++i) acc += (double)p[i * ld];
    seg.out[s][c] = (float)acc;
}

int launch_sum_partials(const float *partials, int64_t n_partials, int64_t ld, const SegPtrs &seg,
                        cudaStream_t stream) {
    int maxlen = 0;
    for (int i = 0; i < seg.n; ++i)
        if (seg.out[i] && seg.len[i] > maxlen) maxlen = seg.len[i];
    if (maxlen == 0) return DCNR_OK;
    dim3 grid((unsigned)ceil_div(maxlen, 128), (unsigned)seg.n);
    k_sum_partials<<<grid, 128, 0, stream>>>(partials, n_partials, ld, seg);
    DCNR_LAUNCHED();
    return DCNR_OK;
}

__global__ void k_sum_partials_2d(const float *__restrict__ partials, int64_t n_partials, int32_t rows,
                                  int32_t cols_pad, int32_t cols, float *__restrict__ out, int64_t ldo) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)rows * cols_pad) return;
    int r = (int)(e / cols_pad), c = (int)(e % cols_pad);
    if (c >= cols) return;
    double acc = 0.0;
    const int64_t stride = (int64_t)rows * cols_pad;
    int64_t p = 0;
    for (; p + 8 <= n_partials; p += 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldg(partials + (p + j) * stride + e);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += (double)v[j];
    }
    for (; p < n_partials; ++p) acc += (double)partials[p * stride + e];
    out[(int64_t)r * ldo + c] = (float)acc;
}

int launch_sum_partials_2d(const float *partials, int64_t n_partials, int32_t rows, int32_t cols_pad,
                           int32_t cols, float *out, int64_t ldo, cudaStream_t stream) {
    int64_t total = (int64_t)rows * cols_pad;
    k_sum_partials_2d<<<(unsigned)ceil_div(total, 256), 256, 0, stream>>>(partials, n_partials, rows, cols_pad,
                                                                        cols, out, ldo);
    DCNR_LAUNCHED();
    return DCNR_OK;
}

}  // namespace dcnr

extern "C" {
int dcnr_abi_version(void) { return DCNR_ABI_VERSION; }
const char *dcnr_last_error_string(void) { return dcnr::g_err; }
int64_t dcnr_launch_count(int reset) {
    return reset ? dcnr::g_launches.exchange(0) : dcnr::g_launches.load();
}
}
