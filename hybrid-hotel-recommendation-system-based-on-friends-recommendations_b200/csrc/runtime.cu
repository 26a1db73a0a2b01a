// Error reporting, launch accounting and the generic deterministic partial-sum finalizers.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <vector>

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

// Optional device-side timing of the dense-layer GEMM launches (dcnr_gemm_timing_begin / _end): a CUDA event pair on the
// launching stream around every GEMM launch, summed at _end.  bench.py uses it to report the dominant kernel's achieved
// rate over the very steps it times.  Off by default: no events are created or recorded.
namespace {
struct GemmTimer {
    std::mutex mu;
    bool on = false;
    std::vector<cudaEvent_t> pool;       // 2 events per launch, reused across sessions
    size_t used = 0;
    std::vector<double> flops;
} g_timer;
}  // namespace

void gemm_timer_before(cudaStream_t stream, double flops) {
    if (!g_timer.on) return;
    std::lock_guard<std::mutex> lk(g_timer.mu);
    if (!g_timer.on) return;
    while (g_timer.pool.size() < g_timer.used + 2) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        g_timer.pool.push_back(e);
    }
    cudaEventRecord(g_timer.pool[g_timer.used], stream);
    g_timer.flops.push_back(flops);
    g_timer.used += 2;
}
void gemm_timer_after(cudaStream_t stream) {
    if (!g_timer.on) return;
    std::lock_guard<std::mutex> lk(g_timer.mu);
    if (!g_timer.on || g_timer.used < 2) return;
    cudaEventRecord(g_timer.pool[g_timer.used - 1], stream);
}

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

// 32 columns x 8 slices per CTA: slice y adds its contiguous range of partial rows in ascending order in double,
// then the 8 slice sums are added in slice order -- a fixed association, so the result does not depend on how the
// producing grid was scheduled.  (One thread per column over ALL partials was a 20 us latency chain per call.)
// Slices per CTA = blockDim.x / 32: 8 for short partial lists, 32 from 64 partials on (a B = 65 536 step reduces 256 chunk
// partials a dozen times: 32 slices make that one batch of 8 loads per thread instead of four dependent ones).
constexpr int kSumSlices = 8, kSumSlicesMax = 32;
static inline int sum_slices(int64_t n_partials) { return n_partials >= 64 ? kSumSlicesMax : kSumSlices; }
__global__ void __launch_bounds__(32 * kSumSlicesMax)
k_sum_partials(const float *__restrict__ partials, int64_t n_partials, int64_t ld, SegPtrs seg) {
    __shared__ double sh[kSumSlicesMax][33];
    const int kSumSlices = blockDim.x >> 5;
    const int s = blockIdx.y;
    if (s >= seg.n || seg.out[s] == nullptr) return;                 // CTA-uniform
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    const int64_t per = (n_partials + kSumSlices - 1) / kSumSlices;
    const int64_t i0 = min(n_partials, ty * per), i1 = min(n_partials, i0 + per);
    double acc = 0.0;
    if (c < seg.len[s]) {
        const float *p = partials + seg.offset[s] + c;
        int64_t i = i0;
        for (; i + 8 <= i1; i += 8) {      // 8 loads in flight, added in ascending order
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldg(p + (i + j) * ld);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += (double)v[j];
        }
        for (; i < i1; ++i) acc += (double)__ldg(p + i * ld);
    }
    sh[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && c < seg.len[s]) {
        double t = sh[0][tx];
#pragma unroll
        for (int y = 1; y < kSumSlices; ++y) t += sh[y][tx];
        seg.out[s][c] = (float)t;
    }
}

int launch_sum_partials(const float *partials, int64_t n_partials, int64_t ld, const SegPtrs &seg,
                        cudaStream_t stream) {
    int maxlen = 0;
    for (int i = 0; i < seg.n; ++i)
        if (seg.out[i] && seg.len[i] > maxlen) maxlen = seg.len[i];
    if (maxlen == 0) return DCNR_OK;
    dim3 grid((unsigned)ceil_div(maxlen, 32), (unsigned)seg.n);
    k_sum_partials<<<grid, 32 * sum_slices(n_partials), 0, stream>>>(partials, n_partials, ld, seg);
    DCNR_LAUNCHED();
    return DCNR_OK;
}

// out[r, c] = sum over partial tables p of partials[p][r][c], same 32 x 8 scheme as k_sum_partials (fixed slices,
// ascending order inside a slice, slices added in order).  A single chain over ~1 500 partial tables (tiny-table scatter
// at B = 1 M) took 130 us.
__global__ void __launch_bounds__(32 * kSumSlicesMax)
k_sum_partials_2d(const float *__restrict__ partials, int64_t n_partials, int32_t rows, int32_t cols_pad, int32_t cols,
                  float *__restrict__ out, int64_t ldo, int64_t pitch,         // pitch: floats between consecutive partial tables
                  const float *__restrict__ u, const float *__restrict__ v) {  // optional rank-1 term u[r] * v[c]
    __shared__ double sh[kSumSlicesMax][33];
    const int kSumSlices = blockDim.x >> 5;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t e = (int64_t)blockIdx.x * 32 + tx;
    const int64_t total = (int64_t)rows * cols_pad;
    const int64_t per = (n_partials + kSumSlices - 1) / kSumSlices;
    const int64_t p0 = min(n_partials, ty * per), p1 = min(n_partials, p0 + per);
    double acc = 0.0;
    if (e < total) {
        int64_t p = p0;
        for (; p + 8 <= p1; p += 8) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldg(partials + (p + j) * pitch + e);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += (double)v[j];
        }
        for (; p < p1; ++p) acc += (double)__ldg(partials + p * pitch + e);
    }
    sh[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && e < total) {
        const int r = (int)(e / cols_pad), c = (int)(e % cols_pad);
        if (c < cols) {
            double t = sh[0][tx];
#pragma unroll
            for (int y = 1; y < kSumSlices; ++y) t += sh[y][tx];
            if (u != nullptr) t += (double)__ldg(u + r) * (double)__ldg(v + c);
            out[(int64_t)r * ldo + c] = (float)t;
        }
    }
}

// float4 form of k_sum_partials_2d for cols_pad % 4 == 0 (the weight-gradient slabs: ~150 tables of 256 KB): a CTA folds 128
// consecutive elements, so the grid is four times smaller and every load is 16 bytes -- same slices, same order of additions.
__global__ void __launch_bounds__(32 * kSumSlicesMax)
k_sum_partials_2d_v4(const float *__restrict__ partials, int64_t n_partials, int32_t rows, int32_t cols_pad, int32_t cols,
                     float *__restrict__ out, int64_t ldo, int64_t pitch, const float *__restrict__ u,
                     const float *__restrict__ v) {
    __shared__ double sh[kSumSlicesMax][32][5];
    const int kSumSlices = blockDim.x >> 5;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t e = ((int64_t)blockIdx.x * 32 + tx) * 4;
    const int64_t total = (int64_t)rows * cols_pad;
    const int64_t per = (n_partials + kSumSlices - 1) / kSumSlices;
    const int64_t p0 = min(n_partials, ty * per), p1 = min(n_partials, p0 + per);
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    if (e < total) {
        int64_t p = p0;
        for (; p + 4 <= p1; p += 4) {
            float4 w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = ldg4(partials + (p + j) * pitch + e);
#pragma unroll
            for (int j = 0; j < 4; ++j) { a0 += (double)w[j].x; a1 += (double)w[j].y; a2 += (double)w[j].z; a3 += (double)w[j].w; }
        }
        for (; p < p1; ++p) {
            const float4 w = ldg4(partials + p * pitch + e);
            a0 += (double)w.x; a1 += (double)w.y; a2 += (double)w.z; a3 += (double)w.w;
        }
    }
    sh[ty][tx][0] = a0; sh[ty][tx][1] = a1; sh[ty][tx][2] = a2; sh[ty][tx][3] = a3;
    __syncthreads();
    if (ty < 4 && e < total) {                         // warp `ty` finishes component `ty` of the 32 float4
        const int64_t ee = e + ty;
        const int r = (int)(ee / cols_pad), c = (int)(ee % cols_pad);
        if (c < cols) {
            double t = sh[0][tx][ty];
            for (int y = 1; y < kSumSlices; ++y) t += sh[y][tx][ty];
            if (u != nullptr) t += (double)__ldg(u + r) * (double)__ldg(v + c);
            out[(int64_t)r * ldo + c] = (float)t;
        }
    }
}

int launch_sum_partials_2d(const float *partials, int64_t n_partials, int32_t rows, int32_t cols_pad,
                           int32_t cols, float *out, int64_t ldo, cudaStream_t stream, int64_t pitch, const float *u,
                           const float *v) {
    int64_t total = (int64_t)rows * cols_pad;
    if (pitch <= 0) pitch = total;
    if (cols_pad % 4 == 0 && pitch % 4 == 0 && ((uintptr_t)partials & 15) == 0 && n_partials >= 8) {
        k_sum_partials_2d_v4<<<(unsigned)ceil_div(total, 128), 32 * sum_slices(n_partials), 0, stream>>>(
            partials, n_partials, rows, cols_pad, cols, out, ldo, pitch, v != nullptr ? u : nullptr, v);
        DCNR_LAUNCHED();
        return DCNR_OK;
    }
    k_sum_partials_2d<<<(unsigned)ceil_div(total, 32), 32 * sum_slices(n_partials), 0, stream>>>(partials, n_partials, rows, cols_pad,
                                                                                   cols, out, ldo, pitch, v != nullptr ? u : nullptr, v);
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
int dcnr_gemm_timing_begin(void) {
    std::lock_guard<std::mutex> lk(dcnr::g_timer.mu);
    dcnr::g_timer.used = 0;
    dcnr::g_timer.flops.clear();
    dcnr::g_timer.on = true;
    return DCNR_OK;
}
int dcnr_gemm_timing_end(double *total_ms, int64_t *launches, double *total_flops) {
    std::lock_guard<std::mutex> lk(dcnr::g_timer.mu);
    dcnr::g_timer.on = false;
    double ms = 0.0, fl = 0.0;
    int64_t n = 0;
    for (size_t i = 0; i + 1 < dcnr::g_timer.used; i += 2) {
        DCNR_CUDA_CHECK(cudaEventSynchronize(dcnr::g_timer.pool[i + 1]));
        float t = 0.f;
        DCNR_CUDA_CHECK(cudaEventElapsedTime(&t, dcnr::g_timer.pool[i], dcnr::g_timer.pool[i + 1]));
        ms += t;
        fl += dcnr::g_timer.flops[i / 2];
        ++n;
    }
    dcnr::g_timer.used = 0;
    dcnr::g_timer.flops.clear();
    if (total_ms) *total_ms = ms;
    if (launches) *launches = n;
    if (total_flops) *total_flops = fl;
    return DCNR_OK;
}
}
