// Fused dense Adam / AdamW step (the optimizer at train.py:201-204,226).
// Dense semantics like torch.optim.Adam(W): every row of every table is updated every step
// (moments decay and weight decay act on untouched rows too), so this is one HBM-bound pass
// over {param, grad, exp_avg, exp_avg_sq}.
#include "kernels.cuh"

namespace dcnr {

__global__ void __launch_bounds__(256)
k_adam(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, int64_t n,
       float lr, float beta1, float beta2, float eps, float weight_decay, int decoupled, float bc1, float bc2_sqrt) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float pv = p[i], gv = g[i];
        if (decoupled) pv *= 1.f - lr * weight_decay;          // AdamW: p <- p (1 - lr*wd)
        else gv = fmaf(weight_decay, pv, gv);                  // Adam : g <- g + wd*p
        const float mv = fmaf(beta1, m[i], (1.f - beta1) * gv);
        const float vv = fmaf(beta2, v[i], (1.f - beta2) * gv * gv);
        m[i] = mv;
        v[i] = vv;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;        // torch: (sqrt(v)/sqrt(bias_correction2)) + eps
        p[i] = pv - (lr / bc1) * (mv / denom);
    }
}

}  // namespace dcnr

using namespace dcnr;

extern "C" int dcnr_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr,
                              float beta1, float beta2, float eps, float weight_decay, int decoupled_weight_decay,
                              int64_t step, dcnr_stream_t stream) {
    DCNR_REQUIRE(param && grad && exp_avg && exp_avg_sq && step >= 1, "null argument / step < 1");
    if (n <= 0) return DCNR_OK;
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n, 256), (int64_t)sm_count() * 16);
    k_adam<<<grid, 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                               decoupled_weight_decay, bc1, bc2_sqrt);
    DCNR_LAUNCHED();
    return DCNR_OK;
}
