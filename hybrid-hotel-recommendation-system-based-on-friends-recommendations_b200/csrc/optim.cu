// Fused dense Adam / AdamW step (the optimizer at train.py:201-204,226).
// Dense semantics like torch.optim.Adam(W): every row of every table is updated every step
// (moments decay and weight decay act on untouched rows too), so this is one HBM-bound pass
// over {param, grad, exp_avg, exp_avg_sq}.
#include "kernels.cuh"

namespace dcnr {

// Mirrors torch's foreach Adam arithmetic: the scalar factors (1 - beta, lr / bias_correction1, sqrt(bias_correction2))
// are formed in DOUBLE on the host like Python does and only then rounded to fp32:
//   m = lerp(m, g, 1 - beta1);  v = v * beta2 + (1 - beta2) * g * g;  p -= step_size * m / (sqrt(v) / bc2_sqrt + eps)
__global__ void __launch_bounds__(256)
k_adam(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v, int64_t n,
       float lr_wd, float w1, float beta2, float w2, float eps, float weight_decay, int decoupled, float step_size,
       float bc2_sqrt) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float pv = p[i], gv = g[i];
        if (decoupled) pv *= 1.f - lr_wd;                      // AdamW: p <- p (1 - lr*wd)
        else gv = fmaf(weight_decay, pv, gv);                  // Adam : g <- g + wd*p
        const float m0 = m[i];
        const float mv = fmaf(w1, gv - m0, m0);                // lerp
        const float vv = fmaf(w2 * gv, gv, v[i] * beta2);      // mul_ then addcmul_
        m[i] = mv;
        v[i] = vv;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        p[i] = pv - step_size * (mv / denom);
    }
}

}  // namespace dcnr

using namespace dcnr;

extern "C" int dcnr_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, double lr,
                              double beta1, double beta2, double eps, double weight_decay, int decoupled_weight_decay,
                              int64_t step, dcnr_stream_t stream) {
    DCNR_REQUIRE(param && grad && exp_avg && exp_avg_sq && step >= 1, "null argument / step < 1");
    if (n <= 0) return DCNR_OK;
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2_sqrt = sqrt(1.0 - pow(beta2, (double)step));
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n, 256), (int64_t)sm_count() * 16);
    k_adam<<<grid, 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, (float)(lr * weight_decay),
                                               (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps,
                                               (float)weight_decay, decoupled_weight_decay, (float)(lr / bc1),
                                               (float)bc2_sqrt);
    DCNR_LAUNCHED();
    return DCNR_OK;
}
