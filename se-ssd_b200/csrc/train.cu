// train.cu -- optimiser-side kernels of the SE-SSD training step over FLAT parameter arenas (one launch per model, not per tensor).
//
// The reference walks the parameter list in Python for the teacher's exponential moving average
// (det3d/torchie/trainer/trainer_sessd.py:315-318: ema = alpha * ema + (1 - alpha) * param, ~300 tiny launches), flattens / unflattens
// the gradients around the all-reduce (det3d/core/utils/dist_utils.py:8-29) and steps a fastai-style Adam with decoupled weight decay
// (det3d/solver/fastai_optim.py).  Here every parameter of a model is a view into one contiguous fp32 buffer (sessd_b200/train.py:
// ParamArena), likewise its gradient, so each of these is ONE grid-stride kernel at the HBM roofline (8-28 bytes per element), and the
// all-reduce runs in place on the gradient arena (no flatten / unflatten copies).
#include "common.cuh"

namespace sessd {

// y = a * y + b * x    (EMA: a = alpha, b = 1 - alpha; gradient averaging: a = 1 / world, b = 0 with x = y)
__global__ void __launch_bounds__(256) axpby_kernel(float4 *__restrict__ y, const float4 *__restrict__ x, float a, float b, long long n4,
                                                    float *__restrict__ y_tail, const float *__restrict__ x_tail, int ntail) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = y[i];
        const float4 u = x ? __ldg(x + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        v.x = fmaf(a, v.x, b * u.x); v.y = fmaf(a, v.y, b * u.y); v.z = fmaf(a, v.z, b * u.z); v.w = fmaf(a, v.w, b * u.w);
        y[i] = v;
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < ntail) y_tail[threadIdx.x] = fmaf(a, y_tail[threadIdx.x], b * (x_tail ? x_tail[threadIdx.x] : 0.f));
}

// Adam with decoupled weight decay (torch.optim.AdamW semantics == fastai true_wd): p -= lr * wd * p; m, v updated; p -= step_size * m / (sqrt(v) / sqrt(bc2) + eps)
__global__ void __launch_bounds__(256) adamw_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                                                    long long n, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt) {
    const float step_size = lr / bc1;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = __ldg(g + i);
        float pi = p[i];
        pi *= 1.f - lr * wd;
        const float mi = m[i] + (1.f - beta1) * (gi - m[i]);          // torch: exp_avg.lerp_(grad, 1 - beta1)
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - step_size * (mi / denom);
    }
}

}  // namespace sessd

using namespace sessd;

// d_y[i] = a * d_y[i] + b * d_x[i]  (d_x nullable => b term dropped).  Teacher EMA of trainer_sessd.py:315-318 with a = alpha, b = 1 - alpha;
// gradient averaging after a SUM all-reduce with d_x = NULL, a = 1 / world_size.  Pointers 16-byte aligned.
extern "C" int sessd_axpby(float *d_y, const float *d_x, float a, float b, long long n, void *stream) {
    if (!d_y || n < 0 || ((uintptr_t)d_y & 15) || ((uintptr_t)d_x & 15)) return SESSD_EINVAL;
    if (n == 0) return 0;
    const long long n4 = n / 4;
    SESSD_LAUNCH(axpby_kernel, persistent_grid(n4 > 0 ? n4 : 1, 256), 256, 0, stream, reinterpret_cast<float4 *>(d_y), reinterpret_cast<const float4 *>(d_x),
                 a, b, n4, d_y + n4 * 4, d_x ? d_x + n4 * 4 : nullptr, (int)(n - n4 * 4));
    return last_error();
}

// One AdamW step over flat fp32 arenas (parameters, gradients, first / second moments); step = 1, 2, ... (bias corrections computed on the host).
extern "C" int sessd_adamw_step(float *d_param, const float *d_grad, float *d_exp_avg, float *d_exp_avg_sq, long long n, float lr, float beta1,
                                float beta2, float eps, float weight_decay, int step, void *stream) {
    if (!d_param || !d_grad || !d_exp_avg || !d_exp_avg_sq || n < 0 || step < 1) return SESSD_EINVAL;
    if (n == 0) return 0;
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    SESSD_LAUNCH(adamw_kernel, persistent_grid(n, 256), 256, 0, stream, d_param, d_grad, d_exp_avg, d_exp_avg_sq, n, lr, beta1, beta2, eps,
                 weight_decay, (float)bc1, (float)sqrt(bc2));
    return last_error();
}
