// Gradient-norm clipping + Adam on ONE flat parameter buffer: what the learners run after every backward pass
// (torch.nn.utils.clip_grad_norm_ followed by torch.optim.Adam.step: offpolicy/algorithms/qmix/qmix.py:200-208,
// mappo/algorithms/ramppo_network.py:181-209, vdn/_train.py:96-100).  The networks on this path have 10^4..10^5 parameters in
// ~40 tensors; PyTorch's foreach implementation is ~25 launches (0.4 ms) per step, this is two launches of a few microseconds:
//   1. per-CTA partial sums of squares of the gradient (fixed order -> deterministic), the step counter advanced on the device;
//   2. every CTA re-adds the partials in the same order, forms the clip coefficient and applies the Adam update.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxCtas = 148;

__global__ void __launch_bounds__(kThreads) gradnorm_partial_kernel(const float* __restrict__ g, long long n,
                                                                    double* __restrict__ partial, float* __restrict__ step) {
    __shared__ double sm[kThreads / 32];
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
        const float v = g[i];
        acc += (double)v * (double)v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(MM_FULL_MASK, acc, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) s += sm[w];
        partial[blockIdx.x] = s;
        if (blockIdx.x == 0 && step) *step += 1.0f;
    }
}

struct AdamArgs {
    float* p;
    const float* g;
    float* m;
    float* v;
    const float* step;
    const double* partial;
    const float* lr_dev;
    float* norm_out;
    long long n, clip_n;
    int n_partial;
    float lr, beta1, beta2, eps, weight_decay, max_norm;
};

__global__ void __launch_bounds__(kThreads) clip_adam_kernel(AdamArgs a) {
    __shared__ float s_coef;
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < a.n_partial; ++i) s += a.partial[i];
        const float norm = (float)sqrt(s);
        float coef = 1.f;
        if (a.max_norm > 0.f) coef = fminf(a.max_norm / (norm + 1e-6f), 1.f);     // clip_grad_norm_'s clamp
        if (!(norm == norm)) coef = norm;                                          // NaN propagates as in PyTorch
        s_coef = coef;
        if (blockIdx.x == 0 && a.norm_out) *a.norm_out = norm;
    }
    __syncthreads();
    const float coef = s_coef;
    const float t = *a.step;
    const float bc1 = 1.f - powf(a.beta1, t);
    const float bc2_sqrt = sqrtf(1.f - powf(a.beta2, t));
    const float step_size = (a.lr_dev ? *a.lr_dev : a.lr) / bc1;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < a.n; i += (long long)gridDim.x * kThreads) {
        float p = a.p[i];
        float g = a.g[i];
        if (i < a.clip_n) g *= coef;
        if (a.weight_decay != 0.f) g = fmaf(a.weight_decay, p, g);
        float m = a.m[i], v = a.v[i];
        m = m + (g - m) * (1.f - a.beta1);                                        // exp_avg.lerp_(grad, 1 - beta1)
        v = v * a.beta2 + (1.f - a.beta2) * g * g;
        const float denom = sqrtf(v) / bc2_sqrt + a.eps;
        p -= step_size * (m / denom);
        a.p[i] = p;
        a.m[i] = m;
        a.v[i] = v;
    }
}

}  // namespace

extern "C" int mm_clip_adam_scratch_doubles(void) { return kMaxCtas; }

extern "C" int mm_clip_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* step,
                                 double* scratch, float* norm_out, int64_t n, int64_t clip_n, float lr, const float* lr_dev,
                                 float beta1, float beta2, float eps, float weight_decay, float max_norm, void* stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || !step || !scratch || n <= 0 || clip_n < 0 || clip_n > n)
        return MM_ERR_BAD_ARG;
    auto st = static_cast<cudaStream_t>(stream);
    long long ctas = (n + kThreads * 4 - 1) / (kThreads * 4);
    if (ctas > kMaxCtas) ctas = kMaxCtas;
    gradnorm_partial_kernel<<<(unsigned)ctas, kThreads, 0, st>>>(grads, clip_n, scratch, step);
    AdamArgs a{params, grads, exp_avg, exp_avg_sq, step, scratch, lr_dev, norm_out, n, clip_n, (int)ctas,
               lr, beta1, beta2, eps, weight_decay, max_norm};
    clip_adam_kernel<<<(unsigned)ctas, kThreads, 0, st>>>(a);
    return mm::check_launch();
}
