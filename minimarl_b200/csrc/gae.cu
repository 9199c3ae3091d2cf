// GAE / discounted-return reverse scan over T for sm_100a.
//
// Replaces SharedReplayBuffer.compute_returns (mappo/runner/shared/shared_buffer.py:131-157) with
// ValueNorm.denormalize (mappo/utils/valuenorm.py:67-78) folded in, and the advantage line of
// R_MAPPO.train (mappo/algorithms/ramppo_network.py:225-230).
//
// One thread owns one (env, agent) row and walks t = T-1 .. 0; consecutive threads own consecutive
// rows, so every load/store of a warp is one contiguous 128-byte line.  HBM-bound: 16 B per
// (t, row) (+4 B with advantages).  Loads of U steps are issued together before the dependent
// float64 chain consumes them, so each warp keeps 3*U lines in flight.
//
// Arithmetic mirrors the reference's mixed precision operation by operation (numpy float32 arrays,
// Python-float scalars, a float64 `gae` accumulator), with contraction into FMA forbidden, so the
// result is bit-identical to the reference and not merely within tolerance:
//   Vhat(x) = fl32(fl32(x * scale) + shift)                          torch float32
//   delta   = fl32(fl32(r + fl32(fl32(g32 * Vhat[t+1]) * m[t+1])) - Vhat[t])
//   gae     = fl64(delta + fl64(fl32(gl32 * m[t+1]) * gae))          gae is float64
//   ret[t]  = fl32(fl64(gae + Vhat[t]))
#include "common.cuh"

namespace {

constexpr int kUnroll = 16;

struct GaeParams {
    const float* __restrict__ rewards;
    const float* __restrict__ value_preds;
    const float* __restrict__ masks;
    float* __restrict__ returns;
    float* __restrict__ advantages;
    int T;
    long long R;
    float g32;       // float32(gamma)
    float gl32;      // float32(gamma * gae_lambda), the product taken in double first
    float scale, shift;
};

template <bool DENORM>
__device__ __forceinline__ float vhat(float x, float scale, float shift) {
    return DENORM ? __fadd_rn(__fmul_rn(x, scale), shift) : x;
}

template <bool USE_GAE, bool DENORM, bool ADV>
__global__ void __launch_bounds__(128) gae_kernel(const GaeParams p) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= p.R) return;
    const long long R = p.R;
    const float* rew = p.rewards + r;
    const float* vp = p.value_preds + r;
    const float* mk = p.masks + r;
    float* ret = p.returns + r;
    float* adv = ADV ? p.advantages + r : nullptr;

    float m_next = mm::ld_stream(mk + (long long)p.T * R);
    float vh_next = 0.f, ret_next = 0.f;
    if (USE_GAE) vh_next = vhat<DENORM>(mm::ld_stream(vp + (long long)p.T * R), p.scale, p.shift);
    else         ret_next = ret[(long long)p.T * R];
    double gae = 0.0;

    for (int t_hi = p.T; t_hi > 0; t_hi -= kUnroll) {
        float rw[kUnroll], vv[kUnroll], mm_[kUnroll];
#pragma unroll
        for (int i = 0; i < kUnroll; ++i) {
            const int t = t_hi - 1 - i;
            if (t >= 0) {
                rw[i] = mm::ld_stream(rew + (long long)t * R);
                if (USE_GAE) vv[i] = mm::ld_stream(vp + (long long)t * R);
                mm_[i] = (t > 0) ? mm::ld_stream(mk + (long long)t * R) : 0.f;   // masks[0] is never used
            }
        }
#pragma unroll
        for (int i = 0; i < kUnroll; ++i) {
            const int t = t_hi - 1 - i;
            if (t >= 0) {
                float out;
                if (USE_GAE) {
                    const float vh = vhat<DENORM>(vv[i], p.scale, p.shift);
                    const float boot = __fmul_rn(__fmul_rn(p.g32, vh_next), m_next);
                    const float delta = __fsub_rn(__fadd_rn(rw[i], boot), vh);
                    const float c = __fmul_rn(p.gl32, m_next);
                    gae = __dadd_rn((double)delta, __dmul_rn((double)c, gae));
                    out = (float)__dadd_rn(gae, (double)vh);
                    if (ADV) adv[(long long)t * R] = __fsub_rn(out, vh);
                    vh_next = vh;
                } else {
                    out = __fadd_rn(rw[i], __fmul_rn(__fmul_rn(ret_next, p.g32), m_next));
                    ret_next = out;
                }
                ret[(long long)t * R] = out;
                m_next = mm_[i];
            }
        }
    }
}

}  // namespace

extern "C" int mm_gae_returns(const float* rewards, const float* value_preds, const float* masks,
                              float* returns, float* advantages, int32_t T, int64_t R, double gamma,
                              double gae_lambda, int32_t use_gae, int32_t use_denorm, float denorm_scale,
                              float denorm_shift, void* stream) {
    if (!rewards || !masks || !returns || T <= 0 || R <= 0) return MM_ERR_BAD_ARG;
    if (use_gae && !value_preds) return MM_ERR_BAD_ARG;
    if (advantages && !use_gae) return MM_ERR_BAD_ARG;
    GaeParams p;
    p.rewards = rewards; p.value_preds = value_preds; p.masks = masks;
    p.returns = returns; p.advantages = advantages;
    p.T = T; p.R = R;
    p.g32 = (float)gamma;
    p.gl32 = (float)(gamma * gae_lambda);
    p.scale = denorm_scale; p.shift = denorm_shift;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // few rows: 1-warp blocks spread them over all 148 SMs; many rows: 128-thread blocks
    const int threads = (R <= 148LL * 32 * 16) ? 32 : 128;
    const dim3 grid((unsigned)((R + threads - 1) / threads));
    const bool adv = advantages != nullptr;
    if (use_gae) {
        if (use_denorm) { if (adv) gae_kernel<true, true, true><<<grid, threads, 0, st>>>(p);
                          else     gae_kernel<true, true, false><<<grid, threads, 0, st>>>(p); }
        else            { if (adv) gae_kernel<true, false, true><<<grid, threads, 0, st>>>(p);
                          else     gae_kernel<true, false, false><<<grid, threads, 0, st>>>(p); }
    } else {
        gae_kernel<false, false, false><<<grid, threads, 0, st>>>(p);
    }
    return mm::check_launch();
}
