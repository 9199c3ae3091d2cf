// Action-value selection (chosen-action gather, double-Q arg-max / max-Q) and the TD target / Bellman
// error / PER priorities of the recurrent QMIX/VDN update, for sm_100a.
//
// Replaces the elementwise tail of QMix.train_policy_on_batch
// (offpolicy/algorithms/qmix/qmix.py:134-191): q_values_from_actions (QMixPolicy.py:92-111),
// actions_from_q's greedy arg-max (QMixPolicy.py:123-175), the bootstrapped target, the bad-transition
// mask, the PER-weighted loss and the R2D2 priority mix.  All HBM-bound elementwise / small-reduction work.
#include "common.cuh"

namespace {

// rows r = a*B + b (agents stacked on the batch dim, qmix.py:112-113); outputs are [T, B, A]
__global__ void q_select_kernel(const float* __restrict__ q_all, const float* __restrict__ tq_all,
                                const float* __restrict__ acts, float* __restrict__ q_taken,
                                float* __restrict__ next_q, int64_t* __restrict__ act_idx,
                                int64_t* __restrict__ greedy, int T, int B, int A, int NA, int double_q) {
    const long long R = (long long)A * B;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over T1 * R
    if (i >= (long long)(T + 1) * R) return;
    const int t = (int)(i / R);
    const long long r = i - (long long)t * R;
    const int a = (int)(r / B), b = (int)(r - (long long)a * B);
    const float* q = q_all + i * NA;
    // greedy action of the live network: first maximum, like torch.max
    int g = 0;
    float best = q[0];
    for (int k = 1; k < NA; ++k)
        if (q[k] > best) { best = q[k]; g = k; }
    if (greedy) greedy[i] = g;
    if (t < T) {
        const float* ac = acts + i * NA;          // acts is [T, R, NA]: same flat index for t < T
        int ai = 0;
        float am = ac[0];
        for (int k = 1; k < NA; ++k)
            if (ac[k] > am) { am = ac[k]; ai = k; }
        const long long o = ((long long)t * B + b) * A + a;
        if (act_idx) act_idx[i] = ai;
        if (q_taken) q_taken[o] = q[ai];
    }
    if (t >= 1 && tq_all != nullptr && next_q != nullptr) {
        const float* tq = tq_all + i * NA;
        float v;
        if (double_q) {
            v = tq[g];                             // target net evaluated at the live net's greedy action
        } else {
            v = tq[0];
            for (int k = 1; k < NA; ++k) v = fmaxf(v, tq[k]);
        }
        next_q[((long long)(t - 1) * B + b) * A + a] = v;   // "don't need the first Q values for next step"
    }
}

// one thread per episode b, walking t = 0..T-1 (consecutive threads -> consecutive b: coalesced)
__global__ void qmix_td_kernel(const float* __restrict__ q_tot, const float* __restrict__ next_q_tot,
                               const float* __restrict__ rewards, const float* __restrict__ dones_env,
                               const float* __restrict__ weights, float gamma, float per_nu, float per_eps,
                               int T, int B, float* __restrict__ err_out, float* __restrict__ ep_loss,
                               float* __restrict__ ep_valid, float* __restrict__ ep_qsum,
                               float* __restrict__ priorities) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float w = weights ? weights[b] : 1.0f;
    float bad = 0.f, sq = 0.f, valid = 0.f, qsum = 0.f, abs_sum = 0.f, abs_max = 0.f;
    for (int t = 0; t < T; ++t) {
        const long long i = (long long)t * B + b;
        const float d = dones_env[i];
        const float q = q_tot[i];
        const float target = rewards[i] + (1.0f - d) * gamma * next_q_tot[i];
        const float keep = 1.0f - bad;
        const float e = (q - target) * keep;
        if (err_out) err_out[i] = e;
        sq = fmaf(e, e, sq);
        valid += keep;
        qsum = fmaf(q, keep, qsum);
        abs_sum += fabsf(e);
        abs_max = fmaxf(abs_max, fabsf(e));
        bad = d;                                   // bad_transitions_mask[t+1] = dones_env[t]
    }
    ep_loss[b] = sq * w;
    ep_valid[b] = valid;
    ep_qsum[b] = qsum;
    if (priorities) priorities[b] = (1.0f - per_nu) * (abs_sum / (float)T) + per_nu * abs_max + per_eps;
}

}  // namespace

extern "C" int mm_q_select(const float* q_all, const float* target_q_all, const float* acts, float* q_taken,
                           float* next_q, int64_t* act_idx, int64_t* greedy, int32_t T, int32_t B, int32_t A,
                           int32_t NA, int32_t double_q, void* stream) {
    if (!q_all || T <= 0 || B <= 0 || A <= 0 || NA <= 0) return MM_ERR_BAD_ARG;
    if ((q_taken || act_idx) && !acts) return MM_ERR_BAD_ARG;
    const long long n = (long long)(T + 1) * A * B;
    q_select_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        q_all, target_q_all, acts, q_taken, next_q, act_idx, greedy, T, B, A, NA, double_q);
    return mm::check_launch();
}

extern "C" int mm_qmix_td(const float* q_tot, const float* next_q_tot, const float* rewards,
                          const float* dones_env, const float* weights, float gamma, float per_nu,
                          float per_eps, int32_t T, int32_t B, float* err, float* ep_loss, float* ep_valid,
                          float* ep_qsum, float* priorities, void* stream) {
    if (!q_tot || !next_q_tot || !rewards || !dones_env || !ep_loss || !ep_valid || !ep_qsum) return MM_ERR_BAD_ARG;
    if (T <= 0 || B <= 0) return MM_ERR_BAD_ARG;
    qmix_td_kernel<<<(B + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        q_tot, next_q_tot, rewards, dones_env, weights, gamma, per_nu, per_eps, T, B, err, ep_loss, ep_valid,
        ep_qsum, priorities);
    return mm::check_launch();
}
