// PPO losses of R_MAPPO.ppo_update as two pairs of kernels (mappo/algorithms/ramppo_network.py:70-171):
//   actor:  imp = exp(logp(a) - old_logp);  L = -sum(min(imp * adv, clamp(imp, 1-c, 1+c) * adv) * w) / sum(w),
//           H = sum(entropy * w) / sum(w)      (Categorical(logits), w = active mask or 1)
//   critic: clipped value loss with Huber / MSE error on (normalised) returns, max(original, clipped), masked mean.
// The forward kernels reduce the weighted sums (double accumulators), the backward kernels re-derive the per-row terms
// from the same inputs and emit d/dlogits resp. d/dvalues for upstream gradients of the scalar outputs -- one pass over
// the batch each instead of ~40 elementwise launches.  Tie rules follow autograd: torch.min / torch.max split the
// gradient evenly between equal arguments, clamp passes the gradient on its closed interval.
#include "common.cuh"

namespace {

constexpr int kBlock = 256;

// Per-row inputs are either dense [R] vectors (first == null) or tables addressed like the data chunks of
// recurrent_generator: row r = l * n_chunks + c lives at table entry first[c] + l * step.
struct RowMap {
    const long long* first;
    long long step, n_chunks;
    __device__ __forceinline__ long long operator()(long long r) const {
        if (!first) return r;
        const long long l = r / n_chunks;
        return first[r - l * n_chunks] + l * step;
    }
};
constexpr int kMaxNA = 8;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MM_FULL_MASK, v, o);
    return v;
}

template <int N>
__device__ __forceinline__ void block_accumulate(double (&v)[N], double* out) {
    __shared__ double sm[N][kBlock / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double s = warp_sum(v[i]);
        if (lane == 0) sm[i][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < N) {
        double s = 0.0;
        for (int w = 0; w < kBlock / 32; ++w) s += sm[threadIdx.x][w];
        atomicAdd(out + threadIdx.x, s);
    }
}

struct ActorRow {
    float p[kMaxNA], lse, logp_act, ent;
    int act;
};

__device__ __forceinline__ ActorRow actor_row(const float* __restrict__ logits, long long r, int NA, float action) {
    ActorRow o;
    float z[kMaxNA], mx = -INFINITY;
#pragma unroll
    for (int a = 0; a < kMaxNA; ++a) {
        z[a] = a < NA ? logits[r * NA + a] : -INFINITY;
        mx = fmaxf(mx, z[a]);
    }
    float s = 0.f;
#pragma unroll
    for (int a = 0; a < kMaxNA; ++a) {
        o.p[a] = a < NA ? expf(z[a] - mx) : 0.f;
        s += o.p[a];
    }
    o.lse = mx + logf(s);
    const float inv = 1.0f / s;
    o.act = (int)action;
    o.ent = 0.f;
    o.logp_act = 0.f;
#pragma unroll
    for (int a = 0; a < kMaxNA; ++a) {
        o.p[a] *= inv;
        const float lp = z[a] - o.lse;                       // normalised logit, as torch.distributions.Categorical keeps it
        if (a < NA) o.ent -= o.p[a] * lp;
        if (a == o.act) o.logp_act = lp;
    }
    return o;
}

// sums[4] += {sum min(surr1, surr2) w, sum entropy w, sum imp, sum w}
__global__ void __launch_bounds__(kBlock) actor_fwd_kernel(const float* __restrict__ logits, const float* __restrict__ actions,
                                                          const float* __restrict__ old_logp, const float* __restrict__ adv,
                                                          const float* __restrict__ active, float* __restrict__ logp_out,
                                                          long long R, int NA, float clip, double* __restrict__ sums, RowMap map) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (long long r = (long long)blockIdx.x * kBlock + threadIdx.x; r < R; r += (long long)gridDim.x * kBlock) {
        const long long i = map(r);
        const ActorRow o = actor_row(logits, r, NA, actions[i]);
        const float w = active ? active[i] : 1.0f;
        const float imp = expf(o.logp_act - old_logp[i]);
        const float s1 = imp * adv[i], s2 = fminf(fmaxf(imp, 1.0f - clip), 1.0f + clip) * adv[i];
        if (logp_out) logp_out[r] = o.logp_act;
        acc[0] += (double)(fminf(s1, s2) * w);
        acc[1] += (double)(o.ent * w);
        acc[2] += (double)imp;
        acc[3] += (double)w;
    }
    block_accumulate(acc, sums);
}

// dlogits = g_min * d(min(surr) w)/dlogits + g_ent * d(entropy w)/dlogits      (g_* = upstream / sum(w), from the host)
__global__ void __launch_bounds__(kBlock) actor_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ actions,
                                                          const float* __restrict__ old_logp, const float* __restrict__ adv,
                                                          const float* __restrict__ active, const float* __restrict__ g,
                                                          float* __restrict__ dlogits, long long R, int NA, float clip, RowMap map) {
    const float g_min = g[0], g_ent = g[1];
    for (long long r = (long long)blockIdx.x * kBlock + threadIdx.x; r < R; r += (long long)gridDim.x * kBlock) {
        const long long i = map(r);
        const ActorRow o = actor_row(logits, r, NA, actions[i]);
        const float w = active ? active[i] : 1.0f;
        const float imp = expf(o.logp_act - old_logp[i]), A = adv[i];
        const float lo = 1.0f - clip, hi = 1.0f + clip;
        const float s1 = imp * A, s2 = fminf(fmaxf(imp, lo), hi) * A;
        const bool inside = imp >= lo && imp <= hi;          // clamp passes the gradient on the closed interval
        // d min / d logp(a): through surr1 always, through surr2 only when unclamped; ties share the gradient evenly
        const float via1 = s1 < s2 ? 1.0f : (s1 == s2 ? 0.5f : 0.0f);
        const float via2 = s2 < s1 ? 1.0f : (s1 == s2 ? 0.5f : 0.0f);
        const float dlogp = (via1 + (inside ? via2 : 0.0f)) * A * imp * g_min * w;
        const float ge = g_ent * w;
#pragma unroll
        for (int a = 0; a < kMaxNA; ++a) {
            if (a < NA) {
                const float lp = logits[r * NA + a] - o.lse;
                const float d_logp = ((a == o.act) ? 1.0f : 0.0f) - o.p[a];
                const float d_ent = -o.p[a] * (lp + o.ent);
                dlogits[r * NA + a] = dlogp * d_logp + ge * d_ent;
            }
        }
    }
}

struct ValueCfg {
    float clip, delta;
    int huber, clipped;
};

__device__ __forceinline__ float err_loss(float e, const ValueCfg& c) {
    if (!c.huber) return e * e / 2;
    const float ae = fabsf(e);
    return ae <= c.delta ? e * e / 2 : c.delta * (ae - c.delta / 2);
}
__device__ __forceinline__ float err_grad(float e, const ValueCfg& c) {
    if (!c.huber) return e;
    return fabsf(e) <= c.delta ? e : copysignf(c.delta, e);
}

// norm = {mean, 1/sqrt(var)} (device) or NULL;  sums[2] += {sum loss w, sum w}
__global__ void __launch_bounds__(kBlock) value_fwd_kernel(const float* __restrict__ values, const float* __restrict__ value_preds,
                                                          const float* __restrict__ returns, const float* __restrict__ active,
                                                          const float* __restrict__ norm, long long R, ValueCfg c,
                                                          double* __restrict__ sums, RowMap map) {
    const float mean = norm ? norm[0] : 0.f, istd = norm ? norm[1] : 1.f;
    double acc[2] = {0.0, 0.0};
    for (long long r = (long long)blockIdx.x * kBlock + threadIdx.x; r < R; r += (long long)gridDim.x * kBlock) {
        const long long i = map(r);
        const float v = values[r], vp = value_preds[i];
        const float ret = norm ? (returns[i] - mean) * istd : returns[i];
        const float vc = vp + fminf(fmaxf(v - vp, -c.clip), c.clip);
        const float lo = err_loss(ret - v, c), lc = err_loss(ret - vc, c);
        const float w = active ? active[i] : 1.0f;
        acc[0] += (double)((c.clipped ? fmaxf(lo, lc) : lo) * w);
        acc[1] += (double)w;
    }
    block_accumulate(acc, sums);
}

__global__ void __launch_bounds__(kBlock) value_bwd_kernel(const float* __restrict__ values, const float* __restrict__ value_preds,
                                                          const float* __restrict__ returns, const float* __restrict__ active,
                                                          const float* __restrict__ norm, const float* __restrict__ g,
                                                          float* __restrict__ dvalues, long long R, ValueCfg c, RowMap map) {
    const float mean = norm ? norm[0] : 0.f, istd = norm ? norm[1] : 1.f, gs = g[0];
    for (long long r = (long long)blockIdx.x * kBlock + threadIdx.x; r < R; r += (long long)gridDim.x * kBlock) {
        const long long i = map(r);
        const float v = values[r], vp = value_preds[i];
        const float ret = norm ? (returns[i] - mean) * istd : returns[i];
        const float d = v - vp;
        const float vc = vp + fminf(fmaxf(d, -c.clip), c.clip);
        const float eo = ret - v, ec = ret - vc;
        const float lo = err_loss(eo, c), lc = err_loss(ec, c);
        const float go = -err_grad(eo, c);
        const float gc = (d >= -c.clip && d <= c.clip) ? -err_grad(ec, c) : 0.0f;
        float gv = go;
        if (c.clipped) gv = lo > lc ? go : (lo < lc ? gc : 0.5f * (go + gc));
        dvalues[r] = gv * gs * (active ? active[i] : 1.0f);
    }
}

int grid_for(long long R) {
    const long long blocks = (R + kBlock - 1) / kBlock;
    return (int)(blocks < 148 * 8 ? blocks : 148 * 8);
}

}  // namespace

extern "C" int mm_ppo_actor_loss_fwd(const float* logits, const float* actions, const float* old_logp, const float* adv,
                                     const float* active, float* logp_out, int64_t R, int32_t NA, float clip, double* sums4,
                                     const int64_t* first, int64_t step_rows, int64_t n_chunks, void* stream) {
    if (!logits || !actions || !old_logp || !adv || !sums4 || R <= 0) return MM_ERR_BAD_ARG;
    if (NA < 1 || NA > kMaxNA) return MM_ERR_UNSUPPORTED;
    if (first && (n_chunks <= 0 || R % n_chunks)) return MM_ERR_BAD_ARG;
    const RowMap map{reinterpret_cast<const long long*>(first), step_rows, n_chunks};
    actor_fwd_kernel<<<grid_for(R), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(logits, actions, old_logp, adv, active,
                                                                                  logp_out, R, NA, clip, sums4, map);
    return mm::check_launch();
}

extern "C" int mm_ppo_actor_loss_bwd(const float* logits, const float* actions, const float* old_logp, const float* adv,
                                     const float* active, const float* g2, float* dlogits, int64_t R, int32_t NA, float clip,
                                     const int64_t* first, int64_t step_rows, int64_t n_chunks, void* stream) {
    if (!logits || !actions || !old_logp || !adv || !g2 || !dlogits || R <= 0) return MM_ERR_BAD_ARG;
    if (NA < 1 || NA > kMaxNA) return MM_ERR_UNSUPPORTED;
    if (first && (n_chunks <= 0 || R % n_chunks)) return MM_ERR_BAD_ARG;
    const RowMap map{reinterpret_cast<const long long*>(first), step_rows, n_chunks};
    actor_bwd_kernel<<<grid_for(R), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(logits, actions, old_logp, adv, active, g2,
                                                                                  dlogits, R, NA, clip, map);
    return mm::check_launch();
}

extern "C" int mm_ppo_value_loss_fwd(const float* values, const float* value_preds, const float* returns, const float* active,
                                     const float* norm2, int64_t R, float clip, float huber_delta, int32_t use_huber,
                                     int32_t use_clipped, double* sums2, const int64_t* first, int64_t step_rows,
                                     int64_t n_chunks, void* stream) {
    if (!values || !value_preds || !returns || !sums2 || R <= 0) return MM_ERR_BAD_ARG;
    if (first && (n_chunks <= 0 || R % n_chunks)) return MM_ERR_BAD_ARG;
    const ValueCfg c{clip, huber_delta, use_huber, use_clipped};
    const RowMap map{reinterpret_cast<const long long*>(first), step_rows, n_chunks};
    value_fwd_kernel<<<grid_for(R), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(values, value_preds, returns, active, norm2, R,
                                                                                  c, sums2, map);
    return mm::check_launch();
}

extern "C" int mm_ppo_value_loss_bwd(const float* values, const float* value_preds, const float* returns, const float* active,
                                     const float* norm2, const float* g1, float* dvalues, int64_t R, float clip,
                                     float huber_delta, int32_t use_huber, int32_t use_clipped, const int64_t* first,
                                     int64_t step_rows, int64_t n_chunks, void* stream) {
    if (!values || !value_preds || !returns || !g1 || !dvalues || R <= 0) return MM_ERR_BAD_ARG;
    if (first && (n_chunks <= 0 || R % n_chunks)) return MM_ERR_BAD_ARG;
    const ValueCfg c{clip, huber_delta, use_huber, use_clipped};
    const RowMap map{reinterpret_cast<const long long*>(first), step_rows, n_chunks};
    value_bwd_kernel<<<grid_for(R), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(values, value_preds, returns, active, norm2, g1,
                                                                                  dvalues, R, c, map);
    return mm::check_launch();
}
