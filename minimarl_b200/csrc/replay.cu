// Episode replay on the device: float64 segment trees (PER), prioritised index sampling, batch gather.
//
// Replaces offpolicy/utils/segment_tree.py (SegmentTree.__setitem__ :74-89, _reduce_helper :40-56,
// SumSegmentTree.find_prefixsum_idx :115-146) and offpolicy/utils/rec_buffer.py
// (RecPolicyBuffer.sample_inds :192-240, PrioritizedRecReplayBuffer.sample :278-304,
// update_priorities :306-324).  Trees are float64 arrays of 2*capacity nodes (node 1 = root,
// leaves at capacity + i), exactly the reference's layout, so index sampling is bit-exact:
// it only uses IEEE float64 add / subtract / compare.
#include "common.cuh"

namespace {

// ---- tree maintenance -------------------------------------------------------------------------
// One block: the update touches n leaves and their <= n*log2(capacity) ancestors; it is latency bound
// (a handful of dependent levels), not bandwidth bound.  Duplicate indices carry equal values in the
// reference's use (same episode -> same TD error), so the write race between them is benign.
__global__ void __launch_bounds__(1024)
segtree_set_kernel(double* __restrict__ sum_tree, double* __restrict__ min_tree, int capacity,
                   const int64_t* __restrict__ idx, const double* __restrict__ val, double scalar_val,
                   int64_t idx_base, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int64_t leaf = (idx ? idx[i] : idx_base + i) + capacity;
        const double v = val ? val[i] : scalar_val;
        if (sum_tree) sum_tree[leaf] = v;
        if (min_tree) min_tree[leaf] = v;
    }
    __syncthreads();
    for (int shift = 1; (capacity >> shift) >= 1; ++shift) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int64_t node = ((idx ? idx[i] : idx_base + i) + capacity) >> shift;
            if (sum_tree) sum_tree[node] = sum_tree[2 * node] + sum_tree[2 * node + 1];
            if (min_tree) min_tree[node] = fmin(min_tree[2 * node], min_tree[2 * node + 1]);
        }
        __syncthreads();
    }
}

// op over leaves [lo, hi] with the association order of the reference's recursive helper
template <bool IS_MIN>
__device__ double range_reduce(const double* tree, int lo, int hi, int node, int node_lo, int node_hi) {
    if (lo == node_lo && hi == node_hi) return tree[node];
    const int mid = (node_lo + node_hi) / 2;
    if (hi <= mid) return range_reduce<IS_MIN>(tree, lo, hi, 2 * node, node_lo, mid);
    if (lo > mid) return range_reduce<IS_MIN>(tree, lo, hi, 2 * node + 1, mid + 1, node_hi);
    const double a = range_reduce<IS_MIN>(tree, lo, mid, 2 * node, node_lo, mid);
    const double b = range_reduce<IS_MIN>(tree, mid + 1, hi, 2 * node + 1, mid + 1, node_hi);
    return IS_MIN ? fmin(a, b) : a + b;
}

__global__ void segtree_reduce_kernel(const double* __restrict__ tree, int capacity, int start, int end,
                                      int is_min, double* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0)
        *out = is_min ? range_reduce<true>(tree, start, end - 1, 1, 0, capacity - 1)
                      : range_reduce<false>(tree, start, end - 1, 1, 0, capacity - 1);
}

__device__ __forceinline__ int64_t descend(const double* __restrict__ tree, int capacity, double mass) {
    int64_t node = 1;
    while (node < capacity) {
        const int64_t left = 2 * node;
        const double v = tree[left];
        if (v <= mass) {      // reference: go right (and subtract) when value[left] <= prefixsum
            mass -= v;
            node = left + 1;
        } else {
            node = left;
        }
    }
    return node - capacity;
}

__global__ void segtree_find_kernel(const double* __restrict__ tree, int capacity,
                                    const double* __restrict__ mass, int64_t* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = descend(tree, capacity, mass[i]);
}

// PrioritizedRecReplayBuffer.sample: mass = u * sum(0, len-1), tree descent, importance weights.
__global__ void per_sample_kernel(const double* __restrict__ sum_tree, const double* __restrict__ min_tree,
                                  int capacity, int len, const double* __restrict__ u, int n, double beta,
                                  const double* __restrict__ beta_dev, int64_t* __restrict__ idx_out, double* __restrict__ w_out) {
    __shared__ double s_total, s_whole, s_max_w;
    if (beta_dev) beta = *beta_dev;                      // a captured update follows the annealing schedule through this scalar
    if (threadIdx.x == 0) {
        // `sum(0, len(self) - 1)`: end is exclusive and the caller already subtracted one (kept as is)
        s_total = range_reduce<false>(sum_tree, 0, len - 2, 1, 0, capacity - 1);
        s_whole = sum_tree[1];
        const double p_min = min_tree[1] / s_whole;
        s_max_w = pow(p_min * (double)len, -beta);
    }
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int64_t leaf = descend(sum_tree, capacity, u[i] * s_total);
        idx_out[i] = leaf;
        const double p = sum_tree[capacity + leaf] / s_whole;
        w_out[i] = pow(p * (double)len, -beta) / s_max_w;
    }
}

// ---- batch gather -------------------------------------------------------------------------------
// Source episodes are time-major [T(+1), S, ...]; the sampled batch is agent-major for the per-agent
// fields ([A, T(+1), B, d], the reference's `_cast`) and time-major for the shared ones.
// One warp moves everything that belongs to one (t, b): 94 + 94 + 10 + 2 + 2 + 1 floats at the
// BASELINE shapes; reads and writes are contiguous runs of D floats.  HBM-bound: 2 x 81 952 B per episode.
struct GatherParams {
    const float *obs, *share_obs, *acts, *rewards, *dones, *dones_env;   // buffer storage
    float *o_obs, *o_share_obs, *o_acts, *o_rewards, *o_dones, *o_dones_env;
    const int64_t* idx;
    int T, S, B, A, D, SD, NA;
    int stacked;   // 0: obs / acts agent-major [A, T(+1), B, d] (the reference's _cast); 1: [T(+1), A, B, d] = the learner's stacked rows
};

// Loads are issued in batches of 4 per lane BEFORE the dependent stores (read-only path, so the compiler
// may hoist them): a warp keeps 4 x 128 B per field in flight instead of one load-store round trip at a time.
template <typename DstIndex>
__device__ __forceinline__ void copy_run(const float* __restrict__ from, int n, float* __restrict__ dst,
                                         DstIndex dst_index, int lane) {
    for (int base = 0; base < n; base += 128) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = base + lane + 32 * u;
            if (k < n) v[u] = __ldg(from + k);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int k = base + lane + 32 * u;
            if (k < n) dst[dst_index(k)] = v[u];
        }
    }
}

// src [TT, S, A, d] at (t, s, :, :) -> dst [A, TT, B, d] at (:, t, b, :), or (stacked) dst [TT, A, B, d] at (t, :, b, :)
__device__ __forceinline__ void copy_agent_major(const float* __restrict__ src, float* __restrict__ dst,
                                                 int t, int b, int64_t s, int S, int B, int A, int d, int TT,
                                                 int lane, bool stacked = false) {
    const float* from = src + ((int64_t)t * S + s) * A * d;
    const int64_t agent_stride = stacked ? (int64_t)B * d : (int64_t)TT * B * d;
    float* to = dst + (stacked ? ((int64_t)t * A * B + b) : ((int64_t)t * B + b)) * d;
    copy_run(from, A * d, to, [=](int k) { const int a = k / d; return (int64_t)a * agent_stride + (k - a * d); }, lane);
}

__global__ void __launch_bounds__(256) replay_gather_kernel(const GatherParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // warp = (t, b), b fastest
    if (w >= (int64_t)(p.T + 1) * p.B) return;
    const int t = (int)(w / p.B), b = (int)(w - (int64_t)t * p.B);
    const int64_t s = p.idx[b];
    copy_agent_major(p.obs, p.o_obs, t, b, s, p.S, p.B, p.A, p.D, p.T + 1, lane, p.stacked != 0);
    copy_run(p.share_obs + ((int64_t)t * p.S + s) * p.SD, p.SD, p.o_share_obs + ((int64_t)t * p.B + b) * p.SD,
             [](int k) { return (int64_t)k; }, lane);
    if (t < p.T) {
        copy_agent_major(p.acts, p.o_acts, t, b, s, p.S, p.B, p.A, p.NA, p.T, lane, p.stacked != 0);
        copy_agent_major(p.rewards, p.o_rewards, t, b, s, p.S, p.B, p.A, 1, p.T, lane);
        copy_agent_major(p.dones, p.o_dones, t, b, s, p.S, p.B, p.A, 1, p.T, lane);
        if (lane == 0) p.o_dones_env[(int64_t)t * p.B + b] = __ldg(p.dones_env + (int64_t)t * p.S + s);
    }
}

// Fast path for the shapes of the hot path (2 agents, D <= 64, SD <= 96, NA <= 16): no integer divisions, every
// load of the (t, b) chunk is issued before the first store, addresses are hoisted.  The generic kernel above
// was issue-bound (75 % issue-active, 311 us for 4 096 x 101 chunks); this one moves the same bytes with
// ~4x fewer instructions.
__global__ void __launch_bounds__(256) replay_gather_fast_kernel(const GatherParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int T1 = p.T + 1;
    if (w >= (int64_t)T1 * p.B) return;
    const int t = (int)(w / p.B), b = (int)(w - (int64_t)t * p.B);
    const int64_t s = __ldg(p.idx + b);
    const int D = p.D, SD = p.SD, NA = p.NA;
    const int64_t src_tb = (int64_t)t * p.S + s;           // (t, episode) in the storage
    const int64_t dst_tb = (int64_t)t * p.B + b;           // (t, batch slot) in the output
    const float* so = p.obs + src_tb * 2 * D;
    const float* ss = p.share_obs + src_tb * SD;
    float o[2][2], sh[3], ac = 0.f, rw = 0.f, dn = 0.f, de = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int k = lane + 32 * u;
            o[a][u] = k < D ? __ldg(so + a * D + k) : 0.f;
        }
#pragma unroll
    for (int u = 0; u < 3; ++u) {
        const int k = lane + 32 * u;
        sh[u] = k < SD ? __ldg(ss + k) : 0.f;
    }
    const bool step = t < p.T;
    if (step) {
        if (lane < 2 * NA) ac = __ldg(p.acts + src_tb * 2 * NA + lane);
        if (lane < 2) { rw = __ldg(p.rewards + src_tb * 2 + lane); dn = __ldg(p.dones + src_tb * 2 + lane); }
        if (lane == 0) de = __ldg(p.dones_env + src_tb);
    }
    const int64_t agent_stride_obs = p.stacked ? (int64_t)p.B * D : (int64_t)T1 * p.B * D;
    float* dobs = p.o_obs + (p.stacked ? (int64_t)t * 2 * p.B + b : dst_tb) * D;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int k = lane + 32 * u;
            if (k < D) dobs[a * agent_stride_obs + k] = o[a][u];
        }
    float* dsh = p.o_share_obs + dst_tb * SD;
#pragma unroll
    for (int u = 0; u < 3; ++u) {
        const int k = lane + 32 * u;
        if (k < SD) dsh[k] = sh[u];
    }
    if (step) {
        const int64_t agent_stride = (int64_t)p.T * p.B;
        if (lane < 2 * NA) {
            const int a = lane >= NA ? 1 : 0;
            const int64_t row = p.stacked ? ((int64_t)t * 2 + a) * p.B + b : a * agent_stride + dst_tb;
            p.o_acts[row * NA + (lane - a * NA)] = ac;
        }
        if (lane < 2) {
            p.o_rewards[lane * agent_stride + dst_tb] = rw;
            p.o_dones[lane * agent_stride + dst_tb] = dn;
        }
        if (lane == 0) p.o_dones_env[dst_tb] = de;
    }
}

}  // namespace

extern "C" int mm_segtree_set(double* sum_tree, double* min_tree, int32_t capacity, const int64_t* idx,
                              const double* values, double scalar_value, int64_t idx_base, int32_t n,
                              void* stream) {
    if ((!sum_tree && !min_tree) || capacity <= 0 || (capacity & (capacity - 1)) || n <= 0) return MM_ERR_BAD_ARG;
    segtree_set_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(sum_tree, min_tree, capacity, idx,
                                                                          values, scalar_value, idx_base, n);
    return mm::check_launch();
}

extern "C" int mm_segtree_reduce(const double* tree, int32_t capacity, int32_t start, int32_t end,
                                 int32_t is_min, double* out, void* stream) {
    if (!tree || !out || capacity <= 0 || start < 0 || end <= start || end > capacity) return MM_ERR_BAD_ARG;
    segtree_reduce_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(tree, capacity, start, end, is_min, out);
    return mm::check_launch();
}

extern "C" int mm_segtree_find_prefixsum(const double* sum_tree, int32_t capacity, const double* mass,
                                         int64_t* idx_out, int32_t n, void* stream) {
    if (!sum_tree || !mass || !idx_out || capacity <= 0 || n <= 0) return MM_ERR_BAD_ARG;
    segtree_find_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(sum_tree, capacity, mass,
                                                                                       idx_out, n);
    return mm::check_launch();
}

extern "C" int mm_per_sample(const double* sum_tree, const double* min_tree, int32_t capacity, int32_t len,
                             const double* uniforms, int32_t n, double beta, int64_t* idx_out,
                             double* weights_out, void* stream) {
    if (!sum_tree || !min_tree || !uniforms || !idx_out || !weights_out) return MM_ERR_BAD_ARG;
    if (capacity <= 0 || len < 2 || len > capacity || n <= 0 || !(beta > 0)) return MM_ERR_BAD_ARG;
    const int blocks = (n + 255) / 256 < 148 ? (n + 255) / 256 : 148;
    per_sample_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(sum_tree, min_tree, capacity, len,
                                                                            uniforms, n, beta, nullptr, idx_out, weights_out);
    return mm::check_launch();
}

// the same with beta read from device memory at run time (a CUDA-graph replay of the update follows the beta schedule)
extern "C" int mm_per_sample_dev(const double* sum_tree, const double* min_tree, int32_t capacity, int32_t len,
                                 const double* uniforms, int32_t n, const double* beta_dev, int64_t* idx_out, double* weights_out,
                                 void* stream) {
    if (!sum_tree || !min_tree || !uniforms || !beta_dev || !idx_out || !weights_out) return MM_ERR_BAD_ARG;
    if (capacity <= 0 || len < 2 || len > capacity || n <= 0) return MM_ERR_BAD_ARG;
    const int blocks = (n + 255) / 256 < 148 ? (n + 255) / 256 : 148;
    per_sample_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(sum_tree, min_tree, capacity, len,
                                                                            uniforms, n, 0.0, beta_dev, idx_out, weights_out);
    return mm::check_launch();
}

static int gather_launch(const float* obs, const float* share_obs, const float* acts,
                         const float* rewards, const float* dones, const float* dones_env,
                         const int64_t* idx, float* out_obs, float* out_share_obs, float* out_acts,
                         float* out_rewards, float* out_dones, float* out_dones_env, int32_t T,
                         int32_t S, int32_t B, int32_t A, int32_t D, int32_t SD, int32_t NA, int stacked,
                         void* stream) {
    if (!obs || !share_obs || !acts || !rewards || !dones || !dones_env || !idx) return MM_ERR_BAD_ARG;
    if (!out_obs || !out_share_obs || !out_acts || !out_rewards || !out_dones || !out_dones_env) return MM_ERR_BAD_ARG;
    if (T <= 0 || S <= 0 || B <= 0 || A <= 0 || D <= 0 || SD <= 0 || NA <= 0) return MM_ERR_BAD_ARG;
    GatherParams p{obs, share_obs, acts, rewards, dones, dones_env, out_obs, out_share_obs, out_acts,
                   out_rewards, out_dones, out_dones_env, idx, T, S, B, A, D, SD, NA, stacked};
    const int64_t warps = (int64_t)(T + 1) * B;
    const int64_t blocks = (warps * 32 + 255) / 256;
    if (A == 2 && D <= 64 && SD <= 96 && NA <= 16)
        replay_gather_fast_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    else
        replay_gather_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    return mm::check_launch();
}

extern "C" int mm_replay_gather(const float* obs, const float* share_obs, const float* acts,
                                const float* rewards, const float* dones, const float* dones_env,
                                const int64_t* idx, float* out_obs, float* out_share_obs, float* out_acts,
                                float* out_rewards, float* out_dones, float* out_dones_env, int32_t T,
                                int32_t S, int32_t B, int32_t A, int32_t D, int32_t SD, int32_t NA,
                                void* stream) {
    return gather_launch(obs, share_obs, acts, rewards, dones, dones_env, idx, out_obs, out_share_obs, out_acts, out_rewards,
                         out_dones, out_dones_env, T, S, B, A, D, SD, NA, 0, stream);
}

// the same with out_obs [T+1, A, B, D] and out_acts [T, A, B, NA]: the row order (t, a * B + b) the learner stacks the agents
// into (qmix.py:96-100), so that torch.cat over the agents becomes a view
extern "C" int mm_replay_gather_stacked(const float* obs, const float* share_obs, const float* acts,
                                        const float* rewards, const float* dones, const float* dones_env,
                                        const int64_t* idx, float* out_obs, float* out_share_obs, float* out_acts,
                                        float* out_rewards, float* out_dones, float* out_dones_env, int32_t T,
                                        int32_t S, int32_t B, int32_t A, int32_t D, int32_t SD, int32_t NA,
                                        void* stream) {
    return gather_launch(obs, share_obs, acts, rewards, dones, dones_env, idx, out_obs, out_share_obs, out_acts, out_rewards,
                         out_dones, out_dones_env, T, S, B, A, D, SD, NA, 1, stream);
}
