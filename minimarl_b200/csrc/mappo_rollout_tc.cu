// Fused MAPPO rollout: actor + critic (tcgen05) + categorical sampling + Checkers step + rollout-buffer writes,
// T lockstep steps of all envs in ONE launch.
//
// Replaces the per-step body of MAGYM_Runner.run / collect / insert (mappo/runner/shared/magym_runner.py:41-98,
// 113-195) with R_MAPPOPolicy.get_actions (mappo/algorithms/rmappo_policy.py:57-95): R_Actor.forward and
// R_Critic.forward (r_actor_critic.py:60-93,189-208) are each LayerNorm(47) -> [Linear ReLU LN] x2 -> GRU(32) ->
// LN -> head, with hidden size 32 (mappo/_config.py:176-249) and the local-observation critic of the default
// configuration (share_obs == obs).  Everything the reference writes into SharedReplayBuffer per step
// (shared_buffer.py:82-117) is written here directly: obs / share_obs [t+1], rnn_states / rnn_states_critic [t+1],
// actions / action_log_probs / value_preds / rewards [t], masks [t+1].
//
// Layout: one CTA = 128 (env, agent) rows = 64 boards, 512 threads.  The two networks sit side by side in the
// accumulator columns: thread (row, part) owns 16 of the 2 x 32 hidden columns of its row -- parts 0,1 the actor,
// parts 2,3 the critic.  The first layer of both networks is ONE MMA (N = 64) on a shared operand: LayerNorm(47)'s
// affine is folded into the first Linear on the host (W1' = W1 diag(gamma), b1' = b1 + W1 beta), so both nets read
// the same normalised observation tile.  The later layers are two MMAs (one per network) per phase, committed
// together.  After the heads the first four warps sample actions by inverse CDF on softmax(logits) with a Philox
// uniform, step their boards (checkers_core.cuh, auto-reset on done), and expand the next observations into the
// shared obs tile; hidden states are zeroed where the episode ended (masks) before they are published for step t+1.
#include "checkers_core.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kRows = 128;
constexpr int kThreads = 512;
constexpr int H = 32;               // hidden size of both networks
constexpr int D = ck::kObs;         // 47
constexpr int K1 = 48;              // D rounded up to the MMA K step
constexpr int NA = 5;
constexpr int CPT = 16;             // columns per thread
constexpr float kEps = 1e-5f;

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// MUFU.TANH (one special-function op; abs error ~5e-4, below the tf32 operand rounding); sigmoid through the same unit --
// the same activations as the PPO trunk kernels that re-evaluate these networks during the update
__device__ __forceinline__ float tanh_(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_(float x) { return fmaf(tanh_(0.5f * x), 0.5f, 0.5f); }

// Parameter buffer (floats), per network k in {0 actor, 1 critic} unless noted:
//   w1 [2][H][K1] (LayerNorm(47) affine folded in, zero-padded), w2 [2][H][H], wih [2][3H][H], whh [2][3H][H]   (matrices)
//   b1 [2][H] (folded), ln1_w, ln1_b, b2, ln2_w, ln2_b [2][H], bih, bhh [2][3H], ln3_w, ln3_b [2][H],
//   head_w: actor [NA][H] then critic [H]; head_b: actor [NA] then critic [1]                                     (small)
struct Layout {
    int w1, w2, wih, whh, small, b1, ln1_w, ln1_b, b2, ln2_w, ln2_b, bih, bhh, ln3_w, ln3_b, wa, wv, ba, bv, total;
};
__host__ __device__ inline Layout make_layout() {
    Layout L;
    int o = 0;
    auto take = [&](int n) { int at = o; o += (n + 3) & ~3; return at; };
    L.w1 = take(2 * H * K1); L.w2 = take(2 * H * H); L.wih = take(2 * 3 * H * H); L.whh = take(2 * 3 * H * H);
    L.small = o;
    L.b1 = take(2 * H); L.ln1_w = take(2 * H); L.ln1_b = take(2 * H);
    L.b2 = take(2 * H); L.ln2_w = take(2 * H); L.ln2_b = take(2 * H);
    L.bih = take(2 * 3 * H); L.bhh = take(2 * 3 * H);
    L.ln3_w = take(2 * H); L.ln3_b = take(2 * H);
    L.wa = take(NA * H); L.wv = take(H); L.ba = take(NA); L.bv = take(1);
    L.total = o;
    return L;
}

struct Args {
    const float* params;
    uint32_t* env_state;                       // [3][N]
    float *obs, *share_obs;                    // [T+1][N][2][47]   (slab 0 is read unless do_reset)
    float *rnn_a, *rnn_c;                      // [T+1][N][2][1][32]
    float *actions, *log_probs, *values, *rewards;   // [T][N][2][1]
    float* masks;                              // [T+1][N][2][1]
    int T;
    long long n_envs;
    int do_reset;
    unsigned long long step_offset;
    uint2 seed;
    mm_checkers_cfg cfg;
};

// LayerNorm over the 32 columns of one network's row, split over 2 threads (parts 2k, 2k+1)
__device__ __forceinline__ void layer_norm_pair(float (&v)[CPT], float2* red, int row, int part, const float* __restrict__ w,
                                                const float* __restrict__ b) {
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int j = 0; j < CPT; ++j) { s += v[j]; q = fmaf(v[j], v[j], q); }
    red[row * 4 + part] = make_float2(s, q);
    __syncthreads();
    const float4 two = *reinterpret_cast<const float4*>(red + row * 4 + (part & 2));
    const float mean = (two.x + two.z) * (1.0f / H);
    const float var = fmaxf((two.y + two.w) * (1.0f / H) - mean * mean, 0.f);
    const float rstd = rsqrtf(var + kEps);
#pragma unroll
    for (int j = 0; j < CPT; ++j) v[j] = fmaf((v[j] - mean) * rstd, w[j], b[j]);
}

__device__ __forceinline__ void store16_kmajor(uint8_t* tile, int r, int k0, int K, const float (&v)[CPT], bool round) {
#pragma unroll
    for (int k = 0; k < CPT; k += 4)
        *reinterpret_cast<float4*>(tile + tc::kmajor_offset(r, k0 + k, K)) =
            round ? make_float4(to_tf32(v[k]), to_tf32(v[k + 1]), to_tf32(v[k + 2]), to_tf32(v[k + 3]))
                  : make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
}

__global__ void __launch_bounds__(kThreads, 1) mappo_rollout_kernel(const __grid_constant__ Args a) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const Layout L = make_layout();
    // ---- shared memory -----------------------------------------------------------------------------------------
    uint8_t* sW1 = smem;                               // [2H x K1]  rows 0..31 actor, 32..63 critic
    uint8_t* sW2 = sW1 + 2 * H * K1 * 4;               // [2][H x H]
    uint8_t* sWih = sW2 + 2 * H * H * 4;               // [2][3H x H]
    uint8_t* sWhh = sWih + 2 * 3 * H * H * 4;          // [2][3H x H]
    uint8_t* sX = sWhh + 2 * 3 * H * H * 4;            // normalised obs tile [128 x K1]
    uint8_t* sA = sX + kRows * K1 * 4;                 // [2][128 x H] activation tiles of the two nets
    uint8_t* sH = sA + 2 * kRows * H * 4;              // [2][128 x H] hidden states
    float* sObs = reinterpret_cast<float*>(sH + 2 * kRows * H * 4);        // [128][47]
    float2* red = reinterpret_cast<float2*>(sObs + kRows * D);             // [128][4]
    float* sHead = reinterpret_cast<float*>(red + kRows * 4);              // [NA + 1][2][128] partial head sums
    float* sDone = sHead + (NA + 1) * 2 * kRows;                           // [128] keep-mask of the row's env
    float* sP = sDone + kRows - L.small;                                   // small params, biased by L.small

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = (warp & 3) * 32 + lane, part = warp >> 2;
    const int net = part >> 1, j0 = (part & 1) * CPT;                      // this thread: columns j0..j0+15 of `net`
    const int c0 = part * CPT;                                             // its column among the 64 side-by-side columns
    const long long row0 = (long long)blockIdx.x * kRows;
    const long long R = 2 * a.n_envs;
    const int n_valid = (int)min((long long)kRows, R - row0);
    const bool valid = row < n_valid;
    const long long grow = row0 + row;
    const long long env_id = grow >> 1;

    if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
    if (tid == 0) tc::mbar_init(&bar, 1);
    auto stage = [&](uint8_t* dst, int off, int rows, int K) {
        for (int i = tid; i < rows * K; i += kThreads) {
            const int r = i / K, k = i - r * K;
            *reinterpret_cast<float*>(dst + tc::kmajor_offset(r, k, K)) = to_tf32(a.params[off + i]);
        }
    };
    stage(sW1, L.w1, 2 * H, K1);
    for (int k = 0; k < 2; ++k) {
        stage(sW2 + k * H * H * 4, L.w2 + k * H * H, H, H);
        stage(sWih + k * 3 * H * H * 4, L.wih + k * 3 * H * H, 3 * H, H);
        stage(sWhh + k * 3 * H * H * 4, L.whh + k * 3 * H * H, 3 * H, H);
    }
    for (int i = L.small + tid; i < L.total; i += kThreads) sP[i] = a.params[i];

    // env registers (part 0 threads = warps 0..3; adjacent lanes are the two agents of a board)
    ck::EnvRegs env;
    env.food = env.cells = env.step = 0;
    if (part == 0) {
        const long long e = valid ? env_id : a.n_envs - 1;
        if (a.do_reset) ck::reset_regs(env);
        else {
            env.food = a.env_state[e];
            env.cells = a.env_state[a.n_envs + e];
            env.step = a.env_state[2 * a.n_envs + e];
        }
        if (a.do_reset) ck::fill_tile(sObs, row, row & 1, env, a.cfg);
    }
    // hidden state of this thread's network / columns, from rnn_states[0] (times masks[0])
    float h[CPT];
    {
        const float* src = (net == 0 ? a.rnn_a : a.rnn_c) + grow * H + j0;
        const float m0 = valid ? a.masks[grow] : 0.f;
#pragma unroll
        for (int k = 0; k < CPT; k += 4) {
            float4 q = valid ? *reinterpret_cast<const float4*>(src + k) : make_float4(0.f, 0.f, 0.f, 0.f);
            h[k] = q.x * m0; h[k + 1] = q.y * m0; h[k + 2] = q.z * m0; h[k + 3] = q.w * m0;
        }
        store16_kmajor(sH + net * kRows * H * 4, row, j0, H, h, false);
    }
    __syncthreads();
    if (a.do_reset) {
        for (int i = tid; i < n_valid * D; i += kThreads) {
            const float v = sObs[i];
            a.obs[row0 * D + i] = v;
            a.share_obs[row0 * D + i] = v;
        }
    } else {
        for (int i = tid; i < n_valid * D; i += kThreads) sObs[i] = a.obs[row0 * D + i];
    }
    tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t phase = 0;

    auto desc = [&](const uint8_t* base, int row_off, int K) {
        return tc::make_desc(tc::smem_u32(base) + (uint32_t)(row_off >> 3) * (uint32_t)(K >> 2) * 128u, 128u,
                             (uint32_t)(K >> 2) * 128u);
    };
    const uint64_t dX = desc(sX, 0, K1), dW1 = desc(sW1, 0, K1);
    uint64_t dA[2], dHd[2], dW2[2], dWih_rz[2], dWih_n[2], dWhh_rz[2], dWhh_n[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        dA[k] = desc(sA + k * kRows * H * 4, 0, H);
        dHd[k] = desc(sH + k * kRows * H * 4, 0, H);
        dW2[k] = desc(sW2 + k * H * H * 4, 0, H);
        dWih_rz[k] = desc(sWih + k * 3 * H * H * 4, 0, H);
        dWih_n[k] = desc(sWih + k * 3 * H * H * 4, 2 * H, H);
        dWhh_rz[k] = desc(sWhh + k * 3 * H * H * 4, 0, H);
        dWhh_n[k] = desc(sWhh + k * 3 * H * H * 4, 2 * H, H);
    }
    const uint32_t id64 = tc::make_idesc_tf32(128, 64), id32 = tc::make_idesc_tf32(128, 32);
    // TMEM columns: [0,64) fc1 / fc2 (actor | critic); rz of net k at 64 + 64k; i_n of net k at 192 + 32k; h_n at 256 + 32k
    auto publish = [&]() {
        tc::fence_smem_to_async();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    };
    auto wait_mma = [&]() {
        if (warp == 0) tc::mbar_wait(&bar, phase);
        phase ^= 1;
        __syncthreads();
        tc::fence_after_sync();
    };
    const float* ln1_w = sP + L.ln1_w + net * H + j0; const float* ln1_b = sP + L.ln1_b + net * H + j0;
    const float* ln2_w = sP + L.ln2_w + net * H + j0; const float* ln2_b = sP + L.ln2_b + net * H + j0;
    const float* ln3_w = sP + L.ln3_w + net * H + j0; const float* ln3_b = sP + L.ln3_b + net * H + j0;
    const size_t slab_obs = (size_t)R * D, slab_h = (size_t)R * H, slab_1 = (size_t)R;

    for (int t = 0; t < a.T; ++t) {
        // ---- A. plain LayerNorm statistics of the obs row (affine folded into W1'), shared by both nets --------
        {
            const float* o = sObs + row * D;
            float s = 0.f, qv = 0.f;
            for (int g = part; g < K1 / 4; g += 4)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k = 4 * g + u;
                    const float x = (valid && k < D) ? o[k] : 0.f;
                    s += x; qv = fmaf(x, x, qv);
                }
            red[row * 4 + part] = make_float2(s, qv);
            __syncthreads();
            const float4 p01 = *reinterpret_cast<const float4*>(red + row * 4);
            const float4 p23 = *reinterpret_cast<const float4*>(red + row * 4 + 2);
            const float mean = (p01.x + p01.z + p23.x + p23.z) * (1.0f / D);
            const float var = fmaxf((p01.y + p01.w + p23.y + p23.w) * (1.0f / D) - mean * mean, 0.f);
            const float rstd = rsqrtf(var + kEps);
            for (int g = part; g < K1 / 4; g += 4) {
                float x[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k = 4 * g + u;
                    x[u] = k < D ? to_tf32(((valid ? o[k] : 0.f) - mean) * rstd) : 0.f;
                }
                *reinterpret_cast<float4*>(sX + tc::kmajor_offset(row, 4 * g, K1)) = make_float4(x[0], x[1], x[2], x[3]);
            }
        }
        publish();
        if (tid == 0) {
#pragma unroll
            for (int kk = 0; kk < K1 / 8; ++kk) tc::mma_tf32(tmem, dX + 16 * kk, dW1 + 16 * kk, id64, kk > 0);
            tc::mma_commit(&bar);
            // the hidden-state halves of both nets' gates only need h_{t-1}: issued now, outside fc1's commit, they run under the
            // fc1 / fc2 epilogues (columns [64, 320) are not touched until the gate phase)
#pragma unroll
            for (int k = 0; k < 2; ++k) {
#pragma unroll
                for (int kk = 0; kk < H / 8; ++kk) tc::mma_tf32(tmem + 64 + 64 * k, dHd[k] + 16 * kk, dWhh_rz[k] + 16 * kk, id64, kk > 0);
#pragma unroll
                for (int kk = 0; kk < H / 8; ++kk) tc::mma_tf32(tmem + 256 + 32 * k, dHd[k] + 16 * kk, dWhh_n[k] + 16 * kk, id32, kk > 0);
            }
        }
        wait_mma();
        // ---- B. fc1 epilogue -> sA[net] ----------------------------------------------------------------------------
        float v[CPT];
        tc::tmem_ld16(lane_addr + c0, v);
#pragma unroll
        for (int j = 0; j < CPT; ++j) v[j] = fmaxf(v[j] + sP[L.b1 + net * H + j0 + j], 0.f);
        layer_norm_pair(v, red, row, part, ln1_w, ln1_b);
        store16_kmajor(sA + net * kRows * H * 4, row, j0, H, v, true);
        publish();
        if (tid == 0) {
#pragma unroll
            for (int k = 0; k < 2; ++k)
#pragma unroll
                for (int kk = 0; kk < H / 8; ++kk) tc::mma_tf32(tmem + 32 * k, dA[k] + 16 * kk, dW2[k] + 16 * kk, id32, kk > 0);
            tc::mma_commit(&bar);
        }
        wait_mma();
        // ---- C. fc2 epilogue -> sA[net]; gate GEMMs of both nets ----------------------------------------------------
        tc::tmem_ld16(lane_addr + c0, v);
#pragma unroll
        for (int j = 0; j < CPT; ++j) v[j] = fmaxf(v[j] + sP[L.b2 + net * H + j0 + j], 0.f);
        layer_norm_pair(v, red, row, part, ln2_w, ln2_b);
        store16_kmajor(sA + net * kRows * H * 4, row, j0, H, v, true);
        publish();
        if (tid == 0) {
#pragma unroll
            for (int k = 0; k < 2; ++k) {
#pragma unroll
                for (int kk = 0; kk < H / 8; ++kk) tc::mma_tf32(tmem + 64 + 64 * k, dA[k] + 16 * kk, dWih_rz[k] + 16 * kk, id64, true);
#pragma unroll
                for (int kk = 0; kk < H / 8; ++kk) tc::mma_tf32(tmem + 192 + 32 * k, dA[k] + 16 * kk, dWih_n[k] + 16 * kk, id32, kk > 0);
            }
            tc::mma_commit(&bar);
        }
        wait_mma();
        // ---- D. GRU gates -> new hidden (registers), LayerNorm, heads ---------------------------------------------
        {
            float gr[CPT], gz[CPT], gin[CPT], ghn[CPT];
            tc::tmem_ld16_issue(lane_addr + 64 + 64 * net + j0, gr);       // four loads in flight, one wait
            tc::tmem_ld16_issue(lane_addr + 64 + 64 * net + 32 + j0, gz);
            tc::tmem_ld16_issue(lane_addr + 192 + 32 * net + j0, gin);
            tc::tmem_ld16_issue(lane_addr + 256 + 32 * net + j0, ghn);
            tc::tmem_ld_fence<16>(gr, gz, gin, ghn);
            const float* bi = sP + L.bih + net * 3 * H + j0;
            const float* bh = sP + L.bhh + net * 3 * H + j0;
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const float rr = sigmoid_(gr[j] + bi[j] + bh[j]);
                const float zz = sigmoid_(gz[j] + bi[H + j] + bh[H + j]);
                const float nn = tanh_(gin[j] + bi[2 * H + j] + rr * (ghn[j] + bh[2 * H + j]));
                h[j] = fmaf(zz, h[j] - nn, nn);
                v[j] = h[j];
            }
        }
        layer_norm_pair(v, red, row, part, ln3_w, ln3_b);
        // heads: actor 5 logits, critic 1 value; partial dots over this thread's 16 columns, pairs meet in smem
        if (net == 0) {
#pragma unroll
            for (int o = 0; o < NA; ++o) {
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < CPT; ++j) acc = fmaf(v[j], sP[L.wa + o * H + j0 + j], acc);
                sHead[(o * 2 + (part & 1)) * kRows + row] = acc;
            }
        } else {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < CPT; ++j) acc = fmaf(v[j], sP[L.wv + j0 + j], acc);
            sHead[(NA * 2 + (part & 1)) * kRows + row] = acc;
        }
        __syncthreads();
        float keep = 1.0f;
        if (part == 0) {
            // ---- E. sample, step the env, write the per-step scalars (warps 0..3) --------------------------------
            float logit[NA], mx = -INFINITY;
#pragma unroll
            for (int o = 0; o < NA; ++o) {
                logit[o] = sP[L.ba + o] + sHead[(o * 2) * kRows + row] + sHead[(o * 2 + 1) * kRows + row];
                mx = fmaxf(mx, logit[o]);
            }
            const float value = sP[L.bv] + sHead[(NA * 2) * kRows + row] + sHead[(NA * 2 + 1) * kRows + row];
            float p[NA], z = 0.f;
#pragma unroll
            for (int o = 0; o < NA; ++o) { p[o] = expf(logit[o] - mx); z += p[o]; }
            const unsigned long long ctr = a.step_offset + (unsigned long long)t;
            const uint4 rnd = mm::philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)grow, (uint32_t)(grow >> 32)), a.seed);
            const float u = (float)(rnd.x >> 8) * (1.0f / 16777216.0f) * z;     // inverse CDF on the unnormalised masses
            int act = NA - 1;                                                    // smallest o with cumsum(p)[o] > u
            float cum = 0.f;
#pragma unroll
            for (int o = 0; o < NA - 1; ++o) {
                cum += p[o];
                if (u < cum && act == NA - 1) act = o;
            }
            const float logp = (logit[act] - mx) - logf(z);
            const int agent = row & 1;
            float reward;
            uint32_t done;
            bool bad;
            ck::transition(env, act, agent, lane, a.cfg, reward, done, bad);
            if (done) ck::reset_regs(env);                                       // auto-reset: next obs is the reset obs
            keep = done ? 0.f : 1.f;
            sDone[row] = keep;
            if (valid) {
                const size_t o1 = (size_t)t * slab_1 + grow;
                a.actions[o1] = (float)act;
                a.log_probs[o1] = logp;
                a.values[o1] = value;
                a.rewards[o1] = reward;
                a.masks[o1 + slab_1] = keep;
            }
            ck::fill_tile(sObs, row, agent, env, a.cfg);
        }
        __syncthreads();
        // ---- F. publish hidden states (zeroed where the episode ended) and stream obs[t+1] --------------------
        keep = sDone[row];
#pragma unroll
        for (int j = 0; j < CPT; ++j) h[j] *= keep;
        store16_kmajor(sH + net * kRows * H * 4, row, j0, H, h, false);
        if (valid) {
            float* dst = (net == 0 ? a.rnn_a : a.rnn_c) + (size_t)(t + 1) * slab_h + grow * H + j0;
#pragma unroll
            for (int k = 0; k < CPT; k += 4) *reinterpret_cast<float4*>(dst + k) = make_float4(h[k], h[k + 1], h[k + 2], h[k + 3]);
        }
        {
            float* d1 = a.obs + (size_t)(t + 1) * slab_obs + row0 * D;
            float* d2 = a.share_obs + (size_t)(t + 1) * slab_obs + row0 * D;
            const int n = n_valid * D;
            if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(d1) | reinterpret_cast<uintptr_t>(d2)) & 15) == 0) {
                for (int i = tid; i < (n >> 2); i += kThreads) {
                    const float4 q = reinterpret_cast<const float4*>(sObs)[i];
                    mm::st_stream(reinterpret_cast<float4*>(d1) + i, q);
                    mm::st_stream(reinterpret_cast<float4*>(d2) + i, q);
                }
            } else {
                for (int i = tid; i < n; i += kThreads) { d1[i] = sObs[i]; d2[i] = sObs[i]; }
            }
        }
        __syncthreads();          // obs tile and hidden tiles complete before step t+1 reads them (its publish() fences them)
    }
    if (part == 0 && valid && (row & 1) == 0) {
        a.env_state[env_id] = env.food;
        a.env_state[a.n_envs + env_id] = env.cells;
        a.env_state[2 * a.n_envs + env_id] = env.step;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

size_t rollout_smem() {
    const Layout L = make_layout();
    return (size_t)4 * (2 * H * K1 + 2 * H * H + 12 * H * H + kRows * K1 + 4 * kRows * H + kRows * D + kRows * 8 +
                        (NA + 1) * 2 * kRows + kRows + (L.total - L.small)) + 128;
}

}  // namespace

extern "C" int mm_mappo_rollout_param_count(void) { return make_layout().total; }

// offsets[19]: w1, w2, wih, whh, b1, ln1_w, ln1_b, b2, ln2_w, ln2_b, bih, bhh, ln3_w, ln3_b, wa, wv, ba, bv, total
extern "C" int mm_mappo_rollout_param_offsets(int32_t* o) {
    if (!o) return MM_ERR_BAD_ARG;
    const Layout L = make_layout();
    const int v[19] = {L.w1, L.w2, L.wih, L.whh, L.b1, L.ln1_w, L.ln1_b, L.b2, L.ln2_w, L.ln2_b, L.bih, L.bhh, L.ln3_w,
                       L.ln3_b, L.wa, L.wv, L.ba, L.bv, L.total};
    for (int i = 0; i < 19; ++i) o[i] = v[i];
    return MM_OK;
}

extern "C" int mm_mappo_rollout_tc(const float* params, uint32_t* env_state, float* obs, float* share_obs, float* rnn_states,
                                   float* rnn_states_critic, float* actions, float* action_log_probs, float* value_preds,
                                   float* rewards, float* masks, int32_t T, int64_t n_envs, int32_t do_reset,
                                   uint64_t seed, uint64_t step_offset, const mm_checkers_cfg* cfg, void* stream) {
    if (!params || !env_state || !obs || !share_obs || !rnn_states || !rnn_states_critic || !actions || !action_log_probs ||
        !value_preds || !rewards || !masks || !cfg || T <= 0 || n_envs <= 0)
        return MM_ERR_BAD_ARG;
    if (cfg->full_observable) return MM_ERR_UNSUPPORTED;
    if (!mm::aligned(rnn_states, 16) || !mm::aligned(rnn_states_critic, 16)) return MM_ERR_BAD_ARG;
    Args a;
    a.params = params; a.env_state = env_state; a.obs = obs; a.share_obs = share_obs;
    a.rnn_a = rnn_states; a.rnn_c = rnn_states_critic; a.actions = actions; a.log_probs = action_log_probs;
    a.values = value_preds; a.rewards = rewards; a.masks = masks; a.T = T; a.n_envs = n_envs; a.do_reset = do_reset;
    a.step_offset = step_offset; a.seed = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    a.cfg = *cfg; a.cfg.n_envs = (int32_t)n_envs; a.cfg.auto_reset = 1;
    const size_t smem = rollout_smem();
    if (smem > 227 * 1024) return MM_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(mappo_rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { mm::set_last_error(e); return MM_ERR_LAUNCH; }
    const long long tiles = (2 * n_envs + kRows - 1) / kRows;
    mappo_rollout_kernel<<<(unsigned)tiles, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(a);
    return mm::check_launch();
}
