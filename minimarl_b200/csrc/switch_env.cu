// Switch{2,4}-v0 lockstep env kernels for sm_100a.
//
// Semantics: ma-gym 0.0.14 `Switch.step/reset/get_agent_obs`;
// reference call sites qmix/main.py:66-71,103,115.  Lane = (env, agent); the agents of an env sit
// in adjacent lanes.  With at most 4 agents the sequential agent order is cheaper to replay
// redundantly in every lane of the env (the other agents' actions arrive by shuffle) than to
// hand positions from lane to lane.
#include "common.cuh"
#include "switch_core.cuh"

namespace {
using sw::init_cells;

constexpr int kThreads = 128;
template <int A>
__device__ __forceinline__ void write_obs(float* __restrict__ obs, int env, int agent, uint32_t cells,
                                          uint32_t step, const mm_switch_cfg& cfg) {
    const int d = 2 + (cfg.clock ? 1 : 0);
    // Python: step_count / max_steps in double, cast to float32 by the buffers
    const float clk = sw::clock_value(step, cfg);
    if (!cfg.full_observable) {
        float* o = obs + ((size_t)env * A + agent) * d;
        const uint32_t cell = (cells >> (8 * agent)) & 0xffu;
        o[0] = cfg.pos_row[cell >> 3];
        o[1] = cfg.pos_col[cell & 7];
        if (cfg.clock) o[2] = clk;
    } else {
        float* o = obs + ((size_t)env * A + agent) * (size_t)(A * d);
#pragma unroll
        for (int a = 0; a < A; ++a) {
            const uint32_t cell = (cells >> (8 * a)) & 0xffu;
            o[a * d + 0] = cfg.pos_row[cell >> 3];
            o[a * d + 1] = cfg.pos_col[cell & 7];
            if (cfg.clock) o[a * d + 2] = clk;
        }
    }
}

template <int A>
__global__ void __launch_bounds__(kThreads)
switch_reset_kernel(uint32_t* __restrict__ state, const uint8_t* __restrict__ reset_mask,
                    float* __restrict__ obs, const __grid_constant__ mm_switch_cfg cfg) {
    const int gid = blockIdx.x * kThreads + threadIdx.x;
    const int env = gid / A, agent = gid % A;
    if (env >= cfg.n_envs) return;
    const int n = cfg.n_envs;
    uint32_t cells, step;
    if (reset_mask == nullptr || reset_mask[env]) {
        cells = init_cells(A);
        step = 0;
        if (agent == 0) {
            state[env] = cells;
            state[n + env] = 0;
            state[2 * n + env] = 0;
        }
    } else {
        cells = state[env];
        step = state[2 * n + env];
    }
    if (obs != nullptr) write_obs<A>(obs, env, agent, cells, step, cfg);
}

// T lockstep steps in one launch (T = 1 is env.step): env state stays in registers between steps, the action of step
// t + 1 is fetched while step t is computed.  actions == nullptr: uniform actions from Philox4x32-10 keyed by (seed), counter
// (step_offset + t, row), like checkers_rollout_kernel.
template <int A>
__global__ void __launch_bounds__(kThreads)
switch_rollout_kernel(uint32_t* __restrict__ state, const int32_t* __restrict__ actions, int32_t* __restrict__ actions_out,
                      float* __restrict__ obs, float* __restrict__ rewards, uint8_t* __restrict__ dones,
                      int32_t* __restrict__ bad_action_flag, int T, uint2 seed, unsigned long long step_offset,
                      const __grid_constant__ mm_switch_cfg cfg) {
    const int gid = blockIdx.x * kThreads + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int env = gid / A, agent = gid % A;
    const bool valid = env < cfg.n_envs;
    const int n = cfg.n_envs;
    const int env_c = valid ? env : n - 1;
    const size_t rows = (size_t)n * A, row = (size_t)env_c * A + agent;
    const size_t obs_step = rows * (size_t)((2 + (cfg.clock ? 1 : 0)) * (cfg.full_observable ? A : 1));
    uint32_t cells = state[env_c];
    uint32_t done = state[n + env_c];
    uint32_t step = state[2 * n + env_c];
    bool any_bad = false;
    int act_next = actions != nullptr ? actions[row] : 0;
    for (int t = 0; t < T; ++t) {
        int my_act;
        if (actions != nullptr) {
            my_act = act_next;
            if (t + 1 < T) act_next = actions[(size_t)(t + 1) * rows + row];
        } else {
            const unsigned long long ctr = step_offset + (unsigned long long)t;
            const uint4 rnd = mm::philox4x32_10(
                make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)row, (uint32_t)((unsigned long long)row >> 32)), seed);
            my_act = (int)__umulhi(rnd.x, 5u);
            if (actions_out != nullptr && valid) actions_out[(size_t)t * rows + row] = my_act;
        }
        any_bad |= (unsigned)my_act > 4u;
        float reward;
        sw::transition<A>(cells, done, step, my_act, agent, lane, cfg, reward);
        if (valid) {
            if (rewards != nullptr) rewards[(size_t)t * rows + row] = reward;
            if (dones != nullptr) dones[(size_t)t * rows + row] = (uint8_t)((done >> agent) & 1u);
            if (obs != nullptr) write_obs<A>(obs + (size_t)t * obs_step, env, agent, cells, step, cfg);
        }
    }
    if (!valid) return;
    if (agent == 0) {
        state[env] = cells;
        state[n + env] = done;
        state[2 * n + env] = step;
    }
    if (any_bad && bad_action_flag != nullptr) *bad_action_flag = 1;
}

int validate(const uint32_t* state, const mm_switch_cfg* cfg) {
    if (state == nullptr || cfg == nullptr) return MM_ERR_BAD_ARG;
    if (cfg->n_envs <= 0 || cfg->max_steps <= 0) return MM_ERR_BAD_ARG;
    if (cfg->n_agents != 2 && cfg->n_agents != 4) return MM_ERR_UNSUPPORTED;
    return MM_OK;
}

}  // namespace

extern "C" int mm_switch_reset(uint32_t* state, const uint8_t* reset_mask, float* obs,
                               const mm_switch_cfg* cfg, void* stream) {
    if (int rc = validate(state, cfg)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long threads = (long long)cfg->n_envs * cfg->n_agents;
    const dim3 grid((unsigned)((threads + kThreads - 1) / kThreads));
    if (cfg->n_agents == 2) switch_reset_kernel<2><<<grid, kThreads, 0, st>>>(state, reset_mask, obs, *cfg);
    else                    switch_reset_kernel<4><<<grid, kThreads, 0, st>>>(state, reset_mask, obs, *cfg);
    return mm::check_launch();
}

extern "C" int mm_switch_rollout(uint32_t* state, const int32_t* actions, int32_t* actions_out, float* obs, float* rewards,
                                 uint8_t* dones, int32_t* bad_action_flag, int32_t T, uint64_t seed, uint64_t step_offset,
                                 const mm_switch_cfg* cfg, void* stream) {
    if (int rc = validate(state, cfg)) return rc;
    if (T <= 0) return MM_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long threads = (long long)cfg->n_envs * cfg->n_agents;
    const dim3 grid((unsigned)((threads + kThreads - 1) / kThreads));
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    if (cfg->n_agents == 2)
        switch_rollout_kernel<2><<<grid, kThreads, 0, st>>>(state, actions, actions_out, obs, rewards, dones, bad_action_flag, T,
                                                            key, step_offset, *cfg);
    else
        switch_rollout_kernel<4><<<grid, kThreads, 0, st>>>(state, actions, actions_out, obs, rewards, dones, bad_action_flag, T,
                                                            key, step_offset, *cfg);
    return mm::check_launch();
}

extern "C" int mm_switch_step(uint32_t* state, const int32_t* actions, float* obs, float* rewards,
                              uint8_t* dones, int32_t* bad_action_flag, const mm_switch_cfg* cfg,
                              void* stream) {
    if (actions == nullptr || obs == nullptr || rewards == nullptr || dones == nullptr) return MM_ERR_BAD_ARG;
    return mm_switch_rollout(state, actions, nullptr, obs, rewards, dones, bad_action_flag, 1, 0, 0, cfg, stream);
}
