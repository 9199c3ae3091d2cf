// Switch{2,4}-v0 lockstep env kernels for sm_100a.
//
// Semantics: ma-gym 0.0.14 `Switch.step/reset/get_agent_obs`;
// reference call sites qmix/main.py:66-71,103,115.  Lane = (env, agent); the agents of an env sit
// in adjacent lanes.  With at most 4 agents the sequential agent order is cheaper to replay
// redundantly in every lane of the env (the other agents' actions arrive by shuffle) than to
// hand positions from lane to lane.
#include "common.cuh"

namespace {

constexpr int kThreads = 128;
constexpr uint32_t kFreeCells = 0x63u | (0x7fu << 8) | (0x63u << 16);   // row 1 (corridor) + cols 0,1,5,6
// agents 0/1 live in row 0, agents 2/3 in row 2; even agents start left (col 1) and must reach
// col 6, odd agents start right (col 5) and must reach col 0
__host__ __device__ constexpr uint32_t init_cell(int a) { return ((a & 2) ? 16u : 0u) + ((a & 1) ? 5u : 1u); }
__host__ __device__ constexpr uint32_t goal_cell(int a) { return ((a & 2) ? 16u : 0u) + ((a & 1) ? 0u : 6u); }

__device__ __forceinline__ uint32_t init_cells(int n_agents) {
    uint32_t c = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a)
        if (a < n_agents) c |= init_cell(a) << (8 * a);
    return c;
}

template <int A>
__device__ __forceinline__ void write_obs(float* __restrict__ obs, int env, int agent, uint32_t cells,
                                          uint32_t step, const mm_switch_cfg& cfg) {
    const int d = 2 + (cfg.clock ? 1 : 0);
    // Python: step_count / max_steps in double, cast to float32 by the buffers
    const float clk = (float)((double)step / (double)cfg.max_steps);
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

template <int A>
__global__ void __launch_bounds__(kThreads)
switch_step_kernel(uint32_t* __restrict__ state, const int32_t* __restrict__ actions,
                   float* __restrict__ obs, float* __restrict__ rewards, uint8_t* __restrict__ dones,
                   int32_t* __restrict__ bad_action_flag, const __grid_constant__ mm_switch_cfg cfg) {
    const int gid = blockIdx.x * kThreads + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int env = gid / A, agent = gid % A;
    const bool valid = env < cfg.n_envs;
    const int n = cfg.n_envs;
    const int env_c = valid ? env : n - 1;
    uint32_t cells = state[env_c];
    uint32_t done = state[n + env_c];
    uint32_t step = state[2 * n + env_c] + 1;
    const int my_act = actions[(size_t)env_c * A + agent];
    const bool bad = (unsigned)my_act > 4u;
    float reward = cfg.step_cost;
    const int group = lane & ~(A - 1);
#pragma unroll
    for (int a = 0; a < A; ++a) {
        const int act = __shfl_sync(MM_FULL_MASK, my_act, group + a);
        if ((done >> a) & 1u) continue;                      // finished agents no longer move
        const uint32_t cell = (cells >> (8 * a)) & 0xffu;
        const int r = cell >> 3, c = cell & 7;
        const int nr = r + (act == 0) - (act == 2);
        const int nc = c + (act == 3) - (act == 1);
        uint32_t occupied = 0;
#pragma unroll
        for (int b = 0; b < A; ++b) occupied |= 1u << ((cells >> (8 * b)) & 0xffu);
        const uint32_t target = (uint32_t)(nr * 8 + nc);
        const bool vacant = ((unsigned)act < 4u) && ((unsigned)nr < 3u) && ((unsigned)nc < 7u) &&
                            (((kFreeCells & ~occupied) >> target) & 1u);
        const uint32_t now = vacant ? target : cell;
        cells = (cells & ~(0xffu << (8 * a))) | (now << (8 * a));
        if (now == goal_cell(a)) {
            done |= 1u << a;
            if (a == agent) reward = cfg.goal_reward;        // assignment, not +=, as in ma-gym
        }
    }
    if ((int)step >= cfg.max_steps) done = (1u << A) - 1u;
    if (!valid) return;
    if (agent == 0) {
        state[env] = cells;
        state[n + env] = done;
        state[2 * n + env] = step;
    }
    rewards[(size_t)env * A + agent] = reward;
    dones[(size_t)env * A + agent] = (uint8_t)((done >> agent) & 1u);
    if (bad && bad_action_flag != nullptr) *bad_action_flag = 1;
    write_obs<A>(obs, env, agent, cells, step, cfg);
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

extern "C" int mm_switch_step(uint32_t* state, const int32_t* actions, float* obs, float* rewards,
                              uint8_t* dones, int32_t* bad_action_flag, const mm_switch_cfg* cfg,
                              void* stream) {
    if (int rc = validate(state, cfg)) return rc;
    if (actions == nullptr || obs == nullptr || rewards == nullptr || dones == nullptr) return MM_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long threads = (long long)cfg->n_envs * cfg->n_agents;
    const dim3 grid((unsigned)((threads + kThreads - 1) / kThreads));
    if (cfg->n_agents == 2)
        switch_step_kernel<2><<<grid, kThreads, 0, st>>>(state, actions, obs, rewards, dones, bad_action_flag, *cfg);
    else
        switch_step_kernel<4><<<grid, kThreads, 0, st>>>(state, actions, obs, rewards, dones, bad_action_flag, *cfg);
    return mm::check_launch();
}
