// Switch{2,4}-v0 rules as a device function shared by the env kernels (switch_env.cu) and the fused top-level
// rollout kernel (toplevel_qnet.cu).  Semantics: ma-gym 0.0.14 `Switch.step`;
// reference call sites qmix/main.py:66-71,103,115.  Lane mapping: the A agents of an env sit in adjacent lanes
// and every lane of the env replays the sequential agent order redundantly (the other agents' actions arrive by
// shuffle), so all lanes of an env hold identical (cells, done, step) registers.
#pragma once
#include "common.cuh"

namespace sw {

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

// One env.step for the (env, agent) of this lane; all 32 lanes must call it.  `done` holds one bit per agent.
template <int A>
__device__ __forceinline__ void transition(uint32_t& cells, uint32_t& done, uint32_t& step, int my_act, int agent, int lane,
                                           const mm_switch_cfg& cfg, float& reward) {
    const int group = lane & ~(A - 1);
    step += 1;
    reward = cfg.step_cost;
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
}

// Python: step_count / max_steps in double, cast to float32 by the buffers
__device__ __forceinline__ float clock_value(uint32_t step, const mm_switch_cfg& cfg) {
    return (float)((double)step / (double)cfg.max_steps);
}

}  // namespace sw
