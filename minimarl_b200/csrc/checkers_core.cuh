// Checkers-v0 rules as device functions shared by the env kernels (checkers_env.cu) and the fused
// policy + env rollout kernel (rollout_tc.cu).  Semantics: ma-gym 0.0.14 `Checkers`.
// Lane mapping: adjacent lanes (lane & 1 = agent) hold the two agents of one env; both lanes carry an
// identical copy of the env registers, kept in sync with shuffles.
#pragma once
#include "common.cuh"

namespace ck {

constexpr int kAgents = MM_CHECKERS_AGENTS;
constexpr int kObs = MM_CHECKERS_OBS_DIM;          // 47

constexpr uint32_t kFoodCells = 0x3f3f3fu;                            // cols 0..5 of rows 0..2
constexpr uint32_t kAppleCells = 0x15u | (0x2au << 8) | (0x15u << 16);  // (row + col) even
constexpr uint32_t kLemonCells = kFoodCells ^ kAppleCells;
constexpr uint32_t kInitCells = (0u * 8 + 6) | ((2u * 8 + 6) << 8);   // A1 (0,6), A2 (2,6)
constexpr uint32_t kDoneLatch = 0x80000000u;

struct EnvRegs {
    uint32_t food;   // bitmap | latch
    uint32_t cells;  // agent cells, 8 bits each
    uint32_t step;
};

// 3 bits (cols c-1..c+1) of row `row` (-1..3) of a 24-bit cell map; cells outside the grid read 0.
__device__ __forceinline__ uint32_t neigh3(uint32_t map24, int row, int c) {
    uint32_t rowbits = __funnelshift_rc(map24 << 8, 0u, 8 * (row + 1)) & 0xffu;
    return ((rowbits << 1) >> c) & 7u;
}
// bits b0,b1,b2 -> bit positions 0,5,10 (one 5-wide one-hot slot per cell)
__device__ __forceinline__ uint32_t spread5(uint32_t x) { return (x * 0x111u) & 0x421u; }

// 45-bit mask of the 3x3x5 neighbourhood around (r, c): bit (5*cell + channel), channels
// lemon, apple, A1, A2, wall (never set: Checkers has no walls).
__device__ __forceinline__ uint64_t neighbourhood_mask(uint32_t food, uint32_t cells, int r, int c) {
    const uint32_t lemons = food & kLemonCells;
    const uint32_t apples = food & kAppleCells;
    const uint32_t a1 = 1u << (cells & 0xffu);
    const uint32_t a2 = 1u << ((cells >> 8) & 0xffu);
    uint32_t rows[3];
#pragma unroll
    for (int dr = 0; dr < 3; ++dr) {
        const int row = r + dr - 1;
        rows[dr] = spread5(neigh3(lemons, row, c)) | (spread5(neigh3(apples, row, c)) << 1) |
                   (spread5(neigh3(a1, row, c)) << 2) | (spread5(neigh3(a2, row, c)) << 3);
    }
    return (uint64_t)rows[0] | ((uint64_t)rows[1] << 15) | ((uint64_t)rows[2] << 30);
}

// One env.step for the (env, agent) of this lane.  All 32 lanes must call it.
// reward_kind: 0 none, 1 lemon, 2 apple (index into cfg.reward[agent]), 3 = reward forced to 0 (episode already over)
__device__ __forceinline__ void transition(EnvRegs& s, int act, int agent, int lane,
                                           const mm_checkers_cfg& cfg, float& reward, uint32_t& done,
                                           bool& bad_action, int* reward_kind = nullptr) {
    const uint32_t latch = s.food & kDoneLatch;
    const uint32_t mine = (s.cells >> (8 * agent)) & 0xffu;
    const uint32_t other = (s.cells >> (8 * (agent ^ 1))) & 0xffu;
    const int r = mine >> 3, c = mine & 7;
    const int nr = r + (act == 0) - (act == 2);
    const int nc = c + (act == 3) - (act == 1);
    bad_action = (unsigned)act > 4u;
    const bool in_grid = ((unsigned)act < 4u) && ((unsigned)nr < 3u) && ((unsigned)nc < 8u);
    const uint32_t target = (uint32_t)(nr * 8 + nc);
    // agent 0 is blocked only by the cell agent 1 occupied before this step
    const uint32_t new_if_first = (in_grid && target != other) ? target : mine;
    const uint32_t first_new = __shfl_sync(MM_FULL_MASK, new_if_first, lane & ~1);
    // agent 1 is blocked by the cell agent 0 occupies now
    const uint32_t mine_new = (agent == 0) ? new_if_first : ((in_grid && target != first_new) ? target : mine);
    const uint32_t other_new = __shfl_xor_sync(MM_FULL_MASK, mine_new, 1);

    const uint32_t bit = 1u << mine_new;
    const uint32_t eaten = (mine_new != mine) ? (s.food & bit & kFoodCells) : 0u;
    const uint32_t eaten_other = __shfl_xor_sync(MM_FULL_MASK, eaten, 1);
    const int kind = eaten ? ((eaten & kAppleCells) ? 2 : 1) : 0;

    const uint32_t food_new = s.food & kFoodCells & ~(eaten | eaten_other);
    s.cells = (agent == 0) ? (mine_new | (other_new << 8)) : (other_new | (mine_new << 8));
    s.step += 1;
    const bool done_now = ((int)s.step >= cfg.max_steps) || ((food_new & kAppleCells) == 0u);
    // rewards are forced to 0 once the episode had already finished before this step
    reward = latch ? 0.0f : cfg.reward[agent][kind];
    if (reward_kind) *reward_kind = latch ? 3 : kind;
    done = (latch || done_now) ? 1u : 0u;
    s.food = food_new | (done ? kDoneLatch : 0u);
}

__device__ __forceinline__ void reset_regs(EnvRegs& s) {
    s.food = kFoodCells;
    s.cells = kInitCells;
    s.step = 0;
}

// Expand this lane's agent observation (2 position floats + 45 one-hot floats) into the warp tile.
__device__ __forceinline__ void fill_tile(float* tile, int lane, int agent, const EnvRegs& s,
                                          const mm_checkers_cfg& cfg) {
    const uint32_t mine = (s.cells >> (8 * agent)) & 0xffu;
    const int r = mine >> 3, c = mine & 7;
    const uint64_t m = neighbourhood_mask(s.food, s.cells, r, c);
    const uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
    float* row = tile + lane * kObs;   // stride 47 words: odd, so the 32 lanes hit 32 different banks
    row[0] = cfg.pos_row[r];
    row[1] = cfg.pos_col[c];
#pragma unroll
    for (int k = 0; k < 32; ++k) row[2 + k] = (lo & (1u << k)) ? 1.0f : 0.0f;
#pragma unroll
    for (int k = 32; k < 45; ++k) row[2 + k] = (hi & (1u << (k - 32))) ? 1.0f : 0.0f;
}


}  // namespace ck
