// Checkers-v0 lockstep env kernels (reset / step / T-step rollout) for sm_100a.
//
// Semantics: ma-gym 0.0.14 `Checkers.step/reset/get_agent_obs`;
// reference call sites vdn/main.py:83,93, mappo/runner/shared/magym_runner.py:52-57,
// offpolicy/runner/shared/magym_runner.py:94-97.
//
// Mapping: lane = (env, agent).  A warp owns 16 consecutive envs, the two agents of an env sit in
// adjacent lanes.  The sequential agent order of the reference (agent 1 sees the grid agent 0 has
// already changed) is resolved with two shuffles: agent 0's new cell goes to agent 1's lane before
// agent 1 decides, and the "food eaten" bits are exchanged so both lanes keep the same grid.
// The 3x3x5 one-hot neighbourhood of each agent is built as a 45-bit mask with shifts, expanded to
// floats into a per-warp shared tile laid out exactly like the output rows, and the tile is then
// streamed to HBM with 128-bit coalesced stores (the obs write is ~92 % of the bytes of a step).
#include "common.cuh"

namespace {

constexpr int kAgents = MM_CHECKERS_AGENTS;
constexpr int kObs = MM_CHECKERS_OBS_DIM;          // 47
constexpr int kEnvsPerWarp = 32 / kAgents;         // 16
constexpr int kTileFloats = 32 * kObs;             // 1504 floats = one obs row per lane
// Few envs: 1-warp blocks spread the warps evenly over the 148 SMs (16 384 envs = 1024 warps = 6.9 per
// SM); many envs: 4-warp blocks.
constexpr int kSmallGridWarps = 148 * 16;

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
__device__ __forceinline__ void transition(EnvRegs& s, int act, int agent, int lane,
                                           const mm_checkers_cfg& cfg, float& reward, uint32_t& done,
                                           bool& bad_action) {
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

// Stream the warp tile (n_valid envs x 94 floats, contiguous) to obs.
//   FULL == false: out row of env e = tile[e*94 .. e*94+93]                       ([N,2,47])
//   FULL == true : out row of env e = that block twice (every agent sees both)     ([N,2,94])
// VEC4 requires `out` 16-byte aligned; the 8-byte path covers odd n_envs / odd slot offsets.
template <bool FULL, bool VEC4>
__device__ __forceinline__ void write_tile(const float* tile, float* out, int lane, int n_valid) {
    constexpr int kEnvFloats = kAgents * kObs;  // 94
    if (!FULL) {
        const int total = n_valid * kEnvFloats;  // even
        if (VEC4) {
            const int n4 = total >> 2;
            const float4* src = reinterpret_cast<const float4*>(tile);
            float4* dst = reinterpret_cast<float4*>(out);
#pragma unroll 4
            for (int q = lane; q < n4; q += 32) mm::st_stream(dst + q, src[q]);
            if ((total & 3) && lane == 0) {
                const float2* s2 = reinterpret_cast<const float2*>(tile);
                mm::st_stream(reinterpret_cast<float2*>(out) + (total >> 1) - 1, s2[(total >> 1) - 1]);
            }
        } else {
            const int n2 = total >> 1;
            const float2* src = reinterpret_cast<const float2*>(tile);
            float2* dst = reinterpret_cast<float2*>(out);
#pragma unroll 4
            for (int q = lane; q < n2; q += 32) mm::st_stream(dst + q, src[q]);
        }
    } else {
        // 188 floats per env = 47 float4; element j of the row reads tile[e*94 + j % 94]
        const float2* src = reinterpret_cast<const float2*>(tile);
        if (VEC4) {
            const int n4 = n_valid * 47;
            float4* dst = reinterpret_cast<float4*>(out);
#pragma unroll 4
            for (int q = lane; q < n4; q += 32) {
                const int e = q / 47, w = q - e * 47;     // float4 w of env e
                int j0 = 2 * w, j1 = 2 * w + 1;           // float2 indices within the 94-float row
                if (j0 >= 47) j0 -= 47;
                if (j1 >= 47) j1 -= 47;
                const float2 a = src[e * 47 + j0], b = src[e * 47 + j1];
                mm::st_stream(dst + q, make_float4(a.x, a.y, b.x, b.y));
            }
        } else {
            const int n2 = n_valid * 94;
            float2* dst = reinterpret_cast<float2*>(out);
#pragma unroll 4
            for (int q = lane; q < n2; q += 32) {
                const int e = q / 94;
                int j = q - e * 94;
                if (j >= 47) j -= 47;
                mm::st_stream(dst + q, src[e * 47 + j]);
            }
        }
    }
}

// The caller alternates between two tiles, so the only barrier needed per step is the one between
// fill and read-out: a lane can start filling tile B while slower lanes still read tile A, and by the
// time anyone refills tile A every lane has passed the barrier of the step in between.
template <bool FULL, bool VEC4>
__device__ __forceinline__ void emit_obs(float* tile, float* obs, int warp_env0, int lane, int agent,
                                         const EnvRegs& s, const mm_checkers_cfg& cfg) {
    fill_tile(tile, lane, agent, s, cfg);
    __syncwarp();
    const int n_valid = min(kEnvsPerWarp, cfg.n_envs - warp_env0);
    const size_t env_floats = (size_t)kAgents * kObs * (FULL ? 2 : 1);
    write_tile<FULL, VEC4>(tile, obs + (size_t)warp_env0 * env_floats, lane, n_valid);
}

__device__ __forceinline__ EnvRegs load_state(const uint32_t* state, int env, int n) {
    EnvRegs s;
    s.food = state[env];
    s.cells = state[n + env];
    s.step = state[2 * n + env];
    return s;
}
__device__ __forceinline__ void store_state(uint32_t* state, int env, int n, const EnvRegs& s) {
    state[env] = s.food;
    state[n + env] = s.cells;
    state[2 * n + env] = s.step;
}

template <bool FULL, bool VEC4, int WARPS>
__global__ void __launch_bounds__(32 * WARPS)
checkers_reset_kernel(uint32_t* __restrict__ state, const uint8_t* __restrict__ reset_mask,
                      float* __restrict__ obs, const __grid_constant__ mm_checkers_cfg cfg) {
    __shared__ __align__(16) float tiles[WARPS][kTileFloats];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warp_env0 = (blockIdx.x * WARPS + warp) * kEnvsPerWarp;
    if (warp_env0 >= cfg.n_envs) return;
    const int agent = lane & 1;
    const int env = warp_env0 + (lane >> 1);
    const bool valid = env < cfg.n_envs;
    const int env_c = valid ? env : cfg.n_envs - 1;
    EnvRegs s;
    const bool do_reset = reset_mask == nullptr || reset_mask[env_c] != 0;
    if (do_reset) {
        reset_regs(s);
        if (valid && agent == 0) store_state(state, env, cfg.n_envs, s);
    } else {
        s = load_state(state, env_c, cfg.n_envs);
    }
    if (obs != nullptr) emit_obs<FULL, VEC4>(tiles[warp], obs, warp_env0, lane, agent, s, cfg);
}

// T == 1 with actions given is env.step; T > 1 keeps the state in registers across steps.
template <bool FULL, bool VEC4, int WARPS>
__global__ void __launch_bounds__(32 * WARPS)
checkers_rollout_kernel(uint32_t* __restrict__ state, const int32_t* __restrict__ actions,
                        int32_t* __restrict__ actions_out, float* __restrict__ obs,
                        float* __restrict__ rewards, uint8_t* __restrict__ dones,
                        int32_t* __restrict__ bad_action_flag, int T, uint2 seed, uint64_t step_offset,
                        const __grid_constant__ mm_checkers_cfg cfg) {
    __shared__ __align__(16) float tiles[WARPS][2][kTileFloats];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warp_env0 = (blockIdx.x * WARPS + warp) * kEnvsPerWarp;
    if (warp_env0 >= cfg.n_envs) return;
    const int agent = lane & 1;
    const int env = warp_env0 + (lane >> 1);
    const bool valid = env < cfg.n_envs;
    const int env_c = valid ? env : cfg.n_envs - 1;
    const size_t row = (size_t)env_c * kAgents + agent;          // (env, agent) row
    const size_t rows = (size_t)cfg.n_envs * kAgents;
    const size_t obs_step_floats = rows * kObs * (FULL ? 2 : 1);

    EnvRegs s = load_state(state, env_c, cfg.n_envs);
    bool any_bad = false;
    // the action of step t+1 is fetched while step t is computed: the load never sits on the critical path
    int act_next = (actions != nullptr) ? actions[row] : 0;
    for (int t = 0; t < T; ++t) {
        int act;
        if (actions != nullptr) {
            act = act_next;
            if (t + 1 < T) act_next = actions[(size_t)(t + 1) * rows + row];
        } else {
            const uint64_t ctr = step_offset + (uint64_t)t;
            const uint4 rnd = mm::philox4x32_10(
                make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)row, (uint32_t)(row >> 32)), seed);
            act = (int)__umulhi(rnd.x, 5u);
            if (actions_out != nullptr && valid) actions_out[(size_t)t * rows + row] = act;
        }
        float reward;
        uint32_t done;
        bool bad;
        transition(s, act, agent, lane, cfg, reward, done, bad);
        any_bad |= bad;
        if (valid) {
            if (rewards != nullptr) rewards[(size_t)t * rows + row] = reward;
            if (dones != nullptr) dones[(size_t)t * rows + row] = (uint8_t)done;
        }
        if (cfg.auto_reset && done) reset_regs(s);
        if (obs != nullptr)
            emit_obs<FULL, VEC4>(tiles[warp][t & 1], obs + (size_t)t * obs_step_floats, warp_env0, lane, agent, s,
                                 cfg);
    }
    if (valid && agent == 0) store_state(state, env, cfg.n_envs, s);
    if (bad_action_flag != nullptr && any_bad && valid) *bad_action_flag = 1;
}

int validate(const uint32_t* state, const mm_checkers_cfg* cfg) {
    if (state == nullptr || cfg == nullptr) return MM_ERR_BAD_ARG;
    if (cfg->n_envs <= 0 || cfg->max_steps <= 0) return MM_ERR_BAD_ARG;
    return MM_OK;
}

inline int warps_for(int n_envs) { return (n_envs + kEnvsPerWarp - 1) / kEnvsPerWarp; }

}  // namespace

extern "C" int mm_checkers_reset(uint32_t* state, const uint8_t* reset_mask, float* obs,
                                 const mm_checkers_cfg* cfg, void* stream) {
    if (int rc = validate(state, cfg)) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool v4 = obs == nullptr || mm::aligned(obs, 16);
    const int warps = warps_for(cfg->n_envs);
#define MM_LAUNCH(F, V)                                                                                   \
    do {                                                                                                  \
        if (warps <= kSmallGridWarps)                                                                     \
            checkers_reset_kernel<F, V, 1><<<warps, 32, 0, st>>>(state, reset_mask, obs, *cfg);           \
        else                                                                                              \
            checkers_reset_kernel<F, V, 4><<<(warps + 3) / 4, 128, 0, st>>>(state, reset_mask, obs, *cfg); \
    } while (0)
    if (cfg->full_observable) { if (v4) MM_LAUNCH(true, true); else MM_LAUNCH(true, false); }
    else                      { if (v4) MM_LAUNCH(false, true); else MM_LAUNCH(false, false); }
#undef MM_LAUNCH
    return mm::check_launch();
}

extern "C" int mm_checkers_rollout(uint32_t* state, const int32_t* actions, int32_t* actions_out,
                                   float* obs, float* rewards, uint8_t* dones, int32_t* bad_action_flag,
                                   int32_t T, uint64_t seed, uint64_t step_offset,
                                   const mm_checkers_cfg* cfg, void* stream) {
    if (int rc = validate(state, cfg)) return rc;
    if (T <= 0) return MM_ERR_BAD_ARG;
    if (obs != nullptr && !mm::aligned(obs, 8)) return MM_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // every per-step obs slice must keep the alignment of the vector path
    const size_t step_bytes = (size_t)cfg->n_envs * kAgents * kObs * (cfg->full_observable ? 2 : 1) * 4;
    const bool v4 = obs == nullptr || (mm::aligned(obs, 16) && (T == 1 || step_bytes % 16 == 0));
    const int warps = warps_for(cfg->n_envs);
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
#define MM_LAUNCH(F, V)                                                                                       \
    do {                                                                                                      \
        if (warps <= kSmallGridWarps)                                                                         \
            checkers_rollout_kernel<F, V, 1><<<warps, 32, 0, st>>>(state, actions, actions_out, obs, rewards,  \
                                                                   dones, bad_action_flag, T, key, step_offset, *cfg); \
        else                                                                                                  \
            checkers_rollout_kernel<F, V, 4><<<(warps + 3) / 4, 128, 0, st>>>(                                 \
                state, actions, actions_out, obs, rewards, dones, bad_action_flag, T, key, step_offset, *cfg); \
    } while (0)
    if (cfg->full_observable) { if (v4) MM_LAUNCH(true, true); else MM_LAUNCH(true, false); }
    else                      { if (v4) MM_LAUNCH(false, true); else MM_LAUNCH(false, false); }
#undef MM_LAUNCH
    return mm::check_launch();
}

extern "C" int mm_checkers_step(uint32_t* state, const int32_t* actions, float* obs, float* rewards,
                                uint8_t* dones, int32_t* bad_action_flag, const mm_checkers_cfg* cfg,
                                void* stream) {
    if (actions == nullptr) return MM_ERR_BAD_ARG;
    return mm_checkers_rollout(state, actions, nullptr, obs, rewards, dones, bad_action_flag, 1, 0, 0, cfg,
                               stream);
}
