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
#include "checkers_core.cuh"

namespace {

using namespace ck;
constexpr int kEnvsPerWarp = 32 / kAgents;         // 16
constexpr int kTileFloats = 32 * kObs;             // 1504 floats = one obs row per lane
// Few envs: 1-warp blocks spread the warps evenly over the 148 SMs (16 384 envs = 1024 warps = 6.9 per
// SM); many envs: 4-warp blocks.
constexpr int kSmallGridWarps = 148 * 16;

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
