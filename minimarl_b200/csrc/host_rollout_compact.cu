// Host-buffer rollout with a compact device-to-host transport (bench "e2e").
//
// mm_checkers_rollout_host moves the env outputs as they are: 47 float32 per agent-step, 386 B per env-step, so a caller
// holding host arrays is bound by PCIe (54 GB/s = 15 ms per 16 384 x 128 rollout) and, with eight ranks, by the host's DRAM
// write bandwidth.  But an observation is 45 one-hot bits and two grid coordinates, a reward one of four values, a done one
// bit: here the kernel emits ONE 64-bit word per agent-step
//     bits 0..44 neighbourhood mask (3x3 cells x 5 channels) | 48..50 row | 52..54 col | 56..57 reward kind | 60 done
// (8 B instead of 193 B over PCIe), and host threads expand the words into the SAME float32 / uint8 arrays -- bit-identical
// to the direct path: the float values come from the same LUTs -- with non-temporal stores, chunk by chunk while the GPU
// and the link work on the next chunk.  The expansion writes 386 B per env-step to host memory and is what bounds this path.
#include <atomic>
#include <thread>
#include <vector>

#include <emmintrin.h>
#include <sched.h>

#include "checkers_core.cuh"
#include "common.cuh"

namespace {
using namespace ck;

constexpr int kEnvsPerWarp = 16;

template <int WARPS>
__global__ void __launch_bounds__(32 * WARPS)
checkers_rollout_codes_kernel(uint32_t* __restrict__ state, const int32_t* __restrict__ actions,
                              unsigned long long* __restrict__ codes, int32_t* __restrict__ bad_action_flag, int T,
                              const __grid_constant__ mm_checkers_cfg cfg) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warp_env0 = (blockIdx.x * WARPS + warp) * kEnvsPerWarp;
    if (warp_env0 >= cfg.n_envs) return;
    const int agent = lane & 1;
    const int env = warp_env0 + (lane >> 1);
    const bool valid = env < cfg.n_envs;
    const int env_c = valid ? env : cfg.n_envs - 1;
    const size_t row = (size_t)env_c * kAgents + agent, rows = (size_t)cfg.n_envs * kAgents;
    EnvRegs s;
    s.food = state[env_c]; s.cells = state[cfg.n_envs + env_c]; s.step = state[2 * cfg.n_envs + env_c];
    bool any_bad = false;
    int act_next = actions[row];
    for (int t = 0; t < T; ++t) {
        const int act = act_next;
        if (t + 1 < T) act_next = actions[(size_t)(t + 1) * rows + row];
        float reward;
        uint32_t done;
        bool bad;
        int kind;
        transition(s, act, agent, lane, cfg, reward, done, bad, &kind);
        any_bad |= bad;
        if (cfg.auto_reset && done) reset_regs(s);
        const uint32_t mine = (s.cells >> (8 * agent)) & 0xffu;
        const int r = mine >> 3, c = mine & 7;
        const unsigned long long code = neighbourhood_mask(s.food, s.cells, r, c) | ((unsigned long long)r << 48) |
                                        ((unsigned long long)c << 52) | ((unsigned long long)kind << 56) |
                                        ((unsigned long long)done << 60);
        if (valid) codes[(size_t)t * rows + row] = code;
    }
    if (valid && agent == 0) {
        state[env] = s.food; state[cfg.n_envs + env] = s.cells; state[2 * cfg.n_envs + env] = s.step;
    }
    if (bad_action_flag != nullptr && any_bad && valid) *bad_action_flag = 1;
}

struct Lut {
    alignas(16) float byte8[256][8];      // the eight floats of one mask byte
    float pos_row[3], pos_col[8], reward[2][4];
};

void build_lut(Lut& lut, const mm_checkers_cfg& cfg) {
    for (int b = 0; b < 256; ++b)
        for (int k = 0; k < 8; ++k) lut.byte8[b][k] = (b >> k) & 1 ? 1.0f : 0.0f;
    for (int r = 0; r < 3; ++r) lut.pos_row[r] = cfg.pos_row[r];
    for (int c = 0; c < 8; ++c) lut.pos_col[c] = cfg.pos_col[c];
    for (int a = 0; a < 2; ++a) {
        for (int k = 0; k < 3; ++k) lut.reward[a][k] = cfg.reward[a][k];
        lut.reward[a][3] = 0.0f;
    }
}

// the 47 floats of one agent observation
inline void expand47(float* dst, unsigned long long code, const Lut& lut) {
    dst[0] = lut.pos_row[(code >> 48) & 7];
    dst[1] = lut.pos_col[(code >> 52) & 7];
    for (int b = 0; b < 5; ++b) {
        const float* src = lut.byte8[(code >> (8 * b)) & 0xff];
        for (int k = 0; k < 8; ++k) dst[2 + 8 * b + k] = src[k];
    }
    for (int k = 40; k < 45; ++k) dst[2 + k] = (code >> k) & 1 ? 1.0f : 0.0f;
}

inline void stream_out(float* dst, const float* src, size_t n_floats) {          // dst 16-byte aligned, n_floats % 4 == 0
    for (size_t i = 0; i < n_floats; i += 4) _mm_stream_ps(dst + i, _mm_loadu_ps(src + i));
}

// rows [r0, r1) of one chunk (row = (step, env, agent) flattened; r0, r1 multiples of 4)
void expand_rows(const unsigned long long* codes, size_t r0, size_t r1, float* obs, float* rewards, uint8_t* dones, bool full,
                 const Lut& lut) {
    alignas(16) float buf[4 * 94];
    for (size_t r = r0; r < r1; r += 4) {
        const size_t n = r1 - r < 4 ? r1 - r : 4;                   // always 4 except for a ragged tail
        if (obs) {
            if (!full) {
                for (size_t i = 0; i < n; ++i) expand47(buf + 47 * i, codes[r + i], lut);
                if (n == 4 && (reinterpret_cast<uintptr_t>(obs + r * 47) & 15) == 0) stream_out(obs + r * 47, buf, 4 * 47);
                else for (size_t i = 0; i < n * 47; ++i) obs[r * 47 + i] = buf[i];
            } else {
                // rows come in (agent 0, agent 1) pairs of one env; every agent sees concat(obs_0, obs_1)
                for (size_t i = 0; i + 1 < n; i += 2) {
                    expand47(buf, codes[r + i], lut);
                    expand47(buf + 47, codes[r + i + 1], lut);
                    for (int k = 0; k < 94; ++k) buf[94 + k] = buf[k];
                    float* d = obs + (r + i) * 94;
                    if ((reinterpret_cast<uintptr_t>(d) & 15) == 0) stream_out(d, buf, 188);
                    else for (int k = 0; k < 188; ++k) d[k] = buf[k];
                }
            }
        }
        for (size_t i = 0; i < n; ++i) {
            const unsigned long long c = codes[r + i];
            if (rewards) rewards[r + i] = lut.reward[(r + i) & 1][(c >> 56) & 3];
            if (dones) dones[r + i] = (uint8_t)((c >> 60) & 1);
        }
    }
}

#define MM_CUDA(call)                      \
    do {                                   \
        cudaError_t e_ = (call);           \
        if (e_ != cudaSuccess) {           \
            mm::set_last_error(e_);        \
            return MM_ERR_LAUNCH;          \
        }                                  \
    } while (0)

int allowed_cpus() {
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
        const int n = CPU_COUNT(&set);
        if (n > 0) return n;
    }
    const unsigned h = std::thread::hardware_concurrency();
    return h ? (int)h : 1;
}

}  // namespace

// codes_host: page-locked uint64 [T][N][2]; codes_stage: device uint64 [2][chunk][N][2]; actions_stage: device int32 [2][chunk][N][2].
// n_threads <= 0: one expansion thread per CPU this process may run on (at most 32).  Returns after every output landed.
extern "C" int mm_checkers_rollout_host_compact(uint32_t* state, const int32_t* actions_host, float* obs_host,
                                                float* rewards_host, uint8_t* dones_host, uint64_t* codes_host,
                                                int32_t* actions_stage, uint64_t* codes_stage, int32_t* bad_action_flag,
                                                int32_t T, int32_t chunk, int32_t n_threads, const mm_checkers_cfg* cfg,
                                                void* stream, void* copy_stream) {
    if (!state || !cfg || !actions_host || !actions_stage || !codes_host || !codes_stage) return MM_ERR_BAD_ARG;
    if (T <= 0 || chunk <= 0 || cfg->n_envs <= 0 || copy_stream == stream) return MM_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream), cp = static_cast<cudaStream_t>(copy_stream);
    const size_t rows = (size_t)cfg->n_envs * MM_CHECKERS_AGENTS;
    const int n_chunks = (T + chunk - 1) / chunk;
    std::vector<cudaEvent_t> drained(n_chunks, nullptr);
    cudaEvent_t ready[2] = {nullptr, nullptr};
    auto cleanup = [&]() {
        for (auto e : drained) if (e) cudaEventDestroy(e);
        for (auto e : ready) if (e) cudaEventDestroy(e);
    };
    for (auto& e : drained)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cleanup(); return MM_ERR_LAUNCH; }
    for (auto& e : ready)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cleanup(); return MM_ERR_LAUNCH; }

    // ---- device side: everything is enqueued up front (the words are 8 B per agent-step: the link is nearly idle) ----------
    const int warps = (cfg->n_envs + kEnvsPerWarp - 1) / kEnvsPerWarp;
    int rc = MM_OK;
    for (int k = 0; k < n_chunks && rc == MM_OK; ++k) {
        const int b = k & 1, t0 = k * chunk, tc = (T - t0 < chunk) ? T - t0 : chunk;
        const size_t off = (size_t)t0 * rows, stage = (size_t)b * chunk * rows, n = (size_t)tc * rows;
        cudaError_t e = cudaSuccess;
        if (k >= 2) e = cudaStreamWaitEvent(st, drained[k - 2], 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(actions_stage + stage, actions_host + off, n * sizeof(int32_t), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) {
            if (warps <= 148 * 16)
                checkers_rollout_codes_kernel<1><<<warps, 32, 0, st>>>(state, actions_stage + stage,
                    reinterpret_cast<unsigned long long*>(codes_stage + stage), bad_action_flag, tc, *cfg);
            else
                checkers_rollout_codes_kernel<4><<<(warps + 3) / 4, 128, 0, st>>>(state, actions_stage + stage,
                    reinterpret_cast<unsigned long long*>(codes_stage + stage), bad_action_flag, tc, *cfg);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaEventRecord(ready[b], st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(cp, ready[b], 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(codes_host + off, codes_stage + stage, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, cp);
        if (e == cudaSuccess) e = cudaEventRecord(drained[k], cp);
        if (e != cudaSuccess) { mm::set_last_error(e); rc = MM_ERR_LAUNCH; }
    }
    if (rc != MM_OK) { cudaStreamSynchronize(cp); cudaStreamSynchronize(st); cleanup(); return rc; }

    // ---- host side: worker w expands its slice of every chunk as soon as the chunk has landed ---------------------------------
    Lut lut;
    build_lut(lut, *cfg);
    int W = n_threads > 0 ? n_threads : allowed_cpus();
    if (W > 32) W = 32;
    if (W < 1) W = 1;
    std::atomic<int> landed{0};
    std::atomic<bool> failed{false};
    const bool full = cfg->full_observable != 0;
    const size_t obs_row = full ? 94 : 47;
    auto worker = [&](int w) {
        for (int k = 0; k < n_chunks; ++k) {
            while (landed.load(std::memory_order_acquire) <= k) {
                if (failed.load(std::memory_order_relaxed)) return;
                std::this_thread::yield();
            }
            const int t0 = k * chunk, tc = (T - t0 < chunk) ? T - t0 : chunk;
            const size_t off = (size_t)t0 * rows, n = (size_t)tc * rows;
            // slices in units of 4 rows so that every slice starts on a 16-byte boundary of the obs array and on an env pair
            const size_t units = (n + 3) / 4, u0 = units * w / W, u1 = units * (w + 1) / W;
            const size_t r0 = u0 * 4, r1 = u1 * 4 < n ? u1 * 4 : n;
            if (r0 < r1)
                expand_rows(reinterpret_cast<const unsigned long long*>(codes_host) + off, r0, r1, obs_host ? obs_host + off * obs_row : nullptr,
                            rewards_host ? rewards_host + off : nullptr, dones_host ? dones_host + off : nullptr, full, lut);
        }
        _mm_sfence();                                           // this thread's non-temporal stores are globally visible
    };
    std::vector<std::thread> pool;
    pool.reserve(W - 1);
    for (int w = 1; w < W; ++w) pool.emplace_back(worker, w);
    for (int k = 0; k < n_chunks; ++k) {                       // this thread waits for the copies and takes slice 0
        if (cudaEventSynchronize(drained[k]) != cudaSuccess) { failed.store(true); rc = MM_ERR_LAUNCH; break; }
        landed.store(k + 1, std::memory_order_release);
        const int t0 = k * chunk, tc = (T - t0 < chunk) ? T - t0 : chunk;
        const size_t off = (size_t)t0 * rows, n = (size_t)tc * rows;
        const size_t units = (n + 3) / 4, u1 = units / W;
        const size_t r1 = u1 * 4 < n ? u1 * 4 : n;
        if (r1 > 0)
            expand_rows(reinterpret_cast<const unsigned long long*>(codes_host) + off, 0, r1, obs_host ? obs_host + off * obs_row : nullptr,
                        rewards_host ? rewards_host + off : nullptr, dones_host ? dones_host + off : nullptr, full, lut);
    }
    for (auto& t : pool) t.join();
    _mm_sfence();
    cleanup();
    return rc;
}
