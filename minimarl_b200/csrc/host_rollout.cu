// Host-buffer rollout: the env rollout as a caller holding host arrays sees it (bench "e2e").
// The T steps are cut into chunks; chunk k's kernel runs on `stream` while the outputs of chunk k-1 drain
// to the host on `copy_stream`, so the PCIe link never waits for the kernel and the kernel never waits for
// more than one staging buffer.
#include "common.cuh"

namespace {

struct EventPair {
    cudaEvent_t ready[2] = {nullptr, nullptr};    // chunk written by the kernel
    cudaEvent_t drained[2] = {nullptr, nullptr};  // chunk copied to the host, staging reusable
    bool ok = true;
    EventPair() {
        for (int i = 0; i < 2; ++i) {
            ok &= cudaEventCreateWithFlags(&ready[i], cudaEventDisableTiming) == cudaSuccess;
            ok &= cudaEventCreateWithFlags(&drained[i], cudaEventDisableTiming) == cudaSuccess;
        }
    }
    ~EventPair() {   // destroying a pending event is legal: resources are released once it completes
        for (int i = 0; i < 2; ++i) {
            if (ready[i]) cudaEventDestroy(ready[i]);
            if (drained[i]) cudaEventDestroy(drained[i]);
        }
    }
};

#define MM_CUDA(call)                      \
    do {                                   \
        cudaError_t e_ = (call);           \
        if (e_ != cudaSuccess) {           \
            mm::set_last_error(e_);        \
            return MM_ERR_LAUNCH;          \
        }                                  \
    } while (0)

}  // namespace

extern "C" int mm_checkers_rollout_host(uint32_t* state, const int32_t* actions_host, float* obs_host,
                                        float* rewards_host, uint8_t* dones_host, int32_t* actions_stage,
                                        float* obs_stage, float* rewards_stage, uint8_t* dones_stage,
                                        int32_t* bad_action_flag, int32_t T, int32_t chunk,
                                        const mm_checkers_cfg* cfg, void* stream, void* copy_stream) {
    if (state == nullptr || cfg == nullptr || actions_host == nullptr || actions_stage == nullptr) return MM_ERR_BAD_ARG;
    if (T <= 0 || chunk <= 0 || cfg->n_envs <= 0 || copy_stream == stream) return MM_ERR_BAD_ARG;
    if ((obs_host == nullptr) != (obs_stage == nullptr) || (rewards_host == nullptr) != (rewards_stage == nullptr) ||
        (dones_host == nullptr) != (dones_stage == nullptr))
        return MM_ERR_BAD_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream), cp = static_cast<cudaStream_t>(copy_stream);
    const size_t rows = (size_t)cfg->n_envs * MM_CHECKERS_AGENTS;                       // per step
    const size_t obs_row = (size_t)MM_CHECKERS_OBS_DIM * (cfg->full_observable ? 2 : 1);
    EventPair ev;
    if (!ev.ok) return MM_ERR_LAUNCH;
    int k = 0;
    for (int t0 = 0; t0 < T; t0 += chunk, ++k) {
        const int b = k & 1, tc = (T - t0 < chunk) ? T - t0 : chunk;
        const size_t off = (size_t)t0 * rows, stage = (size_t)b * chunk * rows, n = (size_t)tc * rows;
        if (k >= 2) MM_CUDA(cudaStreamWaitEvent(st, ev.drained[b], 0));
        MM_CUDA(cudaMemcpyAsync(actions_stage + stage, actions_host + off, n * sizeof(int32_t),
                                cudaMemcpyHostToDevice, st));
        if (int rc = mm_checkers_rollout(state, actions_stage + stage, nullptr,
                                         obs_stage ? obs_stage + stage * obs_row : nullptr,
                                         rewards_stage ? rewards_stage + stage : nullptr,
                                         dones_stage ? dones_stage + stage : nullptr, bad_action_flag, tc, 0, 0, cfg,
                                         stream))
            return rc;
        MM_CUDA(cudaEventRecord(ev.ready[b], st));
        MM_CUDA(cudaStreamWaitEvent(cp, ev.ready[b], 0));
        if (obs_host)
            MM_CUDA(cudaMemcpyAsync(obs_host + off * obs_row, obs_stage + stage * obs_row, n * obs_row * sizeof(float),
                                    cudaMemcpyDeviceToHost, cp));
        if (rewards_host)
            MM_CUDA(cudaMemcpyAsync(rewards_host + off, rewards_stage + stage, n * sizeof(float),
                                    cudaMemcpyDeviceToHost, cp));
        if (dones_host)
            MM_CUDA(cudaMemcpyAsync(dones_host + off, dones_stage + stage, n, cudaMemcpyDeviceToHost, cp));
        MM_CUDA(cudaEventRecord(ev.drained[b], cp));
    }
    // the caller synchronises `stream` only: make it wait for the last two drains
    MM_CUDA(cudaStreamWaitEvent(st, ev.drained[(k - 1) & 1], 0));
    if (k >= 2) MM_CUDA(cudaStreamWaitEvent(st, ev.drained[k & 1], 0));
    return MM_OK;
}
