/*
 * minimarl_b200 -- C ABI of libminimarl_b200.so (sm_100a kernels behind mini-marl's Python surface).
 *
 * mini-marl has no plugin/FFI interface of its own: the boundary is its Python call
 * surface.  Each entry point below names the reference callable (file:line under
 * singforai/mini-marl) whose arithmetic it replaces.  The Python package `minimarl_b200` binds
 * these symbols with ctypes (INTEGRATION.md shows the stub) and mirrors the reference's classes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the name ends in `_host`;
 *     the library never allocates or frees tensor memory and keeps no global mutable state;
 *   - every call is asynchronous and ordered on `stream` (a cudaStream_t passed as void*);
 *   - return value: 0 (MM_OK) on success, a negative mm_status on a bad argument or a failed
 *     launch; `mm_status_string` / `mm_last_cuda_error` give the message the wrapper raises;
 *   - row-major tensors; N = lockstep env instances, A = agents, D = per-agent obs length,
 *     T = rollout length, R = N*A rows.
 */
#ifndef MINIMARL_B200_H_
#define MINIMARL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum mm_status {
    MM_OK = 0,
    MM_ERR_BAD_ARG = -1,     /* null pointer, non-positive size, unsupported shape */
    MM_ERR_LAUNCH = -2,      /* cudaGetLastError() != cudaSuccess after the launch */
    MM_ERR_UNSUPPORTED = -3  /* shape outside what the kernel was built for */
} mm_status;

const char* mm_status_string(int status);
const char* mm_last_cuda_error(void);   /* cudaGetErrorString of the last failed launch (thread-local) */
int mm_version(void);                   /* ABI version, bumped on any signature change */

/* ------------------------------------------------------------------------------------------------
 * Checkers-v0 (ma-gym 0.0.14 `Checkers`, un-vendored; call sites vdn/main.py:61-64,83,93,
 * mappo/runner/shared/magym_runner.py:52-57,98,108, offpolicy/runner/shared/magym_runner.py:55,97).
 *
 * Env state is structure-of-arrays: `state` points at MM_CHECKERS_STATE_WORDS planes of n_envs
 * uint32 each (plane p of env e at state[p * n_envs + e]):
 *   plane 0  food bitmap, bit (row*8+col) set while the cell still holds its apple/lemon
 *            (apple iff (row+col) even); bit 31 = "episode finished" latch (ma-gym's
 *            steps_beyond_done is not None)
 *   plane 1  agent cells, 8 bits per agent (row*8+col), agent 0 in the low byte
 *   plane 2  step_count
 * ---------------------------------------------------------------------------------------------- */
#define MM_CHECKERS_STATE_WORDS 3
#define MM_CHECKERS_AGENTS 2
#define MM_CHECKERS_OBS_DIM 47   /* 2 + 3*3*5; full_observable doubles it */

typedef struct mm_checkers_cfg {
    int32_t n_envs;
    int32_t max_steps;
    int32_t full_observable;  /* 0: obs [N,2,47]; 1: every agent gets concat(obs_0, obs_1): [N,2,94] */
    int32_t auto_reset;       /* 1: an env that finishes is reset inside the same step and the obs
                                 written is the reset obs (reference never auto-resets: default 0) */
    float reward[2][3];       /* [agent][0 none, 1 lemon, 2 apple] = float32(step_cost (+/- food)) */
    float pos_row[3];         /* float32(round(r / 3, 2)) */
    float pos_col[8];         /* float32(round(c / 7, 2)) */
} mm_checkers_cfg;

/* env.reset() for every env (reset_mask == NULL) or for the envs whose mask byte is non-zero.
 * obs may be NULL (state only). */
int mm_checkers_reset(uint32_t* state, const uint8_t* reset_mask, float* obs,
                      const mm_checkers_cfg* cfg_host, void* stream);

/* One lockstep env.step(actions) over all envs.
 *   actions [N,2] int32 in [0,5)  (0 down, 1 left, 2 up, 3 right, 4 no-op)
 *   obs     [N,2,D] float32, rewards [N,2] float32, dones [N,2] uint8
 *   bad_action_flag: optional int32[1]; set to 1 if any action was outside [0,5) (the reference
 *   raises Exception('Action Not found!'); such an action is executed as a no-op here).
 * obs/rewards/dones may point straight into a rollout or replay buffer slot (zero-copy insert). */
int mm_checkers_step(uint32_t* state, const int32_t* actions, float* obs, float* rewards,
                     uint8_t* dones, int32_t* bad_action_flag,
                     const mm_checkers_cfg* cfg_host, void* stream);

/* T lockstep steps in one launch, state held in registers between steps.
 *   actions [T,N,2] int32, or NULL to draw them uniformly on the device (Philox4x32-10 keyed by
 *   (seed, env, agent), counter = step_offset + t); drawn actions are written to actions_out
 *   [T,N,2] when it is not NULL.
 *   obs [T,N,2,D] (obs after step t), rewards [T,N,2], dones [T,N,2]; any of the three may be
 *   NULL to skip that output.  Strides between consecutive t are the natural ones. */
int mm_checkers_rollout(uint32_t* state, const int32_t* actions, int32_t* actions_out,
                        float* obs, float* rewards, uint8_t* dones, int32_t* bad_action_flag,
                        int32_t T, uint64_t seed, uint64_t step_offset,
                        const mm_checkers_cfg* cfg_host, void* stream);

/* The same rollout for a caller that holds HOST arrays, as the reference's runners do (they hand
 * env.step host lists / numpy arrays: mappo/runner/shared/magym_runner.py:52-57,
 * offpolicy/runner/shared/magym_runner.py:94-97).  `*_host` pointers are host memory (page-locked for
 * the copies to overlap; pageable memory works but serialises), laid out like mm_checkers_rollout's
 * tensors for all T steps.  `*_stage` are caller-owned DEVICE staging buffers for two chunks:
 * actions_stage [2,chunk,N,2] int32, obs_stage [2,chunk,N,2,D], rewards_stage / dones_stage
 * [2,chunk,N,2].  An output is skipped when both its host and stage pointers are NULL.  Steps are
 * processed in chunks of `chunk`: H2D of the chunk's actions and its kernel on `stream`, D2H of its
 * outputs on `copy_stream` (a different stream) while the next chunk computes.  On return `stream`
 * has been made to wait for the last copy, so synchronising `stream` is enough. */
int mm_checkers_rollout_host(uint32_t* state, const int32_t* actions_host, float* obs_host,
                             float* rewards_host, uint8_t* dones_host, int32_t* actions_stage,
                             float* obs_stage, float* rewards_stage, uint8_t* dones_stage,
                             int32_t* bad_action_flag, int32_t T, int32_t chunk,
                             const mm_checkers_cfg* cfg_host, void* stream, void* copy_stream);

/* The same host-array rollout with a compact device-to-host transport: the kernel emits one 64-bit word per agent-step
 * (45 neighbourhood bits | row | col | reward kind | done: 8 B instead of 193 B over PCIe) and n_threads host threads
 * (<= 0: one per CPU this process may run on, at most 32) expand the words into the same obs / rewards / dones arrays,
 * bit-identical to mm_checkers_rollout_host, chunk by chunk while the device works on the next chunk.  Returns after every
 * output landed (it synchronises copy_stream itself).  codes_host: page-locked uint64 [T][N][2]; codes_stage: device
 * uint64 [2][chunk][N][2]; actions_stage: device int32 [2][chunk][N][2]. */
int mm_checkers_rollout_host_compact(uint32_t* state, const int32_t* actions_host, float* obs_host, float* rewards_host,
                                     uint8_t* dones_host, uint64_t* codes_host, int32_t* actions_stage, uint64_t* codes_stage,
                                     int32_t* bad_action_flag, int32_t T, int32_t chunk, int32_t n_threads,
                                     const mm_checkers_cfg* cfg_host, void* stream, void* copy_stream);

/* ------------------------------------------------------------------------------------------------
 * Switch{2,4}-v0 (ma-gym 0.0.14 `Switch`, un-vendored; qmix default env, qmix/_config.py:13-18,
 * qmix/main.py:66-71).  State planes: 0 = agent cells (8 bits per agent), 1 = per-agent done bits,
 * 2 = step_count.
 * ---------------------------------------------------------------------------------------------- */
#define MM_SWITCH_STATE_WORDS 3

typedef struct mm_switch_cfg {
    int32_t n_envs;
    int32_t n_agents;         /* 2..4 */
    int32_t max_steps;
    int32_t full_observable;
    int32_t clock;            /* 1: obs carries step_count/max_steps (ma-gym default) */
    float step_cost;          /* float32(step_cost) */
    float goal_reward;        /* float32(5) */
    float pos_row[3];         /* float32(round(r / 2, 2)) */
    float pos_col[7];         /* float32(round(c / 6, 2)) */
} mm_switch_cfg;

int mm_switch_reset(uint32_t* state, const uint8_t* reset_mask, float* obs,
                    const mm_switch_cfg* cfg_host, void* stream);
int mm_switch_step(uint32_t* state, const int32_t* actions, float* obs, float* rewards,
                   uint8_t* dones, int32_t* bad_action_flag,
                   const mm_switch_cfg* cfg_host, void* stream);
/* T lockstep steps in one launch, state in registers between steps; arguments as mm_checkers_rollout (actions [T,N,A] or
 * NULL for device-drawn uniform actions, obs [T,N,A,D] / rewards [T,N,A] / dones [T,N,A], any of them NULL to skip). */
int mm_switch_rollout(uint32_t* state, const int32_t* actions, int32_t* actions_out, float* obs, float* rewards,
                      uint8_t* dones, int32_t* bad_action_flag, int32_t T, uint64_t seed, uint64_t step_offset,
                      const mm_switch_cfg* cfg_host, void* stream);

/* ------------------------------------------------------------------------------------------------
 * GAE / returns reverse scan: SharedReplayBuffer.compute_returns
 * (mappo/runner/shared/shared_buffer.py:131-157) with ValueNorm.denormalize folded in
 * (mappo/utils/valuenorm.py:67-78).
 *   rewards [T,R], value_preds [T+1,R] (value_preds[T] must already hold next_value),
 *   masks [T+1,R], returns [T+1,R] (rows 0..T-1 written; use_gae==0 also reads returns[T]).
 *   denorm_scale/denorm_shift: sqrt(var) and mean of the value normaliser, or 1/0 when neither
 *   PopArt nor ValueNorm is on.  advantages (optional, [T,R]) = returns - denorm(value_preds).
 *   gamma / gae_lambda are the reference's Python floats (doubles).  The arithmetic follows the
 *   reference's mixed precision operation by operation: delta in float32, the gae accumulator in
 *   float64, the stored return rounded to float32.
 * ---------------------------------------------------------------------------------------------- */
int mm_gae_returns(const float* rewards, const float* value_preds, const float* masks,
                   float* returns, float* advantages, int32_t T, int64_t R,
                   double gamma, double gae_lambda, int32_t use_gae, int32_t use_denorm,
                   float denorm_scale, float denorm_shift, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Episode replay on the device (offpolicy/utils/segment_tree.py, offpolicy/utils/rec_buffer.py).
 * Trees: float64[2 * capacity], node 1 = root, leaf i at capacity + i (the reference's layout).
 * ---------------------------------------------------------------------------------------------- */

/* SegmentTree.__setitem__ (segment_tree.py:74-89) on the sum and/or the min tree (either may be NULL):
 * leaves idx[i] (or idx_base + i when idx == NULL) <- values[i] (or scalar_value when values == NULL),
 * then every ancestor is recomputed level by level. */
int mm_segtree_set(double* sum_tree, double* min_tree, int32_t capacity, const int64_t* idx,
                   const double* values, double scalar_value, int64_t idx_base, int32_t n, void* stream);

/* SegmentTree.reduce(start, end) (segment_tree.py:40-72): op over leaves [start, end), same association
 * order as the reference's recursion.  out: device double[1]. */
int mm_segtree_reduce(const double* tree, int32_t capacity, int32_t start, int32_t end, int32_t is_min,
                      double* out, void* stream);

/* SumSegmentTree.find_prefixsum_idx (segment_tree.py:115-146). */
int mm_segtree_find_prefixsum(const double* sum_tree, int32_t capacity, const double* mass,
                              int64_t* idx_out, int32_t n, void* stream);

/* PrioritizedRecReplayBuffer.sample (rec_buffer.py:272-296) minus the gather: mass = uniforms *
 * sum(0, len - 1), tree descent -> idx_out[n]; importance weights (p * len)^-beta / max_weight -> weights_out[n]. */
int mm_per_sample(const double* sum_tree, const double* min_tree, int32_t capacity, int32_t len,
                  const double* uniforms, int32_t n, double beta, int64_t* idx_out, double* weights_out,
                  void* stream);
/* The same with beta read from device memory when the kernel runs: a captured (CUDA graph) update follows the PER beta schedule
 * (offpolicy/runner/shared/base_runner.py: beta_anneal) without being re-captured. */
int mm_per_sample_dev(const double* sum_tree, const double* min_tree, int32_t capacity, int32_t len, const double* uniforms,
                      int32_t n, const double* beta_dev, int64_t* idx_out, double* weights_out, void* stream);

/* RecPolicyBuffer.sample_inds (rec_buffer.py:192-240): gather B episodes idx[b] out of the time-major
 * storage obs [T+1,S,A,D], share_obs [T+1,S,SD], acts [T,S,A,NA], rewards/dones [T,S,A,1],
 * dones_env [T,S,1] into out_obs [A,T+1,B,D], out_share_obs [T+1,B,SD], out_acts [A,T,B,NA],
 * out_rewards/out_dones [A,T,B,1], out_dones_env [T,B,1]. */
int mm_replay_gather(const float* obs, const float* share_obs, const float* acts, const float* rewards,
                     const float* dones, const float* dones_env, const int64_t* idx, float* out_obs,
                     float* out_share_obs, float* out_acts, float* out_rewards, float* out_dones,
                     float* out_dones_env, int32_t T, int32_t S, int32_t B, int32_t A, int32_t D, int32_t SD,
                     int32_t NA, void* stream);
/* The same with out_obs [T+1, A, B, D] and out_acts [T, A, B, NA] (rewards / dones stay agent-major): the (t, a * B + b) row order
 * QMix.train_policy_on_batch stacks the agents into (qmix.py:96-100) -- permuted views of these are the reference's [A, T, B, d]. */
int mm_replay_gather_stacked(const float* obs, const float* share_obs, const float* acts, const float* rewards, const float* dones,
                             const float* dones_env, const int64_t* idx, float* out_obs, float* out_share_obs, float* out_acts,
                             float* out_rewards, float* out_dones, float* out_dones_env, int32_t T, int32_t S, int32_t B,
                             int32_t A, int32_t D, int32_t SD, int32_t NA, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Recurrent agent network forward, fused over the whole sequence:
 * AgentQFunction.forward (offpolicy/algorithms/qmix/algorithm/agent_q_function.py:34-66) =
 * LayerNorm(D) -> [Linear(D,H) ReLU LayerNorm] -> [Linear(H,H) ReLU LayerNorm] -> GRU(H,H) ->
 * LayerNorm(H) -> Linear(H,NA); H in {32, 64} (MAPPO actor/critic trunk: mappo/algorithms/r_actor_critic.py).
 *   params  packed fp32 parameters, 16-byte aligned (layout: mm_agentq_param_offsets; Linear / GRU
 *           matrices stored TRANSPOSED [in][out], the D->H matrix zero-padded to a multiple of 4 rows,
 *           the head zero-padded to a multiple of 4 columns)
 *   obs [T1,R,D]; h0 [R,H] or NULL (zeros); q [T1,R,NA] or NULL; h_final [R,H] or NULL;
 *   rnn_out [T1,R,H] or NULL (the LayerNorm'ed GRU outputs that feed the head).
 * ---------------------------------------------------------------------------------------------- */
int mm_agentq_param_count(int32_t D, int32_t H, int32_t NA);
int mm_agentq_param_offsets(int32_t D, int32_t H, int32_t NA, int32_t* offsets_host /* int32[18] */);
int mm_agentq_forward(const float* params, const float* obs, const float* h0, float* q, float* h_final,
                      float* rnn_out, int32_t T1, int64_t R, int32_t D, int32_t H, int32_t NA, void* stream);

/* ------------------------------------------------------------------------------------------------
 * QMIX hypernetwork mixer forward (offpolicy/algorithms/qmix/algorithm/q_mixer.py:68-94,
 * hypernet_layers == 2, hypernet_hidden_dim == 64, mixer_hidden_dim == 32) and the VDN sum
 * (offpolicy/algorithms/vdn/algorithm/vdn_mixer.py:28-39).
 *   params packed fp32 (layout: mm_qmixer_param_offsets; matrices transposed [in][out]; the four
 *   first layers concatenated column-wise as hyper_w1.0 | hyper_w2.0 | hyper_b2.0 | hyper_b1 | 32 zero cols)
 *   agent_q [M,A], states [M,SD], q_tot [M]  (M = T*B rows);  A in {2, 4}.
 * ---------------------------------------------------------------------------------------------- */
int mm_qmixer_param_count(int32_t SD, int32_t A);
int mm_qmixer_param_offsets(int32_t SD, int32_t A, int32_t* offsets_host /* int32[8] */);
int mm_qmixer_forward(const float* params, const float* agent_q, const float* states, float* q_tot,
                      int64_t M, int32_t SD, int32_t A, void* stream);
int mm_vdn_mix(const float* agent_q, float* q_tot, int64_t M, int32_t A, void* stream);
/* Tensor-core version of mm_qmixer_forward for 2 agents: every hypernetwork contraction on tcgen05 (kind::tf32, fp32
 * accumulators in TMEM), persistent CTAs with the weights resident in shared memory.  SD even, SD <= 96.
 * params: wcat [224][96] (rows hyper_w1.0 | hyper_w2.0 | hyper_b2.0 | hyper_b1 weights, K zero-padded), w1b [64][64]
 * (hyper_w1.2), w2b [32][64] (hyper_w2.2), bcat [224], b1b [64], b2b [32], wb2 [64] (hyper_b2.2), bb2 [4] -- the PyTorch
 * [out][in] layouts; mm_qmixer_tc_param_offsets gives the 8 offsets in that order.  ~1e-3 relative on Q_tot. */
int mm_qmixer_tc_param_count(int32_t SD, int32_t A);
int mm_qmixer_tc_param_offsets(int32_t SD, int32_t A, int32_t* offsets_host);
int mm_qmixer_tc_forward(const float* params, const float* agent_q, const float* states, float* q_tot,
                         int64_t M, int32_t SD, int32_t A, void* stream);

/* The differentiated (live) mixer of QMix.train_policy_on_batch (offpolicy/algorithms/qmix/qmix.py:150-208 through
 * q_mixer.py:68-94): forward = mm_qmixer_tc_forward on a buffer packed by mm_qmixer_tc_pack straight from the module tensors;
 * backward = mm_qmixer_bwd, ONE kernel that recomputes the activations of a 128-row tile and contracts every weight gradient on
 * tcgen05 (csrc/mixer_bwd.cu), followed by mm_qmixer_bwd_unfold (sum over CTAs, scale removed, the 14 gradient tensors back to
 * back in module order; offsets from mm_qmixer_bwd_grad_layout).
 *   params14: HOST array of the 14 device pointers in QMixer.parameters() order (hyper_w1.0.weight, .bias, hyper_w1.2.*,
 *             hyper_w2.0.*, hyper_w2.2.*, hyper_b1.*, hyper_b2.0.*, hyper_b2.2.*), weights [out][in] fp32
 *   agent_q [M][2], states [M][SD] (SD even, <= 94), dout [M] = dL/dQ_tot, scale = mm_ppo_trunk2_scale(dout); dq [M][2]
 *   acc [n_acc_ctas >= mm_qmixer_bwd_ctas(M)][128][dims[0]], vec [n_acc_ctas][dims[1]]: per-CTA accumulators (scratch)
 *   mm_qmixer_bwd_dims -> {accumulator columns, vector floats, tile rows, gradient floats}. */
int mm_qmixer_tc_pack(const void* const* params14, float* packed, int32_t SD, int32_t n_agents, void* stream);
int mm_qmixer_bwd_dims(int32_t SD, int32_t* dims4);
int mm_qmixer_bwd_ctas(int64_t M);
int mm_qmixer_bwd_grad_layout(int32_t SD, int32_t* offsets15);
int mm_qmixer_bwd(const void* const* params14, const float* agent_q, const float* states, const float* dout, const float* scale,
                  float* dq, float* acc, float* vec, int32_t n_acc_ctas, int64_t M, int32_t SD, int32_t n_agents, void* stream);
int mm_qmixer_bwd_unfold(const float* acc, const float* vec, int32_t n_ctas, const float* scale, float* grads, int32_t SD,
                         void* stream);
int mm_qmixer_bwd_unfold_into(const float* acc, const float* vec, int32_t n_ctas, const float* scale, void* const* grads14,
                              int32_t SD, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Tail of QMix.train_policy_on_batch (offpolicy/algorithms/qmix/qmix.py:134-191).
 * mm_q_select: rows r = a*B + b.  q_all / target_q_all [T+1,R,NA], acts [T,R,NA] (one-hot rows; arg-max
 *   with first-maximum tie-break like torch.max) -> q_taken [T,B,A] (q_values_from_actions,
 *   QMixPolicy.py:92-111), act_idx [T,R], greedy [T+1,R] (live arg-max), next_q [T,B,A] = target Q at
 *   step t+1 of the live greedy action (double_q) or its max over actions.  Outputs may be NULL.
 * mm_qmix_td: q_tot / next_q_tot / rewards / dones_env [T,B], weights [B] or NULL ->
 *   err [T,B] = (q_tot - (r + (1-d) gamma next)) * (1 - bad), bad[t] = dones_env[t-1];
 *   ep_loss [B] = w_b * sum_t err^2; ep_valid [B] = sum_t (1 - bad); ep_qsum [B] = sum_t q_tot (1 - bad);
 *   priorities [B] = (1-nu) mean_t|err| + nu max_t|err| + eps  (may be NULL).
 * ---------------------------------------------------------------------------------------------- */
int mm_q_select(const float* q_all, const float* target_q_all, const float* acts, float* q_taken,
                float* next_q, int64_t* act_idx, int64_t* greedy, int32_t T, int32_t B, int32_t A,
                int32_t NA, int32_t double_q, void* stream);
int mm_qmix_td(const float* q_tot, const float* next_q_tot, const float* rewards, const float* dones_env,
               const float* weights, float gamma, float per_nu, float per_eps, int32_t T, int32_t B,
               float* err, float* ep_loss, float* ep_valid, float* ep_qsum, float* priorities, void* stream);

/* Self-test of the tcgen05 / TMEM building blocks: D[128,N] = A[128,K] * B[N,K]^T in TF32 with fp32
 * accumulation (one CTA).  N % 16 == 0, 16 <= N <= 256; K % 8 == 0, K <= 128. */
int mm_tc_gemm_selftest(const float* A, const float* B, float* D, int32_t N, int32_t K, void* stream);
/* Raw probe: 64 KB shared-memory images (float[16384] each) and descriptor fields (bytes) supplied by the caller. */
int mm_tc_raw_selftest(const float* imgA, const float* imgB, float* D, int32_t N, int32_t steps, int32_t a_step,
                       int32_t a_lbo, int32_t a_sbo, int32_t b_step, int32_t b_lbo, int32_t b_sbo, int32_t trans_a,
                       int32_t trans_b, void* stream);

/* The same probe for kind::f16 (fp16 operands, fp32 accumulate): images are 64 KB of raw bytes, M = 64 or 128, trans_a /
 * trans_b set the MN-major bits; D [128,N] is the dump of all 128 TMEM lanes (pre-filled with -12345). */
int mm_tc_raw_f16_selftest(const void* imgA, const void* imgB, float* D, int32_t M, int32_t N, int32_t steps, int32_t a_step,
                           int32_t a_lbo, int32_t a_sbo, int32_t b_step, int32_t b_lbo, int32_t b_sbo, int32_t trans_a,
                           int32_t trans_b, void* stream);

/* Lane probe: the same product with the instruction descriptor's M = 64 or 128 (A has M rows); every TMEM lane holds
 * -12345 before the MMA and D [128,N] is the dump of all 128 lanes afterwards: shows where a 64-row accumulator lives. */
int mm_tc_lane_probe(const float* A, const float* B, float* D, int32_t M, int32_t N, int32_t K, void* stream);

/* Tensor-core version of mm_agentq_forward: every dense contraction on tcgen05 (kind::tf32, fp32
 * accumulators in TMEM), same arguments.  params: plain fp32 buffer of the PyTorch-layout tensors
 * (offsets from mm_agentq_tc_param_offsets, same order of names as mm_agentq_param_offsets; the
 * Linear(D,H) weight is [H][K1] with K1 = D rounded up to 8 and zero padding).  Returns
 * MM_ERR_UNSUPPORTED when the tile does not fit shared memory (e.g. D = 94): use mm_agentq_forward. */
int mm_agentq_tc_param_count(int32_t D, int32_t H, int32_t NA);
int mm_agentq_tc_param_offsets(int32_t D, int32_t H, int32_t NA, int32_t* offsets_host /* int32[18] */);
int mm_agentq_tc_forward(const float* params, const float* obs, const float* h0, float* q, float* h_final,
                         float* rnn_out, int32_t T1, int64_t R, int32_t D, int32_t H, int32_t NA, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused policy + env rollout (the per-step body of MAGYM_Runner.collect_rollout,
 * offpolicy/runner/shared/magym_runner.py:75-127, and of the vdn/qmix main loops, vdn/main.py:139-177):
 * T steps of  agent network forward (tcgen05) -> eps-greedy action -> Checkers step -> next obs,
 * for n_envs lockstep boards in ONE launch; nothing returns to the host between steps.
 * Rows are (env, agent) pairs, row = 2*env + agent.  Every "t stride" is in elements between consecutive t
 * so that outputs can land directly in replay-buffer slots ([T, S, ...] storage with S != n_envs).
 * ---------------------------------------------------------------------------------------------- */
typedef struct mm_rollout_io {
    uint32_t* env_state;        /* Checkers state planes [3][n_envs] (read unless do_reset, always written back) */
    float* obs;                 /* slab t at obs + t*obs_t_stride, each [n_envs,2,47]; slabs 0..T are written
                                   (slab 0 only when do_reset, otherwise it must hold the current observations) */
    int64_t obs_t_stride;
    int32_t* act_idx;           /* [T] slabs of [n_envs*2] chosen action indices, or NULL */
    int64_t act_t_stride;
    float* act_onehot;          /* [T] slabs of [n_envs*2, NA] one-hot actions, or NULL */
    int64_t onehot_t_stride;
    float* rewards;             /* [T] slabs of [n_envs*2], or NULL */
    int64_t rew_t_stride;
    float* dones;               /* [T] slabs of [n_envs*2] float32 0/1, or NULL */
    int64_t done_t_stride;
    float* dones_env;           /* [T] slabs of [n_envs] float32 0/1 (all agents done), or NULL */
    int64_t denv_t_stride;
    uint64_t step_offset;       /* Philox counter of step 0 (so consecutive rollouts draw fresh numbers) */
    float eps;                  /* exploration rate: P(random action) per draw */
    int32_t eps_per_env;        /* 0: one uniform per agent row (QMixPolicy.py:177-187); 1: one per env, all agents
                                   of the env explore together (vdn/_network.py:52-58) */
    int32_t do_reset;           /* 1: env.reset() first and write obs slab 0 */
    int32_t freeze_after_done;  /* 1: a finished episode is no longer stepped (rewards 0, obs frozen), like
                                   collect_rollout's `if not np.all(dones)` */
    int32_t common_reward;      /* 1: every agent receives the team reward (use_common_reward, base_runner.py:346-348);
                                   common_reward_lut[k0][k1] = float32(double(r0) + double(r1)) by food kind */
} mm_rollout_io;

int mm_qmix_rollout_tc(const float* params /* mm_agentq_tc layout, D = 47 */, const float* h0, float* h_final,
                       float* q_out /* [T, n_envs*2, NA] or NULL */, const mm_rollout_io* io_host,
                       const mm_checkers_cfg* cfg_host, const float* common_reward_lut_host /* float[9] or NULL */,
                       int32_t T, int64_t n_envs, int32_t H, int32_t NA, uint64_t seed, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Chunk-PER sum tree of the top-level vdn/ and qmix/ packages (vdn/replay_buffer/sumtree.py:8-66):
 * float64[2*capacity-1], heap order, root 0, leaf i at node capacity-1+i.
 *   update:      for each (node_idx[i], priority[i]) IN ORDER: propagate the change to the ancestors.
 *   sample:      stratified draw (buffer.py:42-66): s_i = seg*i + (seg*(i+1) - seg*i) * u_i, seg = total/n,
 *                `_retrieve_max` walk -> node index and its priority.
 *   argmin_leaf: first leaf with the minimum priority (the eviction rule of `add` on a full buffer).
 * ---------------------------------------------------------------------------------------------- */
int mm_heaptree_update(double* tree, int32_t capacity, const int64_t* node_idx, const double* priority,
                       int32_t n, void* stream);
int mm_heaptree_sample(const double* tree, int32_t capacity, const double* uniforms, int32_t n,
                       int64_t* node_idx_out, double* priority_out, void* stream);
int mm_heaptree_argmin_leaf(const double* tree, int32_t capacity, int64_t* leaf_out, void* stream);
/* k consecutive SumTree.add calls (vdn/replay_buffer/sumtree.py:38-58, qmix/replay_buffer/sumtree.py:38-64) in one
 * launch -- what N lockstep envs hand the buffer at every chunk boundary.  Item i goes to data slot n_data while the
 * buffer is filling (n_data = fill count before the call), afterwards to the arg-min-priority leaf as it is AFTER
 * items 0..i-1 went in (a later item of the batch may evict an earlier one).  slot_out[i] = data slot of item i.
 * MM_ERR_UNSUPPORTED when the sum tree (16 B per slot) does not fit in shared memory (capacity > ~14 000). */
int mm_heaptree_add_batch(double* tree, int32_t capacity, int32_t n_data, const double* priority, int32_t k,
                          int64_t* slot_out, void* stream);

/* Batch formulations for N lockstep envs (heaptree_batch.cu); the sequential calls above stay the reference-exact path.
 *   add_set:    the valid items (valid == NULL: all) first fill the free slots in order, the rest replace the lowest-priority
 *               slots occupied before the batch (ties: lowest slot; r-th lowest -> r-th remaining item); items of one batch
 *               never evict each other, items beyond capacity are dropped.  slot_out[i] = data slot or -1.
 *               n_data_dev: device int32 fill count, updated.  rebuild != 0: inner nodes are recomputed from the leaves
 *               (the insert itself only reads leaves, so a run of inserts needs one mm_heaptree_rebuild at its end).
 *               workspace: int32[k + 2 * capacity + 4].
 *   set_leaves: tree[node_idx[i]] = priority[i], the last i naming a leaf wins (the order of Replay_buffer.update calls in
 *               vdn/_train.py:96-99); node indices outside the leaf range are ignored; inner nodes recomputed.
 *               workspace: int32[capacity].
 *   rows_scatter: dst[t][slot[i]] = src[t][i] for n <= 8 row-major float tensors (row_floats[t] floats per row); slot -1 skips. */
int mm_heaptree_add_set(double* tree, int32_t capacity, int32_t* n_data_dev, const double* priority, const uint8_t* valid,
                        int32_t k, int64_t* slot_out, int32_t* workspace, int32_t rebuild, void* stream);
int mm_heaptree_rebuild(double* tree, int32_t capacity, void* stream);
/* One episode of N lockstep envs: C consecutive add_set calls (chunk c of every env, in the reference's add order) plus the
 * payload scatter of each (src tensors chunk-major [C][N][row_floats]), one host call; inner nodes recomputed at the end.
 * priority / valid / slot_out: [C][N]; workspace: int32[N + 2 * capacity + 4]. */
int mm_chunk_per_insert_episode(double* tree, int32_t capacity, int32_t* n_data_dev, const double* priority, const uint8_t* valid,
                                int32_t C, int32_t N, int64_t* slot_out, int32_t* workspace, const float* const* src_host,
                                float* const* dst_host, const int32_t* row_floats_host, int32_t n_tensors, void* stream);
int mm_heaptree_set_leaves(double* tree, int32_t capacity, const int64_t* node_idx, const double* priority, int32_t n,
                           int32_t* workspace, void* stream);
int mm_rows_scatter(const float* const* src_host, float* const* dst_host, const int32_t* row_floats_host, int32_t n,
                    const int64_t* slot, int32_t k, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Top-level vdn/ and qmix/ packages: fused per-agent Q_Net and recurrent Mix_Net.
 * Networks are passed as HOST arrays of device pointers to the reference-layout tensors (nn.Linear / nn.GRUCell
 * weight [out][in]); the kernels stage them into shared memory themselves, so optimiser steps need no re-packing.
 *   Q_Net  (vdn/_network.py:61-91): per agent w1 [64,D] b1 w2 [32,64] b2 wih [96,32] whh [96,32] bih bhh wo [NA,32] bo;
 *          params_host = 20 pointers (agent 0 then agent 1).  A = 2 agents, D <= 96, NA <= 8.
 *   Mix_Net (qmix/_network.py:172-220, recurrent): gru.weight_ih [96,S] weight_hh [96,32] bias_ih bias_hh,
 *          hyper_net_weight_1 {w [64,32], b}, hyper_net_weight_2 {w [32,32], b}, hyper_net_bias_1 {w, b},
 *          hyper_net_bias_2.0 {w, b}, hyper_net_bias_2.2 {w [1,32], b [1]}: 14 pointers.  S = A * D <= 96.
 *
 * mm_toplevel_rollout: the rollout body of vdn/main.py:139-177 / qmix/main.py:183-240 for N lockstep envs, T = C * L steps
 *   in ONE launch (env reset at the start): behaviour Q_Net -> epsilon_greedy with one uniform per env
 *   (vdn/_network.py:52-58; Philox4x32-10 keyed by seed, counter (step_offset + t, env)) -> env.step
 *   (env_kind 1 = Checkers-v0 with mm_checkers_cfg, 2 = Switch2-v0 with mm_switch_cfg) -> next observation.
 *   Outputs in chunk layout: states / next_states [C,N,L,2,D], actions / rewards [C,N,L,2], dones [C,N,L] (all agents
 *   done), behv [T,N] = sum_a Q_beh[a][action_a], alive [T,N] (env not finished before step t), score [N] = sum of the
 *   rewards over the alive steps.  forced_actions [T,N,2] (optional) replaces the epsilon-greedy choice.
 * mm_toplevel_qnet_seq: Q_Net over recorded observations.  obs element (c, b, l, agent, d) at
 *   c * chunk_stride + ((row_index[b] * L + l) * 2 + agent) * D + d (row_index NULL = identity); the GRU state starts at
 *   zero, runs over all C * L steps and, with reset_on_done, is zeroed after steps whose dones entry
 *   (c * done_chunk_stride + row_index[b] * L + l) is non-zero.  q_out [C,B,L,2,NA]; save (optional) receives the
 *   activations the backward needs: [2, C*L, B, mm_toplevel_qnet_save_stride()].
 * mm_toplevel_qnet_bwd: BPTT through one chunk (C = 1) given dq [B,L,2,NA]; writes (accumulate = 0) or adds the weight
 *   gradients into grads_host (20 pointers, same order and layout as the parameters).  partials: workspace of
 *   2 * ceil(B / 32) * mm_toplevel_qnet_param_count(D, NA) floats.  Deterministic (fixed reduction order).
 * mm_toplevel_td: cal_td_error (vdn/_utils.py:44-52) summed per chunk over the steps at which the env was alive.
 * mm_vdn_chunk_loss: Target_Dqn.train's per-step targets and squared errors (vdn/_train.py:66-86, quirk Q7), dq and the
 *   last-step |TD| (quirk Q9); actions / rewards / dones are read through row_index from the PER payload.
 * mm_toplevel_mix_seq / _bwd, mm_qmix_chunk_loss: the same for Train_dqn.train (qmix/_train.py:19-121): q fed to the
 *   mixer is the taken action's value (actions != NULL) or max_o (target side); the mixer state is zeroed after terminal
 *   steps; _bwd also writes dq [B,L,2,NA] for mm_toplevel_qnet_bwd.
 * ---------------------------------------------------------------------------------------------- */
int mm_toplevel_qnet_param_count(int32_t D, int32_t NA);
int mm_toplevel_qnet_save_stride(void);
int mm_toplevel_rollout(const float* const* params_host, int32_t env_kind, const void* env_cfg_host, float* states,
                        float* next_states, float* actions, float* rewards, float* dones, float* behv, uint8_t* alive,
                        float* score, uint32_t* env_state, const int32_t* forced_actions, int32_t N, int32_t T, int32_t L,
                        int32_t D, int32_t NA, float eps, uint64_t seed, uint64_t step_offset, void* stream);
int mm_toplevel_qnet_seq(const float* const* params_host, const float* obs, const int64_t* row_index, int64_t chunk_stride,
                         const float* dones, int64_t done_chunk_stride, int32_t reset_on_done, float* q_out, float* save,
                         int32_t B, int32_t L, int32_t C, int32_t D, int32_t NA, void* stream);
int mm_toplevel_qnet_bwd(const float* const* params_host, float* const* grads_host, const float* obs, const int64_t* row_index,
                         const float* dones, int32_t reset_on_done, const float* save, const float* dq, float* partials,
                         int32_t accumulate, int32_t B, int32_t L, int32_t D, int32_t NA, void* stream);
int mm_toplevel_td(const float* q_tgt, const float* rewards, const float* dones, const float* behv, const uint8_t* alive,
                   float* td_sum, uint8_t* chunk_valid, int32_t N, int32_t L, int32_t C, int32_t NA, float gamma, void* stream);
int mm_vdn_chunk_loss(const float* q_beh, const float* q_tgt, const float* actions, const float* rewards, const float* dones,
                      const int64_t* row_index, const float* is_weight, float* se, float* dq, float* td_last, int32_t B,
                      int32_t L, int32_t NA, float gamma, void* stream);
int mm_toplevel_mix_param_count(int32_t S);
int mm_toplevel_mix_save_stride(void);
int mm_toplevel_mix_seq(const float* const* params_host, const float* obs, const int64_t* row_index, const float* dones,
                        const float* q_all, const float* actions, float* qtot, float* save, int32_t B, int32_t L, int32_t S,
                        int32_t NA, void* stream);
int mm_toplevel_mix_bwd(const float* const* params_host, float* const* grads_host, const float* obs, const int64_t* row_index,
                        const float* dones, const float* q_all, const float* actions, const float* save, const float* dqtot,
                        float* dq, float* partials, int32_t accumulate, int32_t B, int32_t L, int32_t S, int32_t NA, void* stream);
int mm_qmix_chunk_loss(const float* qtot, const float* qtot_tgt, const float* rewards, const float* dones,
                       const int64_t* row_index, const float* is_weight, float* se, float* dqtot, float* td_last, int32_t B,
                       int32_t L, float gamma, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LayerNorm over short rows (N <= 128, eps 1e-5) for the differentiated network pass: the nn.LayerNorm
 * calls of MLPLayer / RNNLayer (offpolicy/utils/algorithm_utils/mlp.py:17-29, rnn.py:18,28;
 * mappo/utils/algorithm_utils/mlp.py:17-30, rnn.py:22,78).
 *   fwd: y = (x - mean) * rstd * weight + bias; mean / rstd [M] are saved for the backward.
 *   bwd: dx [M,N]; dweight_accum / dbias_accum [N] are ADDED to (zero them first).
 * ---------------------------------------------------------------------------------------------- */
int mm_layernorm_fwd(const float* x, const float* weight, const float* bias, float* y, float* mean, float* rstd,
                     int64_t M, int32_t N, void* stream);
int mm_layernorm_bwd(const float* dy, const float* x, const float* weight, const float* mean, const float* rstd,
                     float* dx, float* dweight_accum, float* dbias_accum, int64_t M, int32_t N, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Recurrent half of nn.GRU (one layer, H in {32, 64}) on tcgen05, forward and BPTT backward, for the
 * differentiated pass of RNNLayer.forward (offpolicy/utils/algorithm_utils/rnn.py:26-29).
 *   fwd: gi [T,R,3H] = x W_ih^T + b_ih (caller, cuBLAS); h0 [R,H] or NULL; w_hh [3H,H]; b_hh [3H] ->
 *        h_seq [T,R,H]; r, z, n, ghn [T,R,H] saved for the backward (all four NULL for inference).
 *   bwd: dh_seq [T,R,H] (gradient w.r.t. every h_t) + the saved tensors -> dgi [T,R,3H] (gradient w.r.t. gi;
 *        its r and z thirds are also the gradient w.r.t. gh), dghn [T,R,H] (gradient w.r.t. gh_n), dh0 [R,H].
 *        dx, dW_ih, dW_hh, db are row-parallel GEMMs / sums the caller does with cuBLAS.
 * ---------------------------------------------------------------------------------------------- */
int mm_gru_seq_fwd(const float* gi, const float* h0, const float* w_hh, const float* b_hh, float* h_seq, float* r,
                   float* z, float* n, float* ghn, int32_t T, int64_t R, int32_t H, void* stream);
int mm_gru_seq_bwd(const float* dh_seq, const float* r, const float* z, const float* n, const float* ghn,
                   const float* h_seq, const float* h0, const float* w_hh, float* dgi, float* dghn, float* dh0,
                   int32_t T, int64_t R, int32_t H, void* stream);
/* The same with the fp16 consumers' gradient scale as a by-product: scale2 (device float[2]) receives max |dgi| in [1] and the
 * power of two mm_ppo_trunk2_scale(dgi) would compute in [0] -- the backward tracks the maximum while it writes dgi, so
 * mm_agent_mlp_bwd / mm_gru_wgrad need no pass of their own over the 3H-wide gradient. */
int mm_gru_seq_bwd_scaled(const float* dh_seq, const float* r, const float* z, const float* n, const float* ghn,
                          const float* h_seq, const float* h0, const float* w_hh, float* dgi, float* dghn, float* dh0,
                          float* scale2, int32_t T, int64_t R, int32_t H, void* stream);

/* The mixing step of QMixer.forward after the hypernetworks (q_mixer.py:80-93), forward and backward, for the
 * differentiated pass:  out[m] = elu(q[m,:] @ w1[m,:,:] + b1[m,:]) . w2[m,:] + b2[m]
 *   q [M,A]; w1 [M,A,32] (already |.|); b1 [M,32]; w2 [M,32] (already |.|); b2 [M]; A in {2, 4}. */
int mm_qmix_mixing_fwd(const float* q, const float* w1, const float* b1, const float* w2, const float* b2, float* out,
                       int64_t M, int32_t A, void* stream);
int mm_qmix_mixing_bwd(const float* dout, const float* q, const float* w1, const float* b1, const float* w2, float* dq,
                       float* dw1, float* db1, float* dw2, float* db2, int64_t M, int32_t A, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused MAPPO rollout (mappo/runner/shared/magym_runner.py:41-98 run/collect/insert, rmappo_policy.py:57-95
 * get_actions, r_actor_critic.py:60-93,189-208): T lockstep steps of actor + critic (hidden 32, local-obs critic)
 * + categorical sampling + Checkers step (auto-reset) in ONE launch, writing every SharedReplayBuffer field of
 * shared_buffer.py:82-117 in place.  Buffers are the contiguous SharedReplayBuffer tensors:
 *   obs, share_obs [T+1, N, 2, 47]; rnn_states, rnn_states_critic [T+1, N, 2, 1, 32]; masks [T+1, N, 2, 1];
 *   actions, action_log_probs, value_preds (first T slabs), rewards [T, N, 2, 1].
 * Slab 0 of obs / rnn_states / masks is the input (obs[0] is written when do_reset != 0).
 * params: mm_mappo_rollout_param_offsets layout; LayerNorm(47)'s affine is folded into the first Linear on the host.
 * Sampling: inverse CDF on softmax(logits) with one Philox4x32-10 uniform per (step_offset + t, row).
 * ---------------------------------------------------------------------------------------------- */
int mm_mappo_rollout_param_count(void);
int mm_mappo_rollout_param_offsets(int32_t* offsets19);
int mm_mappo_rollout_tc(const float* params, uint32_t* env_state, float* obs, float* share_obs, float* rnn_states,
                        float* rnn_states_critic, float* actions, float* action_log_probs, float* value_preds,
                        float* rewards, float* masks, int32_t T, int64_t n_envs, int32_t do_reset, uint64_t seed,
                        uint64_t step_offset, const mm_checkers_cfg* cfg_host, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused recurrent trunk of the PPO update (R_Actor.evaluate_actions / R_Critic.forward inside R_MAPPO.ppo_update:
 * mappo/algorithms/r_actor_critic.py:95-146,189-208, utils/algorithm_utils/mlp.py, rnn.py:24-80), hidden size 32:
 *   LayerNorm(D) -> [Linear ReLU LayerNorm] x2 -> GRU (state * masks[t] before every step) -> LayerNorm -> Linear(NO)
 * over L-step data chunks, rows = chunks.  obs [L,R,D] (D <= 48), h0 [R,32], masks [L,R] or NULL, out [L,R,NO] (NO <= 8).
 * params: mm_ppo_trunk_param_offsets layout, LayerNorm(D)'s affine folded into the first Linear by the caller.
 * mm_ppo_trunk_dims -> {XS, GS, n_acc, off_S, off_dbo, off_dgamma2, off_dgamma1, NOMAX}.
 * obs_row != NULL: the chunks are gathered in place -- obs is then a row table [*, D] (the rollout buffer's
 *   [T+1, N, A, D] storage) and step t of chunk r is table row obs_row[r] + t * obs_step_rows, which is how
 *   recurrent_generator (shared_buffer.py:318-427) addresses its chunks; nothing is materialised.  In this mode h0 and
 *   masks are tables as well: h0 row obs_row[r] of [*, 32] (rnn_states storage), masks entry obs_row[r] + t * obs_step_rows.
 * fwd, training mode (X != NULL): stores the operands of the weight gradients into X [L*R, XS]:
 *   x0 (normalised obs) [0,48) | x1 [48,80) | x2 [80,112) | hp (masked input state) [112,144); the caller owns columns
 *   [144,XS) and sets X[:,144] = 1 so that the bias gradients fall out of the same GEMM.
 * bwd: recomputes the forward per step (reads x0 and hp from X), needs dout [L,R,NO]; writes the pre-activation gradients
 *   G [L*R, GS] = dA1 [0,32) | dA2 [32,64) | dGi (r,z,n) [64,160) | dGh_n [160,192) and per CTA the column sums
 *   S[o][j] = sum dout[.,o] xhat3[.,j], dbo, dgamma2, dgamma1 into acc_partials [n_partials >= mm_ppo_trunk_max_ctas()][n_acc].
 *   All weight and bias gradients are blocks of the ONE large-K GEMM G^T X (caller, TF32 cuBLAS); the LayerNorm
 *   offsets' gradients follow from the bias gradients (dbeta2 = dbih Wih, dbeta1 = db2 W2, dbeta3 = dbo Wo).
 * ---------------------------------------------------------------------------------------------- */
int mm_ppo_trunk_param_count(void);
int mm_ppo_trunk_param_offsets(int32_t* offsets18);
int mm_ppo_trunk_dims(int32_t* dims8);
int mm_ppo_trunk_max_ctas(void);
/* Debug: CTA 0 of later launches accumulates clock64() deltas per kernel phase into dev_buf16 (int64[16]); NULL = off. */
int mm_ppo_trunk_set_prof(long long* dev_buf16);
int mm_ppo_trunk_fwd(const float* params, const float* obs, const int64_t* obs_row, int64_t obs_step_rows, const float* h0,
                     const float* masks, float* out, float* h_final, float* X, int32_t L, int64_t R, int32_t D, int32_t NO,
                     void* stream);
int mm_ppo_trunk_bwd(const float* params, const int64_t* obs_row, int64_t obs_step_rows, const float* masks, const float* X,
                     const float* dout, float* G, float* acc_partials, int32_t n_partials, int32_t L, int64_t R, int32_t D,
                     int32_t NO, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Second-generation fused trunk (csrc/ppo_trunk2.cu): the same function as mm_ppo_trunk_fwd / _bwd
 * (R_Actor.evaluate_actions / R_Critic.forward in the chunked mode, mappo/algorithms/r_actor_critic.py:95-146,189-208,
 * rnn.py:24-80; the update of mappo/algorithms/ramppo_network.py:103-209) with fp16 operands on tcgen05 (kind::f16, fp32
 * accumulate), thread = row, several tiles in flight per SM and EVERY weight gradient contracted inside the backward kernel
 * (G^T X as MN-major tcgen05.mma over the data path's own tiles, accumulated in TMEM) -- no X / G matrices, no GEMM afterwards.
 *   mm_ppo_obs_table: LayerNorm (no affine) of observation rows, obs fp32 [rows, D] -> xhat fp16 [rows, 48] (zero padded);
 *     once per update, shared by both networks and all epochs.  masks != NULL (D < 48): masks[row] rides in column 47 (the
 *     folded first Linear has a zero there) and the kernels, called with mask_in_table = 1, read it from the 32-byte sector
 *     their observation gather fetches anyway instead of gathering 4 bytes per row and step from the mask table.
 *   params: fp32 device buffer w1 [32][48] | w2 [32][32] | wih [96][32] | whh [96][32] | b1 [32] | b2 [32] | bih [96] | bhh [96] |
 *     wo [8][32] | bo [8] with every LayerNorm affine folded into the Linear that follows it (W' = W diag(g), b' = b + W beta).
 *   Chunks are addressed in place like recurrent_generator (shared_buffer.py:318-427): step t of chunk r is table row
 *     first[r] + t * step_rows of xhat / masks; h0 is row first[r] of a [rows, 32] table.
 *   fwd: out [L, R, NO]; saved != NULL (training): mm_ppo_trunk2_dims()[1] bytes per (128-row tile, step) of operand images
 *     (normalised activations, masked input state, 1 / sigma and ReLU bits) for the backward.
 *   bwd: dout [L, R, NO]; scale: device float, power of two applied to dout inside the kernel (fp16 range); acc
 *     [n_acc_ctas >= mm_ppo_trunk2_bwd_ctas(R)][128][dims[2]]: per CTA the TMEM accumulators
 *       columns [0, dims[3]):  lanes dr | dz | dn_i | dn_h (32 each)  x  xh2 [0,32) | hp [32,64) | 1 [64]
 *       columns [dims[3], ..): lanes dA1 | dA2 | xh3 (32 each) | 1 [96]  x  x0 [0,48) | xh1 [48,80) | dout [80,88) | 1 [95]
 *     i.e. scale * (gradients of the folded weights and biases); the caller sums over CTAs and unfolds the LayerNorm affines.
 *   mm_ppo_trunk2_dims -> {parameter count, saved bytes per (tile, step), accumulator columns, columns of the first
 *     accumulator, tile rows, 48, 8, 32};  mm_ppo_trunk2_set_slots: tiles in flight per SM (tuning; 0 = keep).
 * ---------------------------------------------------------------------------------------------- */
int mm_ppo_trunk2_dims(int32_t* dims8);
int mm_ppo_trunk2_set_slots(int32_t fwd_slots, int32_t bwd_slots);
int mm_ppo_trunk2_bwd_ctas(int64_t R);
/* Debug: page-locked host int32[4]; a bounded wait inside the trunk kernels that times out writes {code, block, thread, 1} there
 * before it traps (NULL = off). */
int mm_ppo_trunk2_set_debug(void* host_mapped_int4);
int mm_ppo_obs_table(const float* obs, const float* masks, void* xhat_f16, int64_t rows, int32_t D, void* stream);
/* The algebra around the two kernels as one launch each.  params18: HOST array of the 18 device pointers ln0_w, ln0_b, w1 [32][D],
 * b1, ln1_w, ln1_b, w2, b2, ln2_w, ln2_b, wih, whh, bih, bhh, ln3_w, ln3_b, wo [NO][32], bo (the module's own tensors).
 *   pack: folds the LayerNorm affines -> packed kernel parameters (dims[0] floats).
 *   unfold: sums acc over the CTAs, applies 1 / scale and unfolds the affines -> the 18 gradients back to back in grads
 *     (offsets from mm_ppo_trunk2_grad_layout, 19 entries); red: scratch of 128 * dims[2] floats. */
/* scale2 (device float[2]): [0] <- 2^floor(log2(64 / max |dout|)), the gradient scale mm_ppo_trunk2_bwd takes; [1] is scratch. */
int mm_ppo_trunk2_scale(const float* dout, int64_t n, float* scale2, void* stream);
int mm_ppo_trunk2_pack(const void* const* params18, float* packed, int32_t D, int32_t NO, void* stream);
int mm_ppo_trunk2_grad_layout(int32_t D, int32_t NO, int32_t* offsets19);
int mm_ppo_trunk2_unfold(const float* acc, int32_t n_ctas, const float* scale, float* red, const void* const* params18,
                         float* grads, int32_t D, int32_t NO, void* stream);
/* The same with every gradient ADDED to its own destination tensor (grads18: HOST array of the 18 device pointers, in the order of
 * params18): the parameters' .grad storage -- what autograd's AccumulateGrad would do one launch per tensor later. */
int mm_ppo_trunk2_unfold_into(const float* acc, int32_t n_ctas, const float* scale, float* red, const void* const* params18,
                              void* const* grads18, int32_t D, int32_t NO, void* stream);
int mm_ppo_trunk2_fwd(const float* params, const void* xhat_f16, const int64_t* first, int64_t step_rows, const float* h0,
                      const float* masks, int32_t mask_in_table, float* out, float* h_final, void* saved, int32_t L, int64_t R,
                      int32_t NO, void* stream);
int mm_ppo_trunk2_bwd(const float* params, const void* xhat_f16, const int64_t* first, int64_t step_rows, const float* masks,
                      int32_t mask_in_table, const void* saved, const float* dout, const float* scale, float* acc,
                      int32_t n_acc_ctas, int32_t L, int64_t R, int32_t NO, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Row-parallel half of the agent network's differentiated pass (csrc/agent_mlp.cu): what MLPBase / MLPLayer
 * (offpolicy/utils/algorithm_utils/mlp.py:7-89) and the input projection of the GRU in RNNLayer (rnn.py:4-47) compute
 * inside AgentQFunction.forward (offpolicy/algorithms/qmix/algorithm/agent_q_function.py:34-66) when
 * QMix.train_policy_on_batch (offpolicy/algorithms/qmix/qmix.py:80-210) differentiates the live network, hidden size 64:
 *   gi = Wih LN(ReLU(W2 LN(ReLU(W1 LN(obs) + b1)) + b2)) + bih          rows = (t, episode, agent) entries
 * forward and backward as one tcgen05 kernel each (fp16 operands, fp32 accumulate), weight gradients inside the backward.
 *   params: fp32 device buffer w1 [64][48] | w2 [64][64] | wih [192][64] | b1 [64] | b2 [64] | bih [192], LayerNorm affines folded
 *     into the Linear that follows (W' = W diag(g), b' = b + W beta) by the caller.
 *   fwd: obs [rows, D] (D <= 48) -> gi [rows, 192] (what mm_gru_seq_fwd takes); training: saved_rows [rows, dims[1]] (fp32
 *     normalised activations, 1 / sigma, ReLU bits) and saved_x0 [tiles][dims[6]] fp16 (operand images of the normalised
 *     observations); both NULL for inference.
 *   bwd: dgi [rows, 192] (from mm_gru_seq_bwd), scale: device float, power of two applied to dgi (mm_ppo_trunk2_scale);
 *     acc [n_acc_ctas >= mm_agent_mlp_bwd_ctas(rows)][128][dims[2]]: per CTA the TMEM accumulators, scale x
 *       columns [0,192): lanes xh2 feature j (0..63), lane 64 = 1      x gi column      -> dWih'^T, dbih'
 *       columns [192,256): lanes xh1 feature j, lane 64 = 1              x fc2 output    -> dW2'^T, db2'
 *       columns [256,320): a 64-row accumulator (row r in lane 32 (r / 16) + r % 16): rows x0 feature k (0..47), row 48 = 1
 *                                                                          x fc1 output    -> dW1'^T, db1'
 *   mm_agent_mlp_dims -> {parameter count, saved floats per row, accumulator columns, tile rows, 48, 64, saved x0 halves per tile, 0}.
 * ---------------------------------------------------------------------------------------------- */
int mm_agent_mlp_dims(int32_t* dims8);
/* The algebra around the two kernels, one launch each.  params12: HOST array of the 12 device pointers ln0_w, ln0_b, w1 [64][D], b1,
 * ln1_w, ln1_b, w2, b2, ln2_w, ln2_b, w_ih [192][64], b_ih (the module's own tensors).  pack folds the LayerNorm affines into the
 * packed parameters; unfold sums acc over the CTAs, applies 1 / scale and writes the 12 gradients back to back into grads
 * (offsets from mm_agent_mlp_grad_layout, 13 entries); red: scratch of 128 * dims[2] floats. */
int mm_agent_mlp_pack(const void* const* params12, float* packed, int32_t D, void* stream);
int mm_agent_mlp_grad_layout(int32_t D, int32_t* offsets13);
int mm_agent_mlp_unfold(const float* acc, int32_t n_ctas, const float* scale, float* red, const void* const* params12, float* grads,
                        int32_t D, void* stream);
/* ..._into: every gradient ADDED to its own destination tensor (HOST array of device pointers, same order as the parameters):
 * the parameters' .grad storage, so that autograd's per-tensor accumulation launches disappear (nn/_direct.py). */
int mm_agent_mlp_unfold_into(const float* acc, int32_t n_ctas, const float* scale, float* red, const void* const* params12,
                             void* const* grads12, int32_t D, void* stream);
int mm_agent_mlp_bwd_ctas(int64_t rows);
int mm_agent_mlp_fwd(const float* params, const float* obs, float* gi, float* saved_rows, void* saved_x0, int64_t rows, int32_t D,
                     void* stream);
int mm_agent_mlp_bwd(const float* params, const float* saved_rows, const void* saved_x0, const float* dgi, const float* scale,
                     float* acc, int32_t n_acc_ctas, int64_t rows, void* stream);

/* ------------------------------------------------------------------------------------------------
 * PPO losses of R_MAPPO.ppo_update / cal_value_loss (mappo/algorithms/ramppo_network.py:70-171), one pass each.
 * Rows R = minibatch entries; every per-row input is a float32 vector of length R (logits [R, NA], NA <= 8);
 * active == NULL means weight 1.
 *   actor fwd:  sums4 (double, caller zeroes) += {sum min(imp adv, clamp(imp, 1-clip, 1+clip) adv) w, sum entropy w,
 *               sum imp, sum w};  logp_out (optional) = log-prob of the taken action.
 *   actor bwd:  dlogits = g2[0] * d(sum min w)/dlogits + g2[1] * d(sum entropy w)/dlogits   (g2 on the device).
 *   value fwd:  sums2 += {sum max(loss(ret - v), loss(ret - clipped v)) w, sum w}; ret = (returns - norm2[0]) * norm2[1]
 *               when norm2 != NULL (ValueNorm mean and 1/sqrt(var), device); Huber(delta) or e^2/2.
 *   value bwd:  dvalues = g1[0] * d(sum loss w)/dvalues.
 * first != NULL: every per-row input except logits / values is a TABLE addressed like recurrent_generator's chunks --
 *   row r = l * n_chunks + c reads entry first[c] + l * step_rows (the rollout buffer's [T, N, A, 1] storage as rows), so
 *   no minibatch tensor is gathered.
 * ---------------------------------------------------------------------------------------------- */
int mm_ppo_actor_loss_fwd(const float* logits, const float* actions, const float* old_logp, const float* adv,
                          const float* active, float* logp_out, int64_t R, int32_t NA, float clip, double* sums4,
                          const int64_t* first, int64_t step_rows, int64_t n_chunks, void* stream);
int mm_ppo_actor_loss_bwd(const float* logits, const float* actions, const float* old_logp, const float* adv,
                          const float* active, const float* g2, float* dlogits, int64_t R, int32_t NA, float clip,
                          const int64_t* first, int64_t step_rows, int64_t n_chunks, void* stream);
int mm_ppo_value_loss_fwd(const float* values, const float* value_preds, const float* returns, const float* active,
                          const float* norm2, int64_t R, float clip, float huber_delta, int32_t use_huber, int32_t use_clipped,
                          double* sums2, const int64_t* first, int64_t step_rows, int64_t n_chunks, void* stream);
int mm_ppo_value_loss_bwd(const float* values, const float* value_preds, const float* returns, const float* active,
                          const float* norm2, const float* g1, float* dvalues, int64_t R, float clip, float huber_delta,
                          int32_t use_huber, int32_t use_clipped, const int64_t* first, int64_t step_rows, int64_t n_chunks,
                          void* stream);

/* Column sums of a row-major float32 matrix x [rows, C] (C <= 256) accumulated into out [C] (caller zeroes): the bias
 * gradients of the learner's Linear / GRU layers (what autograd computes with dY.sum(0)). */
int mm_colsum(const float* x, float* out, int64_t rows, int32_t C, void* stream);

/* Gradient-norm clipping + Adam on one flat fp32 parameter buffer (csrc/optim.cu): what the learners run after each backward
 * pass -- torch.nn.utils.clip_grad_norm_(params, max_norm) followed by torch.optim.Adam.step()
 * (offpolicy/algorithms/qmix/qmix.py:200-208, mappo/algorithms/ramppo_network.py:181-209, vdn/_train.py:96-100) -- as two
 * launches.  params / grads / exp_avg / exp_avg_sq [n] (device); step: device float, advanced by one inside the call (Adam's
 * bias-correction exponent); scratch: mm_clip_adam_scratch_doubles() doubles; norm_out (optional): the total gradient 2-norm
 * before clipping (what clip_grad_norm_ returns).  max_norm <= 0: no clipping.  clip_n <= n: the norm is taken over, and the
 * coefficient applied to, the first clip_n elements only (qmix/_train.py:113-115 clips the q-network but steps the mixer too).
 * lr_dev (optional, device float) overrides lr: a captured CUDA graph then follows a learning-rate schedule.  Update rule and operation order of
 * torch.optim.Adam (no amsgrad / maximize): g' = g * min(1, max_norm / (norm + 1e-6)) (+ weight_decay * p),
 * m += (g' - m)(1 - beta1), v = beta2 v + (1 - beta2) g'^2, p -= lr / (1 - beta1^t) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps). */
int mm_clip_adam_scratch_doubles(void);
int mm_clip_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* step, double* scratch,
                      float* norm_out, int64_t n, int64_t clip_n, float lr, const float* lr_dev, float beta1, float beta2,
                      float eps, float weight_decay, float max_norm, void* stream);

/* LayerNorm(H) + the Q head Linear(H, NA) of the recurrent agent network on the GRU outputs, one kernel per direction
 * (csrc/ln_head.cu; rnn.py:18,45 and act.py as agent_q_function.py:60-66 applies them, differentiated by qmix.py:196).
 * H in {32, 64}, NA <= 8.  h [M, H], q [M, NA]; mean / rstd [M] are written by the forward and read by the backward.
 * Backward: dh [M, H]; dwq [NA, H], dbq [NA], dln_w [H], dln_b [H] are ACCUMULATED into (the caller zeroes them). */
int mm_ln_head_fwd(const float* h, const float* ln_w, const float* ln_b, const float* wq, const float* bq, float* q, float* mean,
                   float* rstd, int64_t M, int32_t H, int32_t NA, void* stream);
int mm_ln_head_bwd(const float* dq, const float* h, const float* mean, const float* rstd, const float* ln_w, const float* ln_b,
                   const float* wq, float* dh, float* dwq, float* dbq, float* dln_w, float* dln_b, int64_t M, int32_t H,
                   int32_t NA, void* stream);

/* Recurrent weight gradients of the GRU on tcgen05 (csrc/gru_wgrad.cu): dw_hh [3H, H] = [dgi_r | dgi_z | dgh_n]^T h_prev and
 * db_hh [3H] = the column sums, contracted over all T * R rows -- the weight_hh_l0 / bias_hh_l0 gradients of nn.GRU in RNNLayer
 * (rnn.py:4-47) under qmix.py:196's backward.  dgi [T R, 3H] and dghn [T R, H] from mm_gru_seq_bwd, h_seq [T R, H] from
 * mm_gru_seq_fwd (h_prev of step t is h_seq[t - 1], of step 0 h0 or zero), scale = mm_ppo_trunk2_scale(dgi) (device), acc: scratch
 * of n_acc_ctas >= mm_gru_wgrad_ctas(T R) x 128 x 3H floats; accumulate != 0 adds to dw_hh / db_hh instead of writing.  H = 64. */
int mm_gru_wgrad_ctas(int64_t rows);
int mm_gru_wgrad(const float* dgi, const float* dghn, const float* h_seq, const float* h0, const float* scale, float* acc,
                 int32_t n_acc_ctas, float* dw_hh, float* db_hh, int32_t accumulate, int32_t T, int64_t R, int32_t H, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* MINIMARL_B200_H_ */
