"""Fused policy + env rollouts: T steps of (agent network -> eps-greedy -> Checkers step -> next obs) for all
lockstep envs in ONE kernel launch (``mm_qmix_rollout_tc``).

``qmix_rollout`` is the raw call.  ``EpisodeCollector`` mirrors ``MAGYM_Runner.collect_rollout``
(offpolicy/runner/shared/magym_runner.py:39-141) for N lockstep envs: one call collects N whole episodes
directly into the slots of the device replay buffer (zero-copy insert) and returns the same ``env_info``.
Reference behaviours: the team reward (``use_common_reward``), stepping stops once an episode is done,
``dones`` default to 1.  ``faithful=True`` additionally reproduces quirk Q11 of SURVEY.md Appendix B (the
action INDEX broadcast over the one-hot dim; the final obs written at index T-1 instead of T).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def qmix_rollout(net, env, T, eps, seed=0, step_offset=0, obs=None, act_idx=None, act_onehot=None, rewards=None,
                 dones=None, dones_env=None, q_out=None, h0=None, reset=True, freeze_after_done=True,
                 common_reward=True, eps_per_env=False, obs_t_stride=None, strides=None):
    """net: minimarl_b200.nn.AgentQFunction (H in 32/64, input 47); env: BatchedCheckers (partial obs).
    Output tensors are allocated when not given (natural strides).  Returns a dict of the tensors used."""
    if env.full_observable:
        raise RuntimeError("the fused rollout kernel covers partial observations (D = 47)")
    if not net.tc_ok:
        raise RuntimeError("agent network does not fit the tensor-core kernel")
    dev, N, A, D, NA, H = env.device, env.n_envs, 2, 47, net.act_dim, net.hidden_size
    R = N * A
    z = dict(dtype=torch.float32, device=dev)
    if obs is None:
        obs = torch.zeros((T + 1, N, A, D), **z)
    if act_idx is None:
        act_idx = torch.empty((T, N, A), dtype=torch.int32, device=dev)
    if rewards is None:
        rewards = torch.empty((T, N, A), **z)
    if dones is None:
        dones = torch.empty((T, N, A), **z)
    if dones_env is None:
        dones_env = torch.empty((T, N), **z)
    h_final = torch.empty((R, H), **z)
    st = strides or {}
    io = _lib.RolloutIO()
    io.env_state, io.obs = env.state.data_ptr(), obs.data_ptr()
    io.obs_t_stride = obs_t_stride if obs_t_stride is not None else obs.stride(0)
    io.act_idx, io.act_t_stride = act_idx.data_ptr(), st.get("act_idx", act_idx.stride(0))
    io.act_onehot = act_onehot.data_ptr() if act_onehot is not None else None
    io.onehot_t_stride = st.get("act_onehot", act_onehot.stride(0) if act_onehot is not None else 0)
    io.rewards, io.rew_t_stride = rewards.data_ptr(), st.get("rewards", rewards.stride(0))
    io.dones, io.done_t_stride = dones.data_ptr(), st.get("dones", dones.stride(0))
    io.dones_env, io.denv_t_stride = dones_env.data_ptr(), st.get("dones_env", dones_env.stride(0))
    io.step_offset, io.eps = int(step_offset), float(eps)
    io.eps_per_env, io.do_reset = int(bool(eps_per_env)), int(bool(reset))
    io.freeze_after_done, io.common_reward = int(bool(freeze_after_done)), int(bool(common_reward))
    # float32 of the Python-double team reward for every (food kind of agent 0, of agent 1)
    food = {0: (0, -10, 10), 1: (0, -1, 1)}
    lut = (C.c_float * 9)(*[np.float32((env.step_cost + food[0][k0]) + (env.step_cost + food[1][k1]))
                            for k0 in range(3) for k1 in range(3)])
    if h0 is not None:
        h0 = h0.reshape(R, H).contiguous()
    _lib.check(_lib.lib().mm_qmix_rollout_tc(
        _lib.ptr(net.packed_tc()), _lib.ptr(h0), _lib.ptr(h_final), _lib.ptr(q_out), C.byref(io), C.byref(env.cfg), lut,
        int(T), int(N), int(H), int(NA), int(seed), _lib.stream()))
    return dict(obs=obs, act_idx=act_idx, act_onehot=act_onehot, rewards=rewards, dones=dones, dones_env=dones_env,
                h_final=h_final, q=q_out)


_MAPPO_OFFSETS = None


def _mappo_offsets():
    global _MAPPO_OFFSETS
    if _MAPPO_OFFSETS is None:
        o = (C.c_int32 * 19)()
        _lib.check(_lib.lib().mm_mappo_rollout_param_offsets(o))
        names = ("w1 w2 wih whh b1 ln1_w ln1_b b2 ln2_w ln2_b bih bhh ln3_w ln3_b wa wv ba bv total").split()
        _MAPPO_OFFSETS = dict(zip(names, list(o)))
    return _MAPPO_OFFSETS


def mappo_rollout_fusable(policy, env):
    """True when ``mm_mappo_rollout_tc`` covers this policy / env pair: hidden 32, one MLP layer + one GRU layer with
    feature normalisation and ReLU (the defaults of mappo/_config.py:176-249), local-observation critic, Checkers with
    partial observations."""
    a, c = policy.actor, policy.critic
    return (getattr(a, "_fusable", False) and getattr(c, "_fusable", False) and a.hidden_size == 32 and
            a._in == 47 and c._in == 47 and a._n_out == 5 and a._precision == "tf32" and
            env.__class__.__name__ == "BatchedCheckers" and not env.full_observable and env.n_agents == 2)


def pack_mappo_rollout(actor, critic, out=None):
    """Flatten actor + critic into the ``mm_mappo_rollout_tc`` parameter buffer.  LayerNorm(47)'s affine is folded
    into the first Linear: W1' = W1 diag(gamma), b1' = b1 + W1 beta (r_actor_critic.py / mlp.py feature_norm -> fc1)."""
    o = _mappo_offsets()
    dev = next(actor.parameters()).device
    buf = torch.zeros(o["total"], dtype=torch.float32, device=dev) if out is None else out
    H, K1 = 32, 48
    with torch.no_grad():
        for k, net in enumerate((actor, critic)):
            sd = dict(net.named_parameters())
            g, b = sd["base.feature_norm.weight"], sd["base.feature_norm.bias"]
            W1, b1 = sd["base.mlp.fc1.0.weight"], sd["base.mlp.fc1.0.bias"]
            w1 = buf[o["w1"] + k * H * K1: o["w1"] + (k + 1) * H * K1].view(H, K1)
            w1.zero_()
            w1[:, :47] = W1 * g
            def put(name, val, n):
                buf[o[name] + k * n: o[name] + (k + 1) * n] = val.reshape(-1)
            put("b1", b1 + W1 @ b, H)
            put("ln1_w", sd["base.mlp.fc1.2.weight"], H); put("ln1_b", sd["base.mlp.fc1.2.bias"], H)
            put("w2", sd["base.mlp.fc2.0.0.weight"], H * H); put("b2", sd["base.mlp.fc2.0.0.bias"], H)
            put("ln2_w", sd["base.mlp.fc2.0.2.weight"], H); put("ln2_b", sd["base.mlp.fc2.0.2.bias"], H)
            put("wih", sd["rnn.rnn.weight_ih_l0"], 3 * H * H); put("whh", sd["rnn.rnn.weight_hh_l0"], 3 * H * H)
            put("bih", sd["rnn.rnn.bias_ih_l0"], 3 * H); put("bhh", sd["rnn.rnn.bias_hh_l0"], 3 * H)
            put("ln3_w", sd["rnn.norm.weight"], H); put("ln3_b", sd["rnn.norm.bias"], H)
        sa, sc = dict(actor.named_parameters()), dict(critic.named_parameters())
        buf[o["wa"]: o["wa"] + 5 * H] = sa["act.action_out.linear.weight"].reshape(-1)
        buf[o["ba"]: o["ba"] + 5] = sa["act.action_out.linear.bias"]
        buf[o["wv"]: o["wv"] + H] = sc["v_out.weight"].reshape(-1)
        buf[o["bv"]: o["bv"] + 1] = sc["v_out.bias"]
    return buf


def mappo_rollout(params, env, buffer, T=None, seed=0, step_offset=0, reset=False):
    """T lockstep steps of policy + env written straight into ``buffer`` (a device SharedReplayBuffer whose
    n_rollout_threads == env.n_envs): obs/share_obs/rnn_states/rnn_states_critic/masks [1..T], actions,
    action_log_probs, value_preds, rewards [0..T-1].  Slab 0 is the input (``reset=True`` resets the envs first)."""
    T = buffer.episode_length if T is None else T
    N = env.n_envs
    b = buffer
    assert b.obs.shape[1] == N and b.obs.shape[0] >= T + 1 and b.obs.shape[-1] == 47 and b.share_obs.shape[-1] == 47
    for x in (b.obs, b.share_obs, b.rnn_states, b.rnn_states_critic, b.actions, b.action_log_probs, b.value_preds,
              b.rewards, b.masks):
        assert x.is_contiguous() and x.dtype == torch.float32
    _lib.check(_lib.lib().mm_mappo_rollout_tc(
        _lib.ptr(params), _lib.ptr(env.state), _lib.ptr(b.obs), _lib.ptr(b.share_obs), _lib.ptr(b.rnn_states),
        _lib.ptr(b.rnn_states_critic), _lib.ptr(b.actions), _lib.ptr(b.action_log_probs), _lib.ptr(b.value_preds),
        _lib.ptr(b.rewards), _lib.ptr(b.masks), int(T), int(N), int(bool(reset)), int(seed), int(step_offset),
        C.byref(env.cfg), _lib.stream()))


class EpisodeCollector:
    """N lockstep envs -> N episodes per call, written straight into a PrioritizedRecReplayBuffer / RecReplayBuffer."""

    def __init__(self, env, policy, buffer, episode_length, seed=0, faithful=False):
        self.env, self.policy, self.buffer, self.T = env, policy, buffer, episode_length
        self.seed, self.faithful = seed, faithful
        self.total_env_steps = 0
        self.num_episodes_collected = 0

    @torch.no_grad()
    def collect_rollout(self, explore=True, training_episode=True, warmup=False):
        env, T, N = self.env, self.T, self.env.n_envs
        pb = self.buffer.policy_buffers["policy_0"]
        S, A, D, NA = pb.buffer_size, pb.num_agents, pb.obs_dim, pb.act_dim
        slots = pb.next_slots(N)
        if not (slots[-1] - slots[0] == N - 1):
            raise RuntimeError("buffer_size must be a multiple of the number of lockstep envs (ring wrap inside a batch)")
        i0 = int(slots[0])
        if warmup:
            eps = 1.0                                  # random actions (base_runner.py:225-238)
        elif explore:
            eps = self.policy.exploration.eval(self.total_env_steps)
        else:
            eps = 0.0
        sl = slice(i0, i0 + N)
        if not (explore or warmup):
            # evaluation episodes are never committed: they go to scratch tensors, not into slots that may hold live
            # training episodes (whose priorities stay in the tree)
            return self._eval_rollout(eps, training_episode)
        onehot = pb.acts[:, sl]
        out = qmix_rollout(
            self.policy.q_network, env, T, eps, seed=self.seed, step_offset=self.total_env_steps,
            obs=pb.obs[:, sl], act_onehot=onehot, rewards=pb.rewards[:, sl, :, 0], dones=pb.dones[:, sl, :, 0],
            dones_env=pb.dones_env[:, sl, 0], obs_t_stride=pb.obs.stride(0),
            strides=dict(act_onehot=pb.acts.stride(0), rewards=pb.rewards.stride(0), dones=pb.dones.stride(0),
                         dones_env=pb.dones_env.stride(0)))
        pb.share_obs[:, sl] = pb.obs[:, sl].reshape(T + 1, N, A * D)       # obs_sharing: concat of all agents' obs
        if self.faithful:
            pb.acts[:, sl] = out["act_idx"].to(torch.float32).unsqueeze(-1).expand(T, N, A, NA)
            pb.obs[T - 1, sl] = pb.obs[T, sl]
            pb.obs[T, sl] = 0
            pb.share_obs[T - 1, sl] = pb.share_obs[T, sl]
            pb.share_obs[T, sl] = 0
        if training_episode or warmup:
            self.total_env_steps += N * T
        if explore or warmup:
            self.num_episodes_collected += N
            if hasattr(self.buffer, "commit"):
                self.buffer.commit(slots)
            else:
                pb.commit(slots)
        # the reference logs the team return of the (single) env; here: mean over the N envs, a device scalar
        return {"average_episode_rewards": pb.rewards[:, sl, 0, 0].sum(0).mean()}

    def _eval_rollout(self, eps, training_episode):
        env, T, N = self.env, self.T, self.env.n_envs
        pb = self.buffer.policy_buffers["policy_0"]
        A, D, NA = pb.num_agents, pb.obs_dim, pb.act_dim
        sc = getattr(self, "_scratch", None)
        if sc is None:
            z = dict(dtype=torch.float32, device=pb.obs.device)
            sc = self._scratch = dict(obs=torch.zeros((T + 1, N, A, D), **z), act_onehot=torch.zeros((T, N, A, NA), **z),
                                      rewards=torch.zeros((T, N, A), **z), dones=torch.zeros((T, N, A), **z),
                                      dones_env=torch.zeros((T, N), **z))
        qmix_rollout(self.policy.q_network, env, T, eps, seed=self.seed, step_offset=self.total_env_steps, **sc)
        if training_episode:
            self.total_env_steps += N * T
        return {"average_episode_rewards": sc["rewards"][:, :, 0].sum(0).mean()}
