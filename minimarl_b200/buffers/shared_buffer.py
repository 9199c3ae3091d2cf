"""Device-resident MAPPO rollout buffer with the reference's interface
(mappo/runner/shared/shared_buffer.py:15-427).

Same field names and shapes (``obs [T+1, N, A, D]``, ``rnn_states [T+1, N, A, 1, H]``, ``value_preds /
returns / masks [T+1, N, A, 1]``, ``actions / action_log_probs / rewards [T, N, A, 1]``) but as CUDA
tensors, so the env kernel writes ``obs[t+1]`` / ``rewards[t]`` in place (``BatchedCheckers.step_into``)
and ``compute_returns`` is one launch of the GAE scan (``mm_gae_returns``) instead of a Python loop
over T with numpy temporaries.  The minibatch generators index the time-major storage directly on the
device; the tuples they yield have the reference's order and shapes.
"""
import numpy as np
import torch

from ..ops.gae import gae_returns


def _shape_from_space(space):
    name = space.__class__.__name__
    if name == "Box":
        return tuple(space.shape)
    if name == "list":
        return tuple(space)
    raise NotImplementedError(name)


def _to_dev(x, device):
    if torch.is_tensor(x):
        return x.to(device=device, dtype=torch.float32)
    return torch.as_tensor(np.asarray(x), dtype=torch.float32, device=device)


class SharedReplayBuffer(object):
    def __init__(self, args, num_agents, obs_space, cent_obs_space, act_space, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("minimarl_b200 buffers live on a CUDA device (no CPU fallback)")
        self._use_gae = args.use_gae
        self._use_popart = args.use_popart
        self._use_valuenorm = args.use_valuenorm
        self.num_agents = num_agents
        self.hidden_size = args.hidden_size
        self.recurrent_N = args.recurrent_N
        self.episode_length = args.max_step
        self.batch_size = args.batch_size
        self.gamma = args.gamma
        self.gae_lambda = args.gae_lambda

        obs_shape = _shape_from_space(obs_space)
        share_obs_shape = _shape_from_space(cent_obs_space)
        T, N, A = self.episode_length, self.batch_size, num_agents
        z = dict(dtype=torch.float32, device=self.device)
        self.share_obs = torch.zeros((T + 1, N, A, *share_obs_shape), **z)
        self.obs = torch.zeros((T + 1, N, A, *obs_shape), **z)
        self.rnn_states = torch.zeros((T + 1, N, A, self.recurrent_N, self.hidden_size), **z)
        self.rnn_states_critic = torch.zeros_like(self.rnn_states)
        self.value_preds = torch.zeros((T + 1, N, A, 1), **z)
        self.returns = torch.zeros((T + 1, N, A, 1), **z)
        self.advantages = torch.zeros((T, N, A, 1), **z)      # filled by compute_returns (same launch)
        self.available_actions = None
        if act_space.__class__.__name__ != "Discrete":
            raise NotImplementedError("ma-gym action spaces are Discrete")
        act_shape = 1
        self.actions = torch.zeros((T, N, A, act_shape), **z)
        self.action_log_probs = torch.zeros((T, N, A, act_shape), **z)
        self.rewards = torch.zeros((T, N, A, 1), **z)
        self.masks = torch.ones((T + 1, N, A, 1), **z)
        self.bad_masks = torch.ones_like(self.masks)
        self.active_masks = torch.ones_like(self.masks)
        self.step = 0

    def insert(self, share_obs, obs, rnn_states_actor, rnn_states_critic, actions, action_log_probs,
               value_preds, rewards, masks, bad_masks=None, active_masks=None, available_actions=None):
        """Same argument list as the reference.  A tensor that already IS the destination slice (the env
        kernel wrote ``obs[step+1]`` / ``rewards[step]`` in place) is skipped instead of copied."""
        d, s = self.device, self.step

        def put(dst, src):
            if src is None:
                return
            if torch.is_tensor(src) and src.data_ptr() == dst.data_ptr() and src.device == dst.device:
                return
            dst.copy_(_to_dev(src, d).reshape(dst.shape))

        put(self.share_obs[s + 1], share_obs)
        put(self.obs[s + 1], obs)
        put(self.rnn_states[s + 1], rnn_states_actor)
        put(self.rnn_states_critic[s + 1], rnn_states_critic)
        put(self.actions[s], actions)
        put(self.action_log_probs[s], action_log_probs)
        put(self.value_preds[s], value_preds)
        put(self.rewards[s], rewards)
        put(self.masks[s + 1], masks)
        put(self.bad_masks[s + 1], bad_masks)
        put(self.active_masks[s + 1], active_masks)
        self.step = (self.step + 1) % self.episode_length

    def after_update(self):
        """Copy last timestep data to first index. Called after update to model."""
        for x in (self.share_obs, self.obs, self.rnn_states, self.rnn_states_critic, self.masks, self.bad_masks,
                  self.active_masks):
            x[0].copy_(x[-1])

    def compute_returns(self, next_value, value_normalizer=None):
        """GAE (or discounted returns) over the whole buffer in one kernel launch; also fills
        ``self.advantages`` = returns[:-1] - denormalize(value_preds[:-1]) when GAE is on."""
        denorm = None
        if self._use_popart or self._use_valuenorm:
            mean, var = value_normalizer.running_mean_var()
            denorm = (float(torch.sqrt(var).reshape(-1)[0]), float(mean.reshape(-1)[0]))
        if self._use_gae:
            self.value_preds[-1].copy_(_to_dev(next_value, self.device).reshape(self.value_preds[-1].shape))
            gae_returns(self.rewards, self.value_preds, self.masks, self.returns, self.gamma, self.gae_lambda,
                        use_gae=True, denorm=denorm, advantages=self.advantages)
        else:
            self.returns[-1].copy_(_to_dev(next_value, self.device).reshape(self.returns[-1].shape))
            gae_returns(self.rewards, None, self.masks, self.returns, self.gamma, self.gae_lambda, use_gae=False)

    # -- minibatch generators ---------------------------------------------------------------------
    def _perm(self, n, perm):
        if perm is not None:
            return torch.as_tensor(np.asarray(perm) if not torch.is_tensor(perm) else perm, device=self.device).long()
        return torch.randperm(n, device=self.device)

    def feed_forward_generator(self, advantages, num_mini_batch=None, mini_batch_size=None, perm=None):
        T, N, A = self.rewards.shape[0:3]
        batch_size = N * T * A
        if mini_batch_size is None:
            assert batch_size >= num_mini_batch
            mini_batch_size = batch_size // num_mini_batch
        rand = self._perm(batch_size, perm)
        flat = lambda x: x.reshape(-1, *x.shape[3:])
        share_obs, obs = flat(self.share_obs[:-1]), flat(self.obs[:-1])
        rnn_states, rnn_states_critic = flat(self.rnn_states[:-1]), flat(self.rnn_states_critic[:-1])
        actions, log_probs = flat(self.actions), flat(self.action_log_probs)
        value_preds, returns = flat(self.value_preds[:-1]), flat(self.returns[:-1])
        masks, active_masks = flat(self.masks[:-1]), flat(self.active_masks[:-1])
        adv = _to_dev(advantages, self.device).reshape(-1, 1)
        for i in range(num_mini_batch):
            ind = rand[i * mini_batch_size:(i + 1) * mini_batch_size]
            yield (share_obs[ind], obs[ind], rnn_states[ind], rnn_states_critic[ind], actions[ind],
                   value_preds[ind], returns[ind], masks[ind], active_masks[ind], log_probs[ind], adv[ind], None)

    def recurrent_generator(self, advantages, num_mini_batch, data_chunk_length, perm=None, lazy_obs=False, tables=None):
        """Chunks of L consecutive rows of the (env, agent, time)-major flattening, shuffled
        (shared_buffer.py:318-427).  Row f of that flattening is storage element [t, env, agent] with
        env = f // (A*T), agent = (f // T) % A, t = f % T, so the chunks are gathered straight from the
        time-major tensors without materialising the transposed copy."""
        T, N, A = self.rewards.shape[0:3]
        L = data_chunk_length
        batch_size = N * T * A
        data_chunks = batch_size // L
        mini_batch_size = data_chunks // num_mini_batch
        rand = self._perm(data_chunks, perm)
        adv = _to_dev(advantages, self.device).reshape(T, N, A, -1)
        steps = torch.arange(L, device=self.device)
        for i in range(num_mini_batch):
            chunk = rand[i * mini_batch_size:(i + 1) * mini_batch_size]
            f0 = chunk * L
            env0, agent0, t0 = f0 // (A * T), (f0 // T) % A, f0 % T
            if lazy_obs and T % L == 0:
                # the fused update reads every chunk in place: one row map (first row of each chunk in the [rows, F] storage
                # of a [T(+1), N, A, F] tensor, N * A rows per step) serves all eleven fields, nothing is gathered
                from ..nn.ppo_trunk import ChunkRows
                first = (t0 * N + env0) * A + agent0
                rows = lambda x, key=None: ChunkRows(x.reshape(-1, x.shape[-1]), first, N * A, L, tables, key)
                state = lambda x: ChunkRows(x.reshape(-1, *x.shape[3:]), first, N * A, 1)
                # ``tables`` (a dict owned by the caller, one per update) keeps the normalised fp16 observation tables
                yield (rows(self.share_obs, "share_obs"), rows(self.obs, "obs"), state(self.rnn_states), state(self.rnn_states_critic),
                       rows(self.actions), rows(self.value_preds), rows(self.returns), rows(self.masks),
                       rows(self.active_masks), rows(self.action_log_probs), rows(adv), None)
                continue
            f = f0.unsqueeze(0) + steps.unsqueeze(1)                        # [L, Nc] flat (env, agent, t) rows
            env, agent, t = f // (A * T), (f // T) % A, f % T
            take = lambda x: x[t, env, agent].reshape(L * mini_batch_size, *x.shape[3:])
            yield (take(self.share_obs), take(self.obs), self.rnn_states[t0, env0, agent0],
                   self.rnn_states_critic[t0, env0, agent0], take(self.actions), take(self.value_preds),
                   take(self.returns), take(self.masks), take(self.active_masks), take(self.action_log_probs),
                   take(adv), None)

    def naive_recurrent_generator(self, advantages, num_mini_batch, perm=None):
        T, N, A = self.rewards.shape[0:3]
        batch_size = N * A
        assert batch_size >= num_mini_batch
        per = batch_size // num_mini_batch
        rand = self._perm(batch_size, perm)
        col = lambda x: x.reshape(x.shape[0], batch_size, *x.shape[3:])
        adv = col(_to_dev(advantages, self.device).reshape(T, N, A, -1))
        for start in range(0, batch_size, per):
            ind = rand[start:start + per]
            seq = lambda x, upto: col(x)[:upto, ind].reshape(upto * ind.numel(), *x.shape[3:])
            yield (seq(self.share_obs, T), seq(self.obs, T), col(self.rnn_states)[0, ind],
                   col(self.rnn_states_critic)[0, ind], seq(self.actions, T), seq(self.value_preds, T),
                   seq(self.returns, T), seq(self.masks, T), seq(self.active_masks, T),
                   seq(self.action_log_probs, T), adv[:, ind].reshape(T * ind.numel(), -1), None)
