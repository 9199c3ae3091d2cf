"""Device-resident whole-episode replay with the reference's interface (offpolicy/utils/rec_buffer.py).

Storage keeps the reference's time-major layout (``obs [T+1, S, A, D]`` ...) as CUDA tensors, so a
lockstep rollout can write episodes straight into its slots; ``sample`` runs the prioritised index
draw (``mm_per_sample``) and the batch gather (``mm_replay_gather``) on the device and returns the
same 9-tuple of per-policy dicts, holding CUDA tensors instead of numpy arrays.

Reference behaviours kept on purpose: ``dones`` / ``dones_env`` default to 1;
``PrioritizedRecReplayBuffer.insert`` seeds the priority of leaves ``0..n-1`` rather than of the slots it
just filled (Q12, ``fix_insert_priority=True`` turns that off); the sampling mass is scaled by
``sum(0, len - 1)``.
"""
import numpy as np
import torch

from .. import _lib
from .segment_tree import MinSegmentTree, SumSegmentTree


def _space_dim(space):
    name = space.__class__.__name__
    if name == "Box":
        return int(space.shape[0])
    if name == "Discrete":
        return int(space.n)
    if name == "list":
        return int(space[0])
    raise NotImplementedError(name)


def _f32(x, device):
    if torch.is_tensor(x):
        return x.to(device=device, dtype=torch.float32)
    return torch.as_tensor(np.asarray(x), dtype=torch.float32, device=device)


class RecPolicyBuffer:
    def __init__(self, buffer_size, episode_length, num_agents, obs_space, share_obs_space, act_space,
                 use_same_share_obs, use_avail_acts, use_reward_normalization=False, device="cuda"):
        if not use_same_share_obs or use_avail_acts or use_reward_normalization:
            raise NotImplementedError("device replay covers the reference's defaults: use_same_share_obs=True, "
                                      "use_avail_acts=False, use_reward_normalization=False")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("minimarl_b200 replay buffers live on a CUDA device (no CPU fallback)")
        self.buffer_size, self.episode_length, self.num_agents = buffer_size, episode_length, num_agents
        self.use_same_share_obs, self.use_avail_acts = use_same_share_obs, use_avail_acts
        self.use_reward_normalization = use_reward_normalization
        self.filled_i = 0
        self.current_i = 0
        T, S, A = episode_length, buffer_size, num_agents
        self.obs_dim, self.share_obs_dim, self.act_dim = _space_dim(obs_space), _space_dim(share_obs_space), _space_dim(act_space)
        z = dict(dtype=torch.float32, device=self.device)
        self.obs = torch.zeros((T + 1, S, A, self.obs_dim), **z)
        self.share_obs = torch.zeros((T + 1, S, self.share_obs_dim), **z)
        self.acts = torch.zeros((T, S, A, self.act_dim), **z)
        self.rewards = torch.zeros((T, S, A, 1), **z)
        self.dones = torch.ones((T, S, A, 1), **z)          # default to done being True
        self.dones_env = torch.ones((T, S, 1), **z)

    def __len__(self):
        return self.filled_i

    def next_slots(self, num_insert_episodes):
        if self.current_i + num_insert_episodes <= self.buffer_size:
            return np.arange(self.current_i, self.current_i + num_insert_episodes)
        left = self.current_i + num_insert_episodes - self.buffer_size
        return np.concatenate((np.arange(self.current_i, self.buffer_size), np.arange(left)))

    def commit(self, idx_range):
        """Bookkeeping after episodes were written into ``idx_range`` (also used by zero-copy rollouts)."""
        self.current_i = int(idx_range[-1]) + 1
        self.filled_i = min(self.filled_i + len(idx_range), self.buffer_size)

    def insert(self, num_insert_episodes, obs, share_obs, acts, rewards, dones, dones_env, avail_acts=None):
        episode_length = acts.shape[0]
        assert episode_length == self.episode_length, ("different dimension!")
        idx_range = self.next_slots(num_insert_episodes)
        dev = self.device
        share_obs = _f32(share_obs, dev)[:, :, 0]          # all agents share the centralized observation
        if idx_range[-1] - idx_range[0] == len(idx_range) - 1:
            sl = slice(int(idx_range[0]), int(idx_range[-1]) + 1)
        else:
            sl = torch.as_tensor(idx_range, device=dev)
        self.obs[:, sl] = _f32(obs, dev)
        self.share_obs[:, sl] = share_obs
        self.acts[:, sl] = _f32(acts, dev)
        self.rewards[:, sl] = _f32(rewards, dev)
        self.dones[:, sl] = _f32(dones, dev)
        self.dones_env[:, sl] = _f32(dones_env, dev)
        self.commit(idx_range)
        return idx_range

    def sample_inds(self, sample_inds):
        """Gather the episodes ``sample_inds`` (int64 CUDA tensor or array): one kernel, agent-major outputs."""
        idx = torch.as_tensor(sample_inds, device=self.device).to(torch.int64).contiguous() \
            if not torch.is_tensor(sample_inds) else sample_inds.to(device=self.device, dtype=torch.int64).contiguous()
        B, T, A = idx.numel(), self.episode_length, self.num_agents
        z = dict(dtype=torch.float32, device=self.device)
        # obs / acts are STORED [T(+1), A, B, d] -- the (t, a * B + b) row order the learner stacks the agents into, so its
        # torch.cat over the agents is a view -- and returned as the reference's [A, T(+1), B, d] through a permuted view
        obs = torch.empty((T + 1, A, B, self.obs_dim), **z)
        share_obs = torch.empty((T + 1, B, self.share_obs_dim), **z)
        acts = torch.empty((T, A, B, self.act_dim), **z)
        rewards = torch.empty((A, T, B, 1), **z)
        dones = torch.empty((A, T, B, 1), **z)
        dones_env = torch.empty((T, B, 1), **z)
        p = _lib.ptr
        _lib.check(_lib.lib().mm_replay_gather_stacked(
            p(self.obs), p(self.share_obs), p(self.acts), p(self.rewards), p(self.dones), p(self.dones_env),
            p(idx), p(obs), p(share_obs), p(acts), p(rewards), p(dones), p(dones_env),
            T, self.buffer_size, B, A, self.obs_dim, self.share_obs_dim, self.act_dim, _lib.stream()))
        return obs.permute(1, 0, 2, 3), share_obs, acts.permute(1, 0, 2, 3), rewards, dones, dones_env, None


class RecReplayBuffer:
    def __init__(self, policy_info, policy_agents, buffer_size, episode_length, use_same_share_obs, use_avail_acts,
                 use_reward_normalization=False, device="cuda"):
        self.policy_info = policy_info
        self.device = torch.device(device)
        self.policy_buffers = {
            p_id: RecPolicyBuffer(buffer_size, episode_length, len(policy_agents[p_id]),
                                  policy_info[p_id]["obs_space"], policy_info[p_id]["share_obs_space"],
                                  policy_info[p_id]["act_space"], use_same_share_obs, use_avail_acts,
                                  use_reward_normalization, device=device)
            for p_id in policy_info.keys()}
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(0)

    def seed(self, seed):
        self._gen.manual_seed(int(seed))

    def __len__(self):
        return self.policy_buffers["policy_0"].filled_i

    def insert(self, num_insert_episodes, obs, share_obs, acts, rewards, dones, dones_env, avail_acts=None):
        for p_id in self.policy_info.keys():
            idx_range = self.policy_buffers[p_id].insert(num_insert_episodes, obs[p_id], share_obs[p_id], acts[p_id],
                                                         rewards[p_id], dones[p_id], dones_env[p_id])
        return idx_range

    def _gather(self, inds):
        obs, share_obs, acts, rewards, dones, dones_env, avail_acts = {}, {}, {}, {}, {}, {}, {}
        for p_id in self.policy_info.keys():
            (obs[p_id], share_obs[p_id], acts[p_id], rewards[p_id], dones[p_id], dones_env[p_id],
             avail_acts[p_id]) = self.policy_buffers[p_id].sample_inds(inds)
        return obs, share_obs, acts, rewards, dones, dones_env, avail_acts

    def sample(self, batch_size):
        inds = torch.randint(0, len(self), (batch_size,), device=self.device, generator=self._gen)
        return self._gather(inds) + (None, None)


class PrioritizedRecReplayBuffer(RecReplayBuffer):
    def __init__(self, alpha, policy_info, policy_agents, buffer_size, episode_length, use_same_share_obs,
                 use_avail_acts, use_reward_normalization=False, device="cuda", fix_insert_priority=False):
        super().__init__(policy_info, policy_agents, buffer_size, episode_length, use_same_share_obs,
                         use_avail_acts, use_reward_normalization, device=device)
        self.alpha = alpha
        self.fix_insert_priority = fix_insert_priority
        it_capacity = 1
        while it_capacity < buffer_size:
            it_capacity *= 2
        self._it_sums = {p_id: SumSegmentTree(it_capacity, device) for p_id in policy_info.keys()}
        self._it_mins = {p_id: MinSegmentTree(it_capacity, device) for p_id in policy_info.keys()}
        self.max_priorities = {p_id: 1.0 for p_id in policy_info.keys()}
        self._max_prio_dev = {}          # running maxima of un-validated updates (device scalars), folded in at the next insert
        # sharded replay (one shard per GPU): importance weights normalised with min p / sum p / len over ALL
        # shards (parallel.per_global_weights) instead of this shard's tree alone
        self.sharded = False

    def _seed_priorities(self, idx_range):
        n = len(idx_range)
        for p_id in self.policy_info.keys():
            if p_id in self._max_prio_dev:
                self.max_priorities[p_id] = max(self.max_priorities[p_id], float(self._max_prio_dev[p_id]))
            v = self.max_priorities[p_id] ** self.alpha
            st, mt = self._it_sums[p_id], self._it_mins[p_id]
            if self.fix_insert_priority:
                idx = torch.as_tensor(idx_range, dtype=torch.int64, device=self.device)
                iptr, base = _lib.ptr(idx), 0
            else:                                            # reference: leaves 0..n-1 (rec_buffer.py:265-268)
                iptr, base = None, 0
            _lib.check(_lib.lib().mm_segtree_set(_lib.ptr(st._value), _lib.ptr(mt._value), st._capacity, iptr, None,
                                                 float(v), base, n, _lib.stream()))

    def insert(self, num_insert_episodes, obs, share_obs, acts, rewards, dones, dones_env, avail_acts=None):
        idx_range = super().insert(num_insert_episodes, obs, share_obs, acts, rewards, dones, dones_env, avail_acts)
        self._seed_priorities(idx_range)
        return idx_range

    def commit(self, idx_range, p_id="policy_0"):
        """Zero-copy insert: the rollout already wrote the episodes into ``idx_range``."""
        self.policy_buffers[p_id].commit(idx_range)
        self._seed_priorities(idx_range)

    def sample(self, batch_size, beta=0, p_id=None, uniforms=None, beta_dev=None):
        """Returns the reference's 9-tuple; ``weights`` (float64) and ``batch_inds`` (int64) are CUDA tensors.
        ``uniforms`` (optional, float64 in [0,1)) replaces the internal draw, for reproducible tests.  ``beta_dev`` (optional,
        0-dim float64 CUDA tensor) is read by the kernel instead of ``beta``: a captured update follows the beta schedule."""
        assert len(self) > batch_size, "Cannot sample with no completed episodes in the buffer!"
        assert beta > 0 or beta_dev is not None
        if uniforms is None:
            uniforms = torch.rand(batch_size, dtype=torch.float64, device=self.device, generator=self._gen)
        else:
            uniforms = torch.as_tensor(np.asarray(uniforms) if not torch.is_tensor(uniforms) else uniforms,
                                       dtype=torch.float64, device=self.device).contiguous()
        st, mt = self._it_sums[p_id], self._it_mins[p_id]
        batch_inds = torch.empty(batch_size, dtype=torch.int64, device=self.device)
        weights = torch.empty(batch_size, dtype=torch.float64, device=self.device)
        if beta_dev is not None and not self.sharded:
            _lib.check(_lib.lib().mm_per_sample_dev(_lib.ptr(st._value), _lib.ptr(mt._value), st._capacity, len(self),
                                                    _lib.ptr(uniforms), batch_size, _lib.ptr(beta_dev), _lib.ptr(batch_inds),
                                                    _lib.ptr(weights), _lib.stream()))
        else:
            _lib.check(_lib.lib().mm_per_sample(_lib.ptr(st._value), _lib.ptr(mt._value), st._capacity, len(self),
                                                _lib.ptr(uniforms), batch_size, float(beta), _lib.ptr(batch_inds),
                                                _lib.ptr(weights), _lib.stream()))
        if self.sharded:
            from ..parallel import per_global_weights
            n = len(self)
            leaves = st._value[st._capacity:st._capacity + n]
            weights = per_global_weights(leaves[batch_inds], mt._value[st._capacity:st._capacity + n].min(), leaves.sum(), n, beta)
        return self._gather(batch_inds) + (weights, batch_inds)

    def update_priorities(self, idxes, priorities, p_id=None, validate=True):
        """``priorities ** alpha`` goes into both trees in one launch.  Host (numpy) priorities -- what the
        reference's trainer hands over -- are raised to alpha on the host in float64 exactly as the reference
        does; device tensors are raised on the device (torch.pow, <= 1 ulp from libm)."""
        assert len(idxes) == len(priorities)
        if torch.is_tensor(priorities):
            pr = priorities.to(device=self.device, dtype=torch.float64).reshape(-1)
            vals = torch.pow(pr, self.alpha).contiguous()
        else:
            host = np.asarray(priorities, dtype=np.float64).reshape(-1)
            pr = torch.as_tensor(host, device=self.device)
            vals = torch.as_tensor(host ** self.alpha, device=self.device).contiguous()
        idx = torch.as_tensor(np.asarray(idxes) if not torch.is_tensor(idxes) else idxes,
                              device=self.device).to(torch.int64).reshape(-1).contiguous()
        if validate:
            # the reference's asserts (rec_buffer.py:315-318) and max tracking need the values on the host: one small sync, BEFORE
            # anything is written, so an out-of-range index can never reach the tree
            stats = torch.stack([pr.min(), pr.max(), idx.min().double(), idx.max().double()]).tolist()
            assert stats[0] > 0
            assert stats[2] >= 0
            assert stats[3] < len(self)
        else:
            # no host round trip (CUDA-graph replay, GraphedUpdate): indices are clamped instead of asserted, the running maximum
            # stays on the device until the next insert reads it
            idx = idx.clamp(0, len(self) - 1)
            m = self._max_prio_dev.get(p_id)
            if m is None:
                m = self._max_prio_dev[p_id] = torch.zeros((), dtype=torch.float64, device=self.device)
            m.copy_(torch.maximum(m, pr.max()))
        # PER samples with replacement: when an index repeats the reference's loop lets the LAST write win.  Every duplicate is
        # given the value of its last occurrence, so the parallel write is deterministic.
        order = torch.arange(idx.numel(), device=self.device)
        last = torch.full((len(self),), -1, dtype=torch.int64, device=self.device).scatter_reduce_(0, idx, order, "amax")
        vals = vals[last[idx]].contiguous()
        st, mt = self._it_sums[p_id], self._it_mins[p_id]
        _lib.check(_lib.lib().mm_segtree_set(_lib.ptr(st._value), _lib.ptr(mt._value), st._capacity, _lib.ptr(idx),
                                             _lib.ptr(vals), 0.0, 0, idx.numel(), _lib.stream()))
        if validate:
            self.max_priorities[p_id] = max(self.max_priorities[p_id], stats[1])
