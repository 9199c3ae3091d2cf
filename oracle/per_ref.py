"""numpy restatement of the off-policy episode replay: segment trees, prioritised sampling, gather.
TEST INFRASTRUCTURE -- never imported by the product.

Follows ``offpolicy/utils/segment_tree.py`` (``SegmentTree.__setitem__`` :74-89, ``_reduce_helper``
:40-56, ``SumSegmentTree.find_prefixsum_idx`` :115-146) and ``offpolicy/utils/rec_buffer.py``
(``RecPolicyBuffer.insert`` :146-190, ``sample_inds`` :192-240, ``PrioritizedRecReplayBuffer``
:243-324), including the behaviours SURVEY.md Appendix B lists as quirks: ``insert`` writes the new
max-priority into leaves ``0..n-1`` rather than into the slots just filled (Q12), and the sampling mass
is scaled by ``sum(0, len - 1)``, a range that leaves out the last two filled leaves' upper end
(``reduce`` treats ``end`` as exclusive and the caller already subtracted one).

Pinned: ``tests/golden/per.npz`` was produced by the reference classes themselves
(``tests/golden/make_golden.py``); ``tests/test_oracle_per.py`` replays the same operations here and
compares trees, indices and gathered batches bit for bit (weights to 1e-12: they go through ``pow``).
"""
import numpy as np


class SegTreeRef:
    """Array-backed binary tree over ``capacity`` (a power of two) float64 leaves; node 1 is the root."""

    def __init__(self, capacity, kind):
        assert capacity > 0 and capacity & (capacity - 1) == 0
        self.capacity = capacity
        self.kind = kind
        self.neutral = 0.0 if kind == "sum" else float("inf")
        self.value = np.full(2 * capacity, self.neutral, dtype=np.float64)

    def _op(self, a, b):
        return a + b if self.kind == "sum" else np.minimum(a, b)

    def set(self, idx, val):
        nodes = np.atleast_1d(np.asarray(idx, dtype=np.int64)) + self.capacity
        self.value[nodes] = val
        nodes = np.unique(nodes // 2)
        while nodes[0] >= 1:                       # level by level up to (and including) the root
            self.value[nodes] = self._op(self.value[2 * nodes], self.value[2 * nodes + 1])
            if nodes[0] == 1:
                break
            nodes = np.unique(nodes // 2)

    def get(self, idx):
        return self.value[self.capacity + np.asarray(idx)]

    def reduce(self, start=0, end=None):
        """op over leaves [start, end) -- same association order as the reference's recursion."""
        if end is None:
            end = self.capacity
        if end < 0:
            end += self.capacity
        return self._range(start, end - 1, 1, 0, self.capacity - 1)

    def _range(self, lo, hi, node, node_lo, node_hi):
        if lo == node_lo and hi == node_hi:
            return self.value[node]
        mid = (node_lo + node_hi) // 2
        if hi <= mid:
            return self._range(lo, hi, 2 * node, node_lo, mid)
        if lo > mid:
            return self._range(lo, hi, 2 * node + 1, mid + 1, node_hi)
        return self._op(self._range(lo, mid, 2 * node, node_lo, mid),
                        self._range(mid + 1, hi, 2 * node + 1, mid + 1, node_hi))

    def find_prefixsum(self, mass):
        """For each mass: descend from the root, going right (and subtracting) when left <= mass."""
        mass = np.array(mass, dtype=np.float64)
        node = np.ones(len(mass), dtype=np.int64)
        while node[0] < self.capacity:             # all walkers are on the same level
            left = 2 * node
            go_right = self.value[left] <= mass
            mass = np.where(go_right, mass - self.value[left], mass)
            node = np.where(go_right, left + 1, left)
        return node - self.capacity


class EpisodeReplayRef:
    """Ring buffer of whole episodes, time-major, plus the PER trees (one policy)."""

    def __init__(self, alpha, buffer_size, episode_length, num_agents, obs_dim, share_obs_dim, act_dim):
        T, S, A = episode_length, buffer_size, num_agents
        self.alpha, self.size, self.T, self.A = alpha, S, T, A
        self.obs = np.zeros((T + 1, S, A, obs_dim), dtype=np.float32)
        self.share_obs = np.zeros((T + 1, S, share_obs_dim), dtype=np.float32)
        self.acts = np.zeros((T, S, A, act_dim), dtype=np.float32)
        self.rewards = np.zeros((T, S, A, 1), dtype=np.float32)
        self.dones = np.ones((T, S, A, 1), dtype=np.float32)
        self.dones_env = np.ones((T, S, 1), dtype=np.float32)
        self.filled, self.cursor = 0, 0
        cap = 1
        while cap < S:
            cap *= 2
        self.sum_tree, self.min_tree = SegTreeRef(cap, "sum"), SegTreeRef(cap, "min")
        self.max_priority = 1.0

    def __len__(self):
        return self.filled

    def insert(self, n, obs, share_obs, acts, rewards, dones, dones_env):
        if self.cursor + n <= self.size:
            slots = np.arange(self.cursor, self.cursor + n)
        else:
            slots = np.concatenate((np.arange(self.cursor, self.size), np.arange(self.cursor + n - self.size)))
        self.obs[:, slots] = obs
        self.share_obs[:, slots] = share_obs[:, :, 0]       # use_same_share_obs: agent 0's copy
        self.acts[:, slots] = acts
        self.rewards[:, slots] = rewards
        self.dones[:, slots] = dones
        self.dones_env[:, slots] = dones_env
        self.cursor = slots[-1] + 1
        self.filled = min(self.filled + n, self.size)
        for leaf in range(n):                               # Q12: leaves 0..n-1, not `slots`
            self.sum_tree.set(leaf, self.max_priority ** self.alpha)
            self.min_tree.set(leaf, self.max_priority ** self.alpha)
        return slots

    def gather(self, inds):
        to_agent_major = lambda x: x[:, inds].transpose(2, 0, 1, 3)
        return (to_agent_major(self.obs), self.share_obs[:, inds], to_agent_major(self.acts),
                to_agent_major(self.rewards), to_agent_major(self.dones), self.dones_env[:, inds])

    def sample(self, uniforms, beta):
        assert len(self) > len(uniforms) and beta > 0
        total = self.sum_tree.reduce(0, len(self) - 1)
        inds = self.sum_tree.find_prefixsum(np.asarray(uniforms) * total)
        whole = self.sum_tree.reduce()
        p_min = self.min_tree.reduce() / whole
        max_weight = (p_min * len(self)) ** (-beta)
        weights = (self.sum_tree.get(inds) / whole * len(self)) ** (-beta) / max_weight
        return self.gather(inds) + (weights, inds)

    def update_priorities(self, idxes, priorities):
        assert len(idxes) == len(priorities) and np.min(priorities) > 0
        assert np.min(idxes) >= 0 and np.max(idxes) < len(self)
        self.sum_tree.set(idxes, priorities ** self.alpha)
        self.min_tree.set(idxes, priorities ** self.alpha)
        self.max_priority = max(self.max_priority, np.max(priorities))
