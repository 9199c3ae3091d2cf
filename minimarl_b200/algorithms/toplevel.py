"""Device mirror of the top-level ``vdn/`` and ``qmix/`` packages.

* ``Q_Net`` -- per-agent ``Linear(D,64) ReLU Linear(64,32) ReLU -> GRUCell(32,32) -> Linear(32,5)`` with the
  reference's ParameterList names (vdn/_network.py:5-91; qmix/_network.py:5-77), ``sample_action`` with ONE
  epsilon draw per batch row (vdn/_network.py:52-58, quirk Q8) and float actions.
* ``Mix_Net`` -- recurrent hypernetwork mixer of qmix/_network.py:172-220.
* ``Prioritized_Experience_Replay`` -- chunk PER (vdn/replay_buffer/buffer.py:9-90, qmix/replay_buffer/per.py:9-81):
  payload as device tensors, priorities in the reference's heap-ordered float64 sum tree maintained by
  ``mm_heaptree_*`` kernels; stratified sampling, whole-tree decay (``step_weight``, vdn only), arg-min eviction.
* ``Target_Dqn`` / ``Train_dqn`` -- the DQN-target learners (vdn/_train.py:56-101, qmix/_train.py:19-121)
  including quirk Q7 (the bootstrap term summed over the agent axis) and the last-step-only priority (Q9).
* ``cal_td_error`` (vdn/_utils.py:44-52).

``Q_Net`` / ``Mix_Net`` are ordinary modules with the reference's parameter names (so ``state_dict`` / optimisers / target
copies work), and their ``forward`` is the module path used for single steps and as the parity statement.  The hot loops --
the rollout body, ``Target_Dqn.train`` and ``Train_dqn.train`` -- run on the fused kernels of ``toplevel_fused.py``
(``csrc/toplevel_qnet.cu``, ``csrc/toplevel_mix.cu``): one launch per chunk pass, BPTT with in-register weight gradients,
payload read through the PER slots without a gather.  ``fused=False`` (or an unsupported shape) keeps the module path.
Both top-level packages hard-wire the CPU in places (``.to("cpu")``, qmix/_network.py:64); the mirror runs wherever its
parameters live.
"""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib
from ..optim import clip_and_step
from . import toplevel_fused as tf


class Network(nn.Module):
    def __init__(self, observation_space, action_space, args):
        super().__init__()
        self.use_recurrent = args.use_recurrent
        self.num_agents = len(observation_space)
        self.observation_space, self.action_space = observation_space, action_space
        self.feature_hidden, self.gru_input, self.gru_hidden = 64, 32, 32
        self.feature_net = nn.ParameterList()
        self.gru_net = nn.ParameterList()
        self.action_net = nn.ParameterList()
        self.value_net = nn.ParameterList()
        self.advantage_net = nn.ParameterList()
        self._gen = None
        # qmix/_network.py registers the same per-agent modules with setattr names (feature_network_0, gru_network_0,
        # action_network_0 ...) where vdn/_network.py uses ParameterLists: accept both spellings when loading
        self._register_load_state_dict_pre_hook(self._accept_qmix_names)

    @staticmethod
    def _accept_qmix_names(state_dict, prefix, *unused):
        import re
        for key in list(state_dict.keys()):
            m = re.match(re.escape(prefix) + r"(feature|gru|action|value|advantage)_network_(\d+)(\..*)$", key)
            if m:
                state_dict[f"{prefix}{m.group(1)}_net.{m.group(2)}{m.group(3)}"] = state_dict.pop(key)

    def qmix_state_dict(self):
        """state_dict with the key spelling of qmix/_network.py (``feature_network_0.0.weight`` ...)."""
        import re
        return {re.sub(r"^(feature|gru|action|value|advantage)_net\.(\d+)\.", r"\1_network_\2.", k): v
                for k, v in self.state_dict().items()}

    def _feature(self, i):
        n_obs = self.observation_space[i].shape[0]
        self.feature_net.append(nn.Sequential(nn.Linear(n_obs, self.feature_hidden), nn.ReLU(),
                                              nn.Linear(self.feature_hidden, self.gru_input), nn.ReLU()))

    def epsilon_greedy(self, q_tensor, epsilon):
        """One uniform per batch row; a row that explores draws random actions for ALL its agents."""
        dev = q_tensor.device
        B, A, NA = q_tensor.shape
        mask = torch.rand(B, device=dev, generator=self._gen) <= epsilon
        rand = torch.randint(0, NA, (B, A), device=dev, generator=self._gen).float()
        action = torch.where(mask.unsqueeze(1), rand, q_tensor.argmax(dim=2).float())
        return action, q_tensor


class Q_Net(Network):
    def __init__(self, observation_space, action_space, args):
        super().__init__(observation_space, action_space, args)
        for i in range(self.num_agents):
            self._feature(i)
            if self.use_recurrent:
                self.gru_net.append(nn.GRUCell(self.gru_input, self.gru_hidden))
            self.action_net.append(nn.Sequential(nn.Linear(self.gru_hidden, self.action_space[i].n)))

    def forward(self, obs, hidden):
        q_values, next_hidden = [], []
        for i in range(self.num_agents):
            feature = self.feature_net[i](obs[:, i, :])
            if self.use_recurrent:
                feature = self.gru_net[i](feature, hidden[:, i, :])
                next_hidden.append(feature.unsqueeze(1))
            q_values.append(self.action_net[i](feature).unsqueeze(1))
        return torch.cat(q_values, dim=1), (torch.cat(next_hidden, dim=1) if next_hidden else hidden)

    def sample_action(self, obs, hidden, epsilon):
        q_tensor, next_hidden = self.forward(obs, hidden)
        action, q_tensor = self.epsilon_greedy(q_tensor, epsilon)
        return action, next_hidden, q_tensor

    def init_hidden(self, batch_size=1):
        p = next(self.parameters())
        return torch.zeros((batch_size, self.num_agents, self.gru_hidden), device=p.device)


class Mix_Net(nn.Module):
    def __init__(self, observation_space, args):
        super().__init__()
        hidden_dim, hx_size = 32, 32
        state_size = sum(s.shape[0] for s in observation_space)
        self.hidden_dim, self.hx_size, self.n_agents = hidden_dim, hx_size, len(observation_space)
        self.recurrent = args.use_recurrent
        inp = state_size
        if self.recurrent:
            self.gru = nn.GRUCell(state_size, hx_size)
            inp = hx_size
        self.hyper_net_weight_1 = nn.Linear(inp, self.n_agents * hidden_dim)
        self.hyper_net_weight_2 = nn.Linear(inp, hidden_dim)
        self.hyper_net_bias_1 = nn.Linear(inp, hidden_dim)
        self.hyper_net_bias_2 = nn.Sequential(nn.Linear(inp, hidden_dim), nn.ReLU(), nn.Linear(hidden_dim, 1))

    def forward(self, q_values, observations, hidden):
        batch_size, n_agents, obs_size = observations.shape
        x = observations.reshape(batch_size, n_agents * obs_size)
        if self.recurrent:
            hidden = self.gru(x, hidden)
            x = hidden
        weight_1 = torch.abs(self.hyper_net_weight_1(x)).view(batch_size, self.hidden_dim, n_agents)
        bias_1 = self.hyper_net_bias_1(x).unsqueeze(-1)
        weight_2 = torch.abs(self.hyper_net_weight_2(x))
        bias_2 = self.hyper_net_bias_2(x)
        x = torch.relu(torch.bmm(weight_1, q_values.unsqueeze(-1)) + bias_1)
        return (weight_2.unsqueeze(-1) * x).sum(dim=1) + bias_2, hidden

    def init_hidden(self, batch_size=1):
        p = next(self.parameters())
        return torch.zeros((batch_size, self.hx_size), device=p.device)


def cal_td_error(action, reward, done, behavior_q, target_q, gamma):
    """|sum r + (1 - done) gamma sum_a max Q_tgt - sum_a Q_beh[a]| for every env of the batch ([B] tensor)."""
    idx = action.long().unsqueeze(-1)
    behavior_value = torch.gather(behavior_q, 2, idx).squeeze(-1).sum(dim=1)
    target_max_q = target_q.max(dim=2)[0].sum(dim=1)
    reward = torch.as_tensor(reward, device=behavior_q.device, dtype=behavior_q.dtype)
    done = torch.as_tensor(done, device=behavior_q.device, dtype=behavior_q.dtype)
    return (reward.sum(dim=-1) + (1 - done) * gamma * target_max_q - behavior_value).abs()


class Test:
    """Greedy evaluation with the TD loss of the behaviour / target pair: ``vdn/_test.py:7-52`` (``qmix/_test.py`` is the
    same loop).  ``test_env`` is a lockstep batch (``BatchedCheckers`` / ``BatchedSwitch``): its boards play
    ``test_episodes`` episodes side by side (several passes if there are fewer boards than episodes); score and loss
    accumulate on the device and are read once at the end."""

    def __init__(self, test_env, args, device):
        self.test_env = test_env
        self.batch_size, self.chunk_size, self.update_iter = args.batch_size, args.chunk_size, args.update_iter
        self.test_episodes, self.gamma, self.grad_clip_norm = args.test_episodes, args.gamma, args.grad_clip_norm
        self.device = device

    @torch.no_grad()
    def execute(self, behavior_network, target_network):
        env, dev = self.test_env, self.device
        N = env.n_envs
        score = torch.zeros((), dtype=torch.float64, device=dev)
        iter_loss = torch.zeros((), dtype=torch.float64, device=dev)
        left = self.test_episodes
        while left > 0:
            count = torch.arange(N, device=dev) < left             # boards that stand for an episode in this pass
            state = env.reset()
            hidden, target_hidden = behavior_network.init_hidden(N), target_network.init_hidden(N)
            alive = count.clone()
            for _ in range(env.max_steps):
                action, next_hidden, behavior_q = behavior_network.sample_action(state, hidden, epsilon=0)
                next_state, reward, done = env.step(action)
                target_q, target_hidden = target_network(next_state, target_hidden)
                done_env = done.all(1)
                td = cal_td_error(action, reward, done_env.float(), behavior_q, target_q, self.gamma)
                iter_loss += (td.double() ** 2 * alive).sum()
                score += (reward.sum(1).double() * alive).sum()
                alive = alive & ~done_env
                state, hidden = next_state, next_hidden
            left -= N
        return float(score) / self.test_episodes, float(iter_loss) / self.test_episodes


class SumTree:
    """Heap-ordered float64 sum tree on the device (same node numbering as the reference)."""

    def __init__(self, capacity, device):
        self.capacity = capacity
        self.device = torch.device(device)
        self.priority_tree = torch.zeros(2 * capacity - 1, dtype=torch.float64, device=self.device)
        self.n_data = 0
        self._n_data_dev = None       # device copy of the fill count, owned by the batch ("set") insert
        self._ws = None

    def total(self):
        return self.priority_tree[0]

    def update(self, node_idx, priority):
        idx = torch.as_tensor(node_idx, device=self.device, dtype=torch.int64).reshape(-1).contiguous()
        pr = torch.as_tensor(priority, device=self.device).to(torch.float64).reshape(-1).contiguous()
        _lib.check(_lib.lib().mm_heaptree_update(_lib.ptr(self.priority_tree), self.capacity, _lib.ptr(idx), _lib.ptr(pr),
                                                 idx.numel(), _lib.stream()))

    def set_leaves(self, node_idx, priority):
        """Batch write-back (``mm_heaptree_set_leaves``): last write to a leaf wins, inner nodes recomputed from the leaves."""
        idx = torch.as_tensor(node_idx, device=self.device, dtype=torch.int64).reshape(-1).contiguous()
        pr = torch.as_tensor(priority, device=self.device).to(torch.float64).reshape(-1).contiguous()
        if self._ws is None or self._ws.numel() < self.capacity:
            self._ws = torch.empty(self.capacity, dtype=torch.int32, device=self.device)
        _lib.check(_lib.lib().mm_heaptree_set_leaves(_lib.ptr(self.priority_tree), self.capacity, _lib.ptr(idx), _lib.ptr(pr),
                                                     idx.numel(), _lib.ptr(self._ws), _lib.stream()))

    def rebuild(self):
        _lib.check(_lib.lib().mm_heaptree_rebuild(_lib.ptr(self.priority_tree), self.capacity, _lib.stream()))

    def add_set(self, priority, valid=None, rebuild=True):
        """Batch insert with set semantics (``mm_heaptree_add_set``); returns the slot of every item (-1 = not stored).  The
        fill count lives on the device; ``sync_n_data`` refreshes the host copy."""
        k = priority.numel()
        if self._n_data_dev is None:
            self._n_data_dev = torch.tensor([self.n_data], dtype=torch.int32, device=self.device)
        ws = torch.empty(k + 2 * self.capacity + 4, dtype=torch.int32, device=self.device)
        slot = torch.empty(k, dtype=torch.int64, device=self.device)
        _lib.check(_lib.lib().mm_heaptree_add_set(_lib.ptr(self.priority_tree), self.capacity, _lib.ptr(self._n_data_dev),
                                                  _lib.ptr(priority), _lib.ptr(valid), k, _lib.ptr(slot), _lib.ptr(ws),
                                                  int(bool(rebuild)), _lib.stream()))
        return slot

    def sync_n_data(self):
        if self._n_data_dev is not None:
            self.n_data = int(self._n_data_dev.item())
        return self.n_data

    def argmin_leaf(self):
        out = torch.empty(1, dtype=torch.int64, device=self.device)
        _lib.check(_lib.lib().mm_heaptree_argmin_leaf(_lib.ptr(self.priority_tree), self.capacity, _lib.ptr(out), _lib.stream()))
        return out

    def sample(self, uniforms):
        n = uniforms.numel()
        node = torch.empty(n, dtype=torch.int64, device=self.device)
        prio = torch.empty(n, dtype=torch.float64, device=self.device)
        _lib.check(_lib.lib().mm_heaptree_sample(_lib.ptr(self.priority_tree), self.capacity, _lib.ptr(uniforms), n,
                                                 _lib.ptr(node), _lib.ptr(prio), _lib.stream()))
        return node, prio


class Prioritized_Experience_Replay:
    def __init__(self, args, n_agents=2, obs_dim=94, device="cuda"):
        self.use_step_weight = getattr(args, "use_step_weight", False)
        self.capacity = args.buffer_limit
        self.step_weight = getattr(args, "step_weight", 1.0)
        self.eps, self.alpha, self.beta = args.eps, args.alpha, args.beta
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("minimarl_b200 replay buffers live on a CUDA device (no CPU fallback)")
        self.sum_tree = SumTree(self.capacity, self.device)
        L = args.chunk_size
        z = dict(dtype=torch.float32, device=self.device)
        # one spare row (index capacity) absorbs the payload of items that a later item of the same batch evicted
        rows = self.capacity + 1
        self.states = torch.zeros((rows, L, n_agents, obs_dim), **z)
        self.actions = torch.zeros((rows, L, n_agents), **z)
        self.rewards = torch.zeros((rows, L, n_agents), **z)
        self.next_states = torch.zeros((rows, L, n_agents, obs_dim), **z)
        self.dones = torch.zeros((rows, L, 1), **z)
        if args.update_alpha_beta:
            self.alpha_increment_per_sampling = (1 - self.alpha) / (args.max_episodes * args.update_iter)
            self.beta_increment_per_sampling = (1 - self.beta) / (args.max_episodes * args.update_iter)
        else:
            self.alpha_increment_per_sampling = self.beta_increment_per_sampling = 0
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(int(getattr(args, "seed", 0)))
        self._sched = None            # graph mode: [alpha, beta] as a device float64 pair, advanced inside the captured graph
        # batch formulations for lockstep envs (heaptree_batch.cu): set-semantics insert, write-back with tree rebuild.  Off by
        # default: the sequential kernels reproduce the reference's trees bit for bit.
        self.fast_tree = bool(getattr(args, "fast_tree", False))

    # -- CUDA-graph support: the annealed exponents must be device values, or a captured sample() would replay stale ones ----
    def enable_graph_mode(self):
        """``sample`` / ``update`` read alpha and beta from a device tensor that ``sample`` itself advances (so one captured
        training iteration can be replayed), and draw their uniforms from the default generator (which torch registers with
        a capture).  The host copies ``self.alpha`` / ``self.beta`` are advanced by ``advance_host_schedule`` once per
        iteration, with the same float64 arithmetic."""
        if self._sched is None:
            self._sched = torch.tensor([self.alpha, self.beta], dtype=torch.float64, device=self.device)
            self._sched_inc = torch.tensor([self.alpha_increment_per_sampling, self.beta_increment_per_sampling],
                                           dtype=torch.float64, device=self.device)

    def advance_host_schedule(self):
        self.alpha = float(np.min([1., self.alpha + self.alpha_increment_per_sampling]))
        self.beta = float(np.min([1., self.beta + self.beta_increment_per_sampling]))

    def _get_priority(self, td_error):
        """(td + eps) ** alpha in the arithmetic of the argument, like the reference: Python floats stay Python
        doubles (collect_sample), float32 tensors are raised in float32 and only then widened (update)."""
        if isinstance(td_error, (float, int, np.floating)):
            return (float(td_error) + self.eps) ** self.alpha
        td = torch.as_tensor(td_error, device=self.device)
        if self._sched is not None:
            return torch.pow(td + self.eps, self._sched[0].to(td.dtype)).to(torch.float64)
        return ((td + self.eps) ** self.alpha).to(torch.float64)

    def collect_sample(self, sample, td_error, warm_up):
        """sample = [states [L,A,D], actions [L,A] or [L,1,A], rewards [L,A], next_states [L,A,D], dones [L]]."""
        t = lambda x: torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x, device=self.device, dtype=torch.float32)
        tree = self.sum_tree
        if tree.n_data < self.capacity:
            slot = torch.tensor([tree.n_data], device=self.device)
        else:
            slot = tree.argmin_leaf()
        s, a, r, s2, d = (t(x) for x in sample)
        self.states[slot] = s.reshape(1, *self.states.shape[1:])
        self.actions[slot] = a.reshape(1, *self.actions.shape[1:])
        self.rewards[slot] = r.reshape(1, *self.rewards.shape[1:])
        self.next_states[slot] = s2.reshape(1, *self.next_states.shape[1:])
        self.dones[slot] = d.reshape(1, *self.dones.shape[1:])
        tree.update(slot + self.capacity - 1,
                    torch.as_tensor(self._get_priority(td_error), dtype=torch.float64, device=self.device).reshape(1))
        if tree.n_data < self.capacity:
            tree.n_data += 1
        return tree.n_data if warm_up else None

    def collect_samples_batch(self, samples, td_errors, warm_up, host_pow=False, valid=None, rebuild=True, priorities=None):
        """K consecutive ``collect_sample`` calls -- the chunks N lockstep envs finish at the same step, in env order
        (the reference's add order) -- without a Python loop: ``mm_heaptree_add_batch`` replays the K ``SumTree.add``s
        on the device (fill, then arg-min eviction where a later chunk may evict an earlier one of the same batch), the
        payloads are scattered into the slots their final owners got.
        samples = [states [K,L,A,D], actions [K,L,A] or [K,L,1,A], rewards [K,L,A], next_states [K,L,A,D], dones [K,L]];
        td_errors [K].  ``host_pow``: raise (td + eps) ** alpha with Python doubles like the reference (one D2H sync)
        instead of torch.pow on the device (<= 1 ulp from libm).
        ``fast_tree`` or ``valid`` (uint8 [K]: chunks of envs that had already finished are not stored): the batch goes in as a
        SET (``mm_heaptree_add_set`` + one ``mm_rows_scatter``), see heaptree_batch.cu; ``rebuild=False`` defers the inner-node
        recomputation (call ``sum_tree.rebuild()`` after the last insert of a run), ``priorities`` (float64 [K]) replaces
        ``(td + eps) ** alpha`` when the caller raised a whole episode's TD sums at once."""
        tree = self.sum_tree
        s, a, r, s2, d = (torch.as_tensor(x, device=self.device, dtype=torch.float32) for x in samples)
        K = s.shape[0]
        td = torch.as_tensor(td_errors, device=self.device).reshape(K)
        if self.fast_tree or valid is not None:
            return self._collect_set(s, a, r, s2, d, td, valid, warm_up, rebuild, priorities)
        if host_pow:
            prio = torch.tensor([(float(v) + self.eps) ** self.alpha for v in td.tolist()], dtype=torch.float64, device=self.device)
        else:
            prio = torch.pow(td.to(torch.float64) + self.eps, self.alpha)
        slot = torch.empty(K, dtype=torch.int64, device=self.device)
        rc = _lib.lib().mm_heaptree_add_batch(_lib.ptr(tree.priority_tree), self.capacity, tree.n_data, _lib.ptr(prio.contiguous()),
                                              K, _lib.ptr(slot), _lib.stream())
        if rc == -3:                                      # arg-min tree larger than shared memory: one by one
            if not getattr(self, "_warned_big_tree", False):
                import warnings
                warnings.warn(f"mm_heaptree_add_batch: a tree of {self.capacity} slots does not fit in shared memory (limit ~14 000); "
                              "the reference-order batch insert falls back to one add per chunk (bit-identical, ~0.2 ms each) -- "
                              "fast_tree=True (set-semantics insert, heaptree_batch.cu) has no such limit", stacklevel=2)
                self._warned_big_tree = True
            for i in range(K):
                self.collect_sample([s[i], a[i], r[i], s2[i], d[i]], float(td[i]), warm_up)
            return tree.n_data if warm_up else None
        _lib.check(rc)
        order = torch.arange(K, device=self.device)
        last = torch.full((self.capacity,), -1, dtype=torch.int64, device=self.device).scatter_reduce_(0, slot, order, "amax")
        dest = torch.where(last[slot] == order, slot, torch.full_like(slot, self.capacity))
        self.states[dest] = s.reshape(K, *self.states.shape[1:])
        self.actions[dest] = a.reshape(K, *self.actions.shape[1:])
        self.rewards[dest] = r.reshape(K, *self.rewards.shape[1:])
        self.next_states[dest] = s2.reshape(K, *self.next_states.shape[1:])
        self.dones[dest] = d.reshape(K, *self.dones.shape[1:])
        tree.n_data = min(self.capacity, tree.n_data + K)
        return tree.n_data if warm_up else None

    def priorities_from_td(self, td):
        """(td + eps) ** alpha in float64 on the device, alpha from the device schedule in graph mode."""
        alpha = self._sched[0] if self._sched is not None else self.alpha
        return torch.pow(td.to(torch.float64) + self.eps, alpha)

    def _collect_set(self, s, a, r, s2, d, td, valid, warm_up, rebuild=True, priorities=None):
        tree = self.sum_tree
        K = s.shape[0]
        prio = (priorities if priorities is not None else self.priorities_from_td(td)).contiguous()
        if valid is not None and (valid.dtype != torch.uint8 or not valid.is_contiguous()):
            valid = valid.to(device=self.device, dtype=torch.uint8).contiguous()
        slot = tree.add_set(prio, valid, rebuild=rebuild)
        src = [s.reshape(K, -1), a.reshape(K, -1), r.reshape(K, -1), s2.reshape(K, -1), d.reshape(K, -1)]
        dst = [self.states, self.actions, self.rewards, self.next_states, self.dones]
        src = [x if x.is_contiguous() else x.contiguous() for x in src]
        sp = (C.c_void_p * 5)(*[x.data_ptr() for x in src])
        dp = (C.c_void_p * 5)(*[x.data_ptr() for x in dst])
        rf = (C.c_int32 * 5)(*[x.shape[1] for x in src])
        _lib.check(_lib.lib().mm_rows_scatter(sp, dp, rf, 5, _lib.ptr(slot), K, _lib.stream()))
        if warm_up:
            return tree.sync_n_data()
        return None

    def collect_episode(self, chunks, td_sums, valid):
        """All chunks of one lockstep episode: ``chunks`` = [states [C,N,L,A,D], actions [C,N,L,A], rewards [C,N,L,A], next_states,
        dones [C,N,L]] (contiguous, chunk-major), ``td_sums`` / ``valid`` [C,N].  Equivalent to C ``collect_samples_batch`` calls
        with set semantics (chunk order = the reference's add order), issued as one library call."""
        tree = self.sum_tree
        Cn, N = td_sums.shape
        prio = self.priorities_from_td(td_sums).contiguous()
        valid = valid.to(device=self.device, dtype=torch.uint8).contiguous()
        if tree._n_data_dev is None:
            tree._n_data_dev = torch.tensor([tree.n_data], dtype=torch.int32, device=self.device)
        ws = torch.empty(N + 2 * self.capacity + 4, dtype=torch.int32, device=self.device)
        slot = torch.empty((Cn, N), dtype=torch.int64, device=self.device)
        dst = [self.states, self.actions, self.rewards, self.next_states, self.dones]
        for x in chunks:
            if not (x.is_cuda and x.is_contiguous() and x.dtype == torch.float32):
                raise RuntimeError("collect_episode takes contiguous float32 CUDA tensors")
        sp = (C.c_void_p * 5)(*[x.data_ptr() for x in chunks])
        dp = (C.c_void_p * 5)(*[x.data_ptr() for x in dst])
        rf = (C.c_int32 * 5)(*[x[0, 0].numel() for x in chunks])
        _lib.check(_lib.lib().mm_chunk_per_insert_episode(
            _lib.ptr(tree.priority_tree), self.capacity, _lib.ptr(tree._n_data_dev), _lib.ptr(prio), _lib.ptr(valid), Cn, N,
            _lib.ptr(slot), _lib.ptr(ws), sp, dp, rf, 5, _lib.stream()))
        return slot

    def sample_indices(self, batch_size, uniforms=None):
        """The stratified draw without the payload gather: (data slots [B] int64, node ids [B] int64, IS weights [B,1] f32)."""
        data, node, w = self._draw(batch_size, uniforms)
        return data, node, w

    def sample(self, batch_size, _chunk_size, uniforms=None):
        """7-tuple of the reference (buffer.py:81-86), all CUDA tensors; priority_idxs is an int64 tensor of NODE ids."""
        data, node, w = self._draw(batch_size, uniforms)
        return (self.states[data], self.actions[data], self.rewards[data], self.next_states[data], self.dones[data],
                node, w)

    def _draw(self, batch_size, uniforms=None):
        tree = self.sum_tree
        if self._sched is not None:
            self._sched.add_(self._sched_inc).clamp_(max=1.0)
            beta = self._sched[1]
        else:
            self.advance_host_schedule()
            beta = self.beta
        if uniforms is None:
            uniforms = torch.rand(batch_size, dtype=torch.float64, device=self.device,
                                  generator=None if self._sched is not None else self._gen)
        else:
            uniforms = torch.as_tensor(np.asarray(uniforms) if not torch.is_tensor(uniforms) else uniforms,
                                       dtype=torch.float64, device=self.device).contiguous()
        node, prio = tree.sample(uniforms)
        data = node - (self.capacity - 1)
        if self.use_step_weight:
            tree.priority_tree.mul_(self.step_weight)
        probs = prio / tree.total()
        w = torch.pow(self.capacity * probs, -beta)
        w = (w / w.max()).to(torch.float32).view(batch_size, 1)
        return data, node, w

    def update(self, priority_idx, new_td_error):
        if self.fast_tree:
            self.sum_tree.set_leaves(priority_idx, self._get_priority(new_td_error))
        else:
            self.sum_tree.update(priority_idx, self._get_priority(new_td_error))


class _GraphedIterations:
    """Runs ``update_iter`` training iterations; with ``use_graph`` the iteration (PER draw, chunk forward of both networks,
    loss, backward, clipping, Adam step, priority write-back: ~700 small launches) is captured ONCE as a CUDA graph and
    replayed -- the loop is launch-bound from Python, not GPU-bound.  The first iteration of the first call runs eagerly on a
    side stream (it is a real update and doubles as the warm-up a capture needs); the optimiser must be ``capturable``."""

    use_graph = False
    _graph = None

    def _run_iterations(self, per, optimizer, iteration):
        if not self.use_graph:
            total = 0
            for _ in range(self.update_iter):
                total = total + iteration(None)
                if per._sched is not None:                       # device-side schedule: the host copy follows
                    per.advance_host_schedule()
            return total / self.update_iter
        done, total = 0, 0
        if self._graph is None:
            per.enable_graph_mode()
            main = torch.cuda.current_stream()
            side = torch.cuda.Stream(device=per.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                total = total + iteration(None).clone()
            main.wait_stream(side)
            per.advance_host_schedule()
            done = 1
            if not getattr(self, "fused", False):
                optimizer.zero_grad(set_to_none=True)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._static_loss = iteration(None)
        for _ in range(self.update_iter - done):
            self._graph.replay()
            per.advance_host_schedule()
            total = total + self._static_loss
        return total / self.update_iter


class Target_Dqn(_GraphedIterations):
    """vdn/_train.py:56-101."""

    def __init__(self, Replay_buffer, behavior_network, target_network, args, device):
        self.replay_buffer, self.behavior_network, self.target_network = Replay_buffer, behavior_network, target_network
        self.batch_size, self.update_iter = args.batch_size, args.update_iter
        self.chunk_size = args.chunk_size if args.use_recurrent else 1
        self.gamma, self.grad_clip_norm, self.device = args.gamma, args.grad_clip_norm, device
        self.use_graph = bool(getattr(args, "train_graph", False))
        self.fused = bool(getattr(args, "fused", True))

    def _iteration_fused(self, target_network, optimizer, u):
        """One iteration on the fused kernels: target pass and behaviour pass (activations kept) read the sampled chunks
        through their PER slots, the loss kernel emits dq, one backward kernel writes every weight gradient into ``p.grad``."""
        per, beh, B, L = self.replay_buffer, self.behavior_network, self.batch_size, self.chunk_size
        data, node, w = per.sample_indices(B, uniforms=u)
        q_tgt = tf.qnet_seq(target_network, per.next_states, B, L, row_index=data, dones=per.dones, reset_on_done=True)
        q_beh, save = tf.qnet_seq(beh, per.states, B, L, row_index=data, dones=per.dones, reset_on_done=True, save=True)
        se, dq, td_last = tf.vdn_chunk_loss(q_beh, q_tgt, per.actions, per.rewards, per.dones, data, w.reshape(-1).contiguous(),
                                            self.gamma)
        tf.qnet_backward(beh, per.states, data, per.dones, save, dq, B, L)
        clip_and_step(optimizer, beh.parameters(), self.grad_clip_norm)
        per.update(node, td_last)
        return se.sum() / B

    def _can_fuse(self, target_network):
        return (self.fused and self.chunk_size > 1 and tf.qnet_supported(self.behavior_network) and tf.qnet_supported(target_network))

    def _iteration(self, target_network, optimizer, u):
        if self._can_fuse(target_network):
            return self._iteration_fused(target_network, optimizer, u)
        states, actions, rewards, next_states, dones, priority_idx, is_weight = \
            self.replay_buffer.sample(self.batch_size, self.chunk_size, uniforms=u)
        hidden = self.behavior_network.init_hidden(self.batch_size)
        target_hidden = target_network.init_hidden(self.batch_size)
        loss = 0.0
        for step in range(self.chunk_size):
            state, action = states[:, step], actions[:, step].unsqueeze(-1).long()
            reward, next_state, done = rewards[:, step], next_states[:, step], dones[:, step]
            behavior_q, next_hidden = self.behavior_network(obs=state, hidden=hidden)
            sum_q = behavior_q.gather(dim=2, index=action).squeeze(-1).sum(dim=1, keepdims=True)
            with torch.no_grad():
                target_q, next_target_hidden = target_network(obs=next_state, hidden=target_hidden.detach())
            sum_max_target_q = target_q.max(dim=2)[0].sum(dim=1, keepdims=True)
            target_values = is_weight * (reward + self.gamma * (1 - done) * sum_max_target_q).sum(dim=1, keepdims=True)
            loss = loss + F.mse_loss(target_values.detach(), sum_q)
            keep = (1 - done).view(-1, 1, 1)                       # hidden rows of finished envs restart at zero
            hidden, target_hidden = next_hidden * keep, next_target_hidden * keep
        optimizer.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(parameters=self.behavior_network.parameters(), max_norm=self.grad_clip_norm, norm_type=2)
        optimizer.step()
        new_td_error = (target_values - sum_q).abs().detach().reshape(-1)   # last chunk step only (Q9)
        self.replay_buffer.update(priority_idx, new_td_error)
        return loss.detach()

    def train(self, target_network, optimizer, epsilon, uniforms=None):
        if uniforms is not None:                                   # recorded uniforms (tests): always the eager loop
            total = 0
            for it in range(self.update_iter):
                total = total + self._iteration(target_network, optimizer, uniforms[it])
            return total / self.update_iter
        return self._run_iterations(self.replay_buffer, optimizer, lambda u: self._iteration(target_network, optimizer, u))


class Train_dqn(_GraphedIterations):
    """qmix/_train.py:19-121: the same loop with the recurrent mixer on both sides; only q-net grads are clipped."""

    def __init__(self, args, device):
        self.batch_size, self.update_iter = args.batch_size, args.update_iter
        self.chunk_size = args.chunk_size if args.use_recurrent else 1
        self.gamma, self.grad_clip_norm, self.device = args.gamma, args.grad_clip_norm, device
        self.use_graph = bool(getattr(args, "train_graph", False))
        self.fused = bool(getattr(args, "fused", True))

    def _iteration_fused(self, per, q_net, mix, tq_net, tmix, optimizer, u):
        """qmix/_train.py:29-121 on the fused kernels: Q_Net target / behaviour passes, Mix_Net on both sides, loss, mixer BPTT
        (emits dq), Q_Net BPTT; only the q-net gradients are clipped, as in the reference."""
        B, L = self.batch_size, self.chunk_size
        data, node, w = per.sample_indices(B, uniforms=u)
        q_tgt = tf.qnet_seq(tq_net, per.next_states, B, L, row_index=data, dones=per.dones, reset_on_done=True)
        qtot_tgt = tf.mix_seq(tmix, per.next_states, data, per.dones, q_tgt, None, B, L)
        q_beh, save = tf.qnet_seq(q_net, per.states, B, L, row_index=data, dones=per.dones, reset_on_done=True, save=True)
        qtot, msave = tf.mix_seq(mix, per.states, data, per.dones, q_beh, per.actions, B, L, save=True)
        se, dqtot, td_last = tf.qmix_chunk_loss(qtot, qtot_tgt, per.rewards, per.dones, data, w.reshape(-1).contiguous(), self.gamma)
        dq = tf.mix_backward(mix, per.states, data, per.dones, q_beh, per.actions, msave, dqtot, B, L)
        tf.qnet_backward(q_net, per.states, data, per.dones, save, dq, B, L)
        clip_and_step(optimizer, q_net.parameters(), self.grad_clip_norm, prefix=True)
        per.update(node, td_last)
        return se.sum() / B

    def _iteration(self, Replay_buffer, behavior_q_net, behavior_mix_net, target_q_net, target_mix_net, optimizer, u):
        if (self.fused and self.chunk_size > 1 and tf.qnet_supported(behavior_q_net) and tf.qnet_supported(target_q_net)
                and tf.mix_supported(behavior_mix_net) and tf.mix_supported(target_mix_net)):
            return self._iteration_fused(Replay_buffer, behavior_q_net, behavior_mix_net, target_q_net, target_mix_net, optimizer, u)
        states, actions, rewards, next_states, dones, priority_idx, w = Replay_buffer.sample(self.batch_size, self.chunk_size, uniforms=u)
        hidden, target_hidden = behavior_q_net.init_hidden(self.batch_size), target_q_net.init_hidden(self.batch_size)
        mix_hidden, target_mix_hidden = behavior_mix_net.init_hidden(self.batch_size), target_mix_net.init_hidden(self.batch_size)
        loss = 0.0
        for step in range(self.chunk_size):
            state, action = states[:, step], actions[:, step].unsqueeze(-1).long()
            reward, next_state, done = rewards[:, step], next_states[:, step], dones[:, step]
            behavior_q, next_hidden = behavior_q_net(state, hidden)
            max_behavior_q = behavior_q.gather(dim=2, index=action).squeeze(-1)
            sum_mix_q, next_mix_hidden = behavior_mix_net(max_behavior_q, state, mix_hidden)
            with torch.no_grad():
                target_q, next_target_hidden = target_q_net(next_state, target_hidden.detach())
                max_target_q = target_q.max(dim=2)[0]
                sum_target_mix_q, next_target_mix_hidden = target_mix_net(max_target_q, next_state, target_mix_hidden.detach())
            target_values = w * (reward + self.gamma * (1 - done) * sum_target_mix_q).sum(dim=1, keepdims=True)
            loss = loss + F.mse_loss(target_values.detach(), sum_mix_q)
            keep = (1 - done)
            hidden, target_hidden = next_hidden * keep.view(-1, 1, 1), next_target_hidden * keep.view(-1, 1, 1)
            mix_hidden, target_mix_hidden = next_mix_hidden * keep, next_target_mix_hidden * keep
        optimizer.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(parameters=behavior_q_net.parameters(), max_norm=self.grad_clip_norm, norm_type=2)
        optimizer.step()
        Replay_buffer.update(priority_idx, (target_values - sum_mix_q).abs().detach().reshape(-1))
        return loss.detach()

    def train(self, Replay_buffer, behavior_q_net, behavior_mix_net, target_q_net, target_mix_net, optimizer, epsilon,
              uniforms=None):
        nets = (Replay_buffer, behavior_q_net, behavior_mix_net, target_q_net, target_mix_net, optimizer)
        if uniforms is not None:                                   # recorded uniforms (tests): always the eager loop
            total = 0
            for it in range(self.update_iter):
                total = total + self._iteration(*nets, uniforms[it])
            return total / self.update_iter
        return self._run_iterations(Replay_buffer, optimizer, lambda u: self._iteration(*nets, u))
