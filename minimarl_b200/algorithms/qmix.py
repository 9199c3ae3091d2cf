"""Recurrent QMIX / VDN policy and trainer with the reference's interface, device-resident.

``QMixPolicy`` mirrors offpolicy/algorithms/qmix/algorithm/QMixPolicy.py:10-226 and ``QMix`` mirrors
offpolicy/algorithms/qmix/qmix.py:13-242 (``VDN`` = ``QMix(..., vdn=True)``, vdn/vdn.py:4-7): same
constructor arguments, ``get_q_values / get_actions / get_random_actions / init_hidden / parameters /
load_state``, ``train_policy_on_batch(batch) -> (train_info, new_priorities, idxes)``,
``soft_target_updates / hard_target_updates / prep_training / prep_rollout``.

What runs where in one update (batch B episodes x T steps x A agents):
  * target network over the whole sequence            -> ONE fused kernel (mm_agentq_forward)
  * greedy arg-max, chosen-action gather, next-Q pick  -> ONE kernel (mm_q_select)
  * target mixer                                       -> ONE fused kernel (mm_qmixer_forward)
  * TD target, bad-transition mask, PER loss, priorities -> ONE kernel (mm_qmix_td) behind an autograd.Function
  * the differentiated live pass (agent net + mixer)   -> the identical eager PyTorch layer chain (autograd)
Multi-GPU: when torch.distributed is initialised the flattened gradient bucket is all-reduced (sum) and
divided by the world size before clipping -- the semantics of the reference's dead ``average_gradients``
(offpolicy/utils/util.py:148-153).
"""
import copy

import numpy as np
import torch

from ..nn import AgentQFunction, QMixer, VDNMixer
from ..nn._direct import direct_param_grads
from ..ops.td import q_select, qmix_td_loss
from ..optim import FlatAdam, clip_and_step, make_adam
from ..parallel import GradBucket, allreduce_mean_grads


def _space_dim(space):
    name = space.__class__.__name__
    if name == "Box":
        return int(space.shape[0])
    if name == "Discrete":
        return int(space.n)
    raise NotImplementedError(name)


class DecayThenFlatSchedule:
    """offpolicy/utils/util.py:78-100 (linear / exp decay then flat)."""

    def __init__(self, start, finish, time_length, decay="exp"):
        self.start, self.finish, self.time_length, self.decay = start, finish, time_length, decay
        self.delta = (start - finish) / time_length
        if decay == "exp":
            self.exp_scaling = (-1) * time_length / np.log(finish) if finish > 0 else 1

    def eval(self, T):
        if self.decay == "linear":
            return max(self.finish, self.start - self.delta * T)
        return min(self.start, max(self.finish, np.exp(-T / self.exp_scaling)))


def soft_update(target, source, tau):
    with torch.no_grad():
        tp, sp = list(target.parameters()), list(source.parameters())
        torch._foreach_mul_(tp, 1.0 - tau)
        torch._foreach_add_(tp, sp, alpha=tau)


def _t(x, device, dtype=torch.float32):
    if torch.is_tensor(x):
        return x.to(device=device, dtype=dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype, device=device)


class QMixPolicy:
    def __init__(self, args, config, policy_config, train=True):
        self.args = config["args"]
        self.device = torch.device(config["device"])
        self.num_agents = config["num_agents"]
        self.batch_size = args.batch_size
        self.obs_space = policy_config["obs_space"]
        self.obs_dim = _space_dim(self.obs_space)
        self.act_space = policy_config["act_space"]
        self.act_dim = _space_dim(self.act_space)
        self.output_dim = self.act_dim
        self.hidden_size = self.args.hidden_size
        self.central_obs_dim = policy_config["cent_obs_dim"]
        self.discrete, self.multidiscrete = True, False
        self.tpdv = dict(dtype=torch.float32, device=self.device)
        self.q_network_input_dim = self.obs_dim + self.act_dim if self.args.prev_act_inp else self.obs_dim
        self.q_network = AgentQFunction(self.args, self.q_network_input_dim, self.act_dim, self.device)
        if train:
            self.exploration = DecayThenFlatSchedule(self.args.epsilon_start, self.args.epsilon_finish,
                                                     self.args.epsilon_anneal_time, decay="linear")
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(int(getattr(self.args, "seed", 0)))

    def get_q_values(self, obs_batch, prev_action_batch, rnn_states, action_batch=None):
        obs_batch = _t(obs_batch, self.device)
        if self.args.prev_act_inp:
            obs_batch = torch.cat((obs_batch, _t(prev_action_batch, self.device)), dim=-1)
        q_batch, new_rnn_states = self.q_network(obs_batch, _t(rnn_states, self.device))
        if action_batch is not None:
            return self.q_values_from_actions(q_batch, _t(action_batch, self.device)), new_rnn_states
        return q_batch, new_rnn_states

    def q_values_from_actions(self, q_batch, action_batch):
        ind = action_batch.max(dim=-1)[1]
        return torch.gather(q_batch, 2, ind.unsqueeze(dim=-1))

    def get_actions(self, obs, prev_actions, rnn_states, available_actions=None, t_env=None, explore=False):
        q_values_out, new_rnn_states = self.get_q_values(obs, prev_actions, rnn_states)
        onehot_actions, greedy_Qs = self.actions_from_q(q_values_out, available_actions=available_actions,
                                                        explore=explore, t_env=t_env)
        return onehot_actions, new_rnn_states, greedy_Qs

    def actions_from_q(self, q_values, available_actions=None, explore=False, t_env=None):
        """eps-greedy (one draw per agent row, linear eps schedule) on the device; one-hot float32 CUDA tensor."""
        no_sequence = q_values.dim() == 2
        batch_size = q_values.shape[0] if no_sequence else q_values.shape[1]
        if available_actions is not None:
            q_values = q_values.clone()
            q_values[_t(available_actions, self.device) == 0] = -1e10
        greedy_Qs, greedy_actions = q_values.max(dim=-1)
        if explore:
            assert no_sequence, "Can only explore on non-sequences"
            eps = self.exploration.eval(t_env)
            rand_numbers = torch.rand(batch_size, device=self.device, generator=self._gen)
            random_actions = torch.randint(0, self.act_dim, (batch_size,), device=self.device, generator=self._gen)
            actions = torch.where(rand_numbers < eps, random_actions, greedy_actions)
        else:
            greedy_Qs = greedy_Qs.unsqueeze(-1)
            actions = greedy_actions
        return torch.nn.functional.one_hot(actions, self.act_dim).to(torch.float32), greedy_Qs

    def get_random_actions(self, obs, available_actions=None):
        batch_size = obs.shape[0]
        a = torch.randint(0, self.act_dim, (batch_size,), device=self.device, generator=self._gen)
        return torch.nn.functional.one_hot(a, self.act_dim).to(torch.float32)

    def init_hidden(self, num_agents, batch_size):
        if num_agents == -1:
            return torch.zeros(batch_size, self.hidden_size, device=self.device)
        return torch.zeros(num_agents, batch_size, self.hidden_size, device=self.device)

    def parameters(self):
        return self.q_network.parameters()

    def load_state(self, source_policy):
        self.q_network.load_state_dict(source_policy.q_network.state_dict())


class QMix:
    def __init__(self, args, num_agents, batch_size, policies, policy_mapping_fn, device=torch.device("cuda:0"),
                 episode_length=None, vdn=False):
        self.args = args
        self.use_per = args.use_per
        self.per_eps = args.per_eps
        self.use_huber_loss = args.use_huber_loss
        if self.use_huber_loss or args.use_popart:
            raise NotImplementedError("reference defaults only: MSE loss, no PopArt (offpolicy/config.py:119-136)")
        self.device = torch.device(device)
        self.tpdv = dict(dtype=torch.float32, device=self.device)
        self.lr, self.tau, self.opti_eps = args.lr, args.tau, args.opti_eps
        self.episode_length = args.episode_length if episode_length is None else episode_length
        self.num_agents = num_agents
        self.policies = policies
        self.policy_mapping_fn = policy_mapping_fn
        self.policy_ids = sorted(list(policies.keys()))
        self.policy_agents = {pid: sorted([a for a in range(num_agents) if policy_mapping_fn(a) == pid])
                              for pid in policies.keys()}
        self.use_same_share_obs = args.use_same_share_obs
        mixer_cls = VDNMixer if vdn else QMixer
        self.mixer = mixer_cls(args, num_agents, policies["policy_0"].central_obs_dim, self.device)
        self.target_policies = {pid: copy.deepcopy(policies[pid]) for pid in self.policy_ids}
        self.target_mixer = copy.deepcopy(self.mixer)
        self.parameters = []
        for policy in policies.values():
            self.parameters += policy.parameters()
        self.parameters += self.mixer.parameters()
        # FlatAdam (optim.py): parameters and gradients live in flat buffers, clip + Adam are two launches, the step counter
        # lives on the device, so a whole update can be replayed as a CUDA graph (GraphedUpdate)
        self.optimizer = make_adam(self.parameters, lr=self.lr, eps=self.opti_eps)
        if isinstance(self.optimizer, FlatAdam):
            self._bucket = self.optimizer.bucket
        self.batch_size = batch_size
        self.overlap_target = True
        self._side_stream = torch.cuda.Stream(device=self.device)

    # -- gradient exchange (the only collective on this path) -------------------------------------
    def _allreduce_grads(self):
        allreduce_mean_grads(self.parameters, bucket=getattr(self, "_bucket", None))

    def _zero_grads(self):
        """Multi-GPU: gradients of every parameter live in one flat ``GradBucket`` (views), see parallel.py."""
        import torch.distributed as dist
        if isinstance(self.optimizer, FlatAdam):
            self.optimizer.zero_grad()
        elif dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            b = getattr(self, "_bucket", None)
            if b is None or not b.attached():
                b = self._bucket = GradBucket(self.parameters)
            b.zero()
        else:
            self.optimizer.zero_grad()

    def train_policy_on_batch(self, batch, update_policy_id=None):
        obs_batch, cent_obs_batch, act_batch, rew_batch, dones_batch, dones_env_batch, avail_act_batch, \
            importance_weights, idxes = batch
        p_id = self.policy_ids[0]
        dev = self.device
        cent_obs = _t(cent_obs_batch[p_id], dev)                     # [T+1, B, SD]
        dones_env = _t(dones_env_batch[p_id], dev)                   # [T, B, 1]
        policy, target_policy = self.policies[p_id], self.target_policies[p_id]
        pol_obs = _t(obs_batch[p_id], dev)                           # [A, T+1, B, D]
        acts = _t(act_batch[p_id], dev)                              # [A, T, B, NA]
        A, T1, B = pol_obs.shape[0], pol_obs.shape[1], pol_obs.shape[2]
        T = T1 - 1
        # rows r = a*B + b (torch.cat(list(x), dim=-2) in the reference); a view when the replay gathered in that order
        stacked_obs = pol_obs.permute(1, 0, 2, 3).reshape(T1, A * B, -1)      # [T+1, A*B, D]
        stacked_acts = acts.permute(1, 0, 2, 3).reshape(T, A * B, -1)         # [T, A*B, NA]
        total = A * B
        prev_acts = None
        if self.args.prev_act_inp:
            prev_acts = torch.cat((torch.zeros(1, total, policy.act_dim, **self.tpdv), stacked_acts))

        # The target side (no gradient: fused kernels of 64 CTAs each, about half the SMs) runs on a second stream next to
        # the live pass: target network ‖ live forward, then target mixer ‖ live mixer; the streams meet at q_select and at
        # the TD loss.  Tensors crossing streams are pinned to their consumer with record_stream.
        main = torch.cuda.current_stream(dev)
        side = self._side_stream if self.overlap_target else main
        if side is not main:
            side.wait_stream(main)
        with torch.no_grad(), torch.cuda.stream(side):
            tq_all, _ = target_policy.get_q_values(stacked_obs, prev_acts, target_policy.init_hidden(-1, total))
        # live network: the differentiated pass
        q_all, _ = policy.get_q_values(stacked_obs, prev_acts, policy.init_hidden(-1, total))
        capturing = torch.cuda.is_current_stream_capturing()       # a graph's private pool needs no cross-stream pinning
        if side is not main:
            main.wait_stream(side)
            if not capturing:
                tq_all.record_stream(main)
        with torch.no_grad():
            _, next_q, act_idx, _ = q_select(q_all.detach(), tq_all, stacked_acts, T, B, A,
                                             double_q=self.args.use_double_q)
        if side is not main:
            side.wait_stream(main)
        with torch.no_grad(), torch.cuda.stream(side):
            next_Q_tot = torch.reshape(self.target_mixer(next_q, cent_obs[1:]), (-1, B, 1))
        # chosen-action Q of the live net, [T, A*B] -> [T, B, A]
        q_taken = torch.gather(q_all[:-1], 2, act_idx.unsqueeze(-1)).squeeze(-1)
        agent_q_seq = q_taken.view(T, A, B).permute(0, 2, 1)

        Q_tot_seq = torch.reshape(self.mixer(agent_q_seq, cent_obs[:-1]), (-1, B, 1))
        if side is not main:
            main.wait_stream(side)
            if not capturing:
                next_Q_tot.record_stream(main)
        rewards = _t(rew_batch[p_id], dev)[0]                        # agents share the reward: agent 0's copy

        w = _t(importance_weights, dev) if (self.use_per and importance_weights is not None) else None
        loss, prio, q_tot_mean, _ = qmix_td_loss(Q_tot_seq, next_Q_tot, rewards, dones_env, w, self.args.gamma,
                                                 self.args.per_nu, self.per_eps)
        new_priorities = prio if self.use_per else None

        self._zero_grads()
        with direct_param_grads():                       # the fused ops add their weight gradients into the bucket views themselves
            loss.backward()
        self._allreduce_grads()
        grad_norm = clip_and_step(self.optimizer, self.parameters, self.args.max_grad_norm)
        train_info = {"loss": loss.detach(), "grad_norm": grad_norm, "Q_tot": q_tot_mean}
        return train_info, new_priorities, idxes

    def hard_target_updates(self):
        for pid in self.policy_ids:
            self.target_policies[pid].load_state(self.policies[pid])
        if self.mixer is not None:
            self.target_mixer.load_state_dict(self.mixer.state_dict())

    def soft_target_updates(self):
        for pid in self.policy_ids:
            soft_update(self.target_policies[pid], self.policies[pid], self.tau)
        if self.mixer is not None and len(list(self.mixer.parameters())):
            soft_update(self.target_mixer, self.mixer, self.tau)

    def prep_training(self):
        for pid in self.policy_ids:
            self.policies[pid].q_network.train()
            self.target_policies[pid].q_network.train()
        self.mixer.train()
        self.target_mixer.train()

    def prep_rollout(self):
        for pid in self.policy_ids:
            self.policies[pid].q_network.eval()
            self.target_policies[pid].q_network.eval()
        self.mixer.eval()
        self.target_mixer.eval()


class GraphedUpdate:
    """One learner update -- PER draw + episode gather, ``train_policy_on_batch`` (both streams), priority write-back, soft target
    update: ~230 launches issued from Python in ~7 ms, against ~6 ms of device work -- captured ONCE as a CUDA graph and replayed.
    The reference's loop (offpolicy/runner/shared/base_runner.py:batch_train_q -> QMix.train_policy_on_batch ->
    buffer.update_priorities) is unchanged in content; what is dropped is the host-side assert of update_priorities (indices are
    clamped on the device instead) and the host copy of the running maximum priority (kept on the device, read at the next
    insert).  The uniforms of the PER draw are generated outside the graph from the buffer's own generator.  The graph is
    re-captured when the number of stored episodes changes (beta is a device scalar)."""

    def __init__(self, trainer, buffer, batch_size, p_id="policy_0"):
        self.trainer, self.buffer, self.batch_size, self.p_id = trainer, buffer, int(batch_size), p_id
        self._graph, self._key, self._info = None, None, None
        self._uniforms = torch.empty(self.batch_size, dtype=torch.float64, device=buffer.device)
        self._beta = torch.zeros((), dtype=torch.float64, device=buffer.device)    # read by the sampling kernel at replay time

    def _step(self, beta):
        sample = self.buffer.sample(self.batch_size, beta=beta, p_id=self.p_id, uniforms=self._uniforms, beta_dev=self._beta)
        info, prio, idx = self.trainer.train_policy_on_batch(sample)
        self.buffer.update_priorities(idx, prio, self.p_id, validate=False)
        self.trainer.soft_target_updates()
        return info

    def __call__(self, beta):
        # beta travels through a device scalar (the annealing schedule does not force a re-capture); the sharded PER normaliser
        # of a multi-GPU run still takes it on the host
        key = (len(self.buffer), float(beta) if getattr(self.buffer, "sharded", False) else None)
        self._beta.fill_(float(beta))
        self._uniforms.copy_(torch.rand(self.batch_size, dtype=torch.float64, device=self.buffer.device, generator=self.buffer._gen))
        if self._graph is None or key != self._key:
            # a real update on a side stream doubles as the warm-up a capture needs (cuBLAS workspaces, lazy kernel loading)
            main = torch.cuda.current_stream()
            side = torch.cuda.Stream(device=self.buffer.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                info = self._step(beta)
            main.wait_stream(side)
            self.trainer.optimizer.zero_grad(set_to_none=True)
            torch.cuda.synchronize()
            self._graph = torch.cuda.CUDAGraph()
            # thread_local: with several GPUs NCCL's watchdog thread queries events while this thread captures (the gradient
            # all-reduce is captured with the rest: one graph per rank, replayed in lockstep)
            with torch.cuda.graph(self._graph, capture_error_mode="thread_local"):
                self._info = self._step(beta)
            self._key = key
            return info
        self._graph.replay()
        # a replay runs no Python: bump the version counters the packed-parameter caches of the rollout kernels key on
        tr = self.trainer
        moved = list(tr.parameters) + [p for pid in tr.policy_ids for p in tr.target_policies[pid].parameters()] + \
            list(tr.target_mixer.parameters())
        torch.autograd.graph.increment_version(moved)
        return self._info


class VDN(QMix):
    def __init__(self, args, num_agents, batch_size, policies, policy_mapping_fn, device=torch.device("cuda:0"),
                 episode_length=None):
        super().__init__(args, num_agents, batch_size, policies, policy_mapping_fn, device, episode_length, vdn=True)
