"""Recurrent MAPPO policy and trainer with the reference's interface, device-resident.

Mirrors ``R_Actor`` / ``R_Critic`` (mappo/algorithms/r_actor_critic.py:13-208), ``R_MAPPOPolicy``
(mappo/algorithms/rmappo_policy.py:7-153), ``R_MAPPO`` (mappo/algorithms/ramppo_network.py:9-295) and
``ValueNorm`` (mappo/utils/valuenorm.py:8-78): same constructor arguments, method names, return tuples and
``state_dict`` keys (``base.feature_norm.*``, ``base.mlp.fc1.0.*``, ``base.mlp.fc2.0.0.*``, ``rnn.rnn.weight_ih_l0``,
``rnn.norm.*``, ``act.action_out.linear.*`` / ``v_out.*``).

Rollout-time calls (``get_actions`` / ``get_values`` / ``act``: one step, no gradient) run the whole
LayerNorm-MLP-GRU-LayerNorm-head chain of actor and critic as ONE fused kernel each
(``mm_agentq_tc_forward`` on tcgen05, or the fp32 ``mm_agentq_forward``); the PPO update
(``evaluate_actions`` over data chunks with mask-segmented BPTT) is the same layer chain in eager PyTorch
so that autograd records it.  Advantages come out of the GAE kernel (``SharedReplayBuffer.compute_returns``).
Multi-GPU: gradients are all-reduced (mean) once per optimiser step, the advantage-normalisation and
ValueNorm batch moments are all-reduced so a sharded run reproduces the single-process formulas.
"""
import numpy as np
import torch
import torch.nn as nn

from ..nn.agent_q import MLPBase, _init_linear, _tc_fits, agentq_forward, pack_agentq, pack_agentq_tc
from ..nn.gru import GRU
from ..nn.layernorm import LayerNorm
from ..nn.ppo_trunk import ChunkRows, direct_param_grads, ppo_trunk, trunk_fits
from ..ops.ppo_loss import ppo_actor_loss, ppo_value_loss, refresh_packs
from ..optim import FlatAdam, clip_and_step, make_adam
from ..parallel import GradBucket, allreduce_mean_grads, allreduce_stats


def _shape(space):
    name = space.__class__.__name__
    if name == "Box":
        return tuple(space.shape)
    if name == "list":
        return tuple(space)
    raise NotImplementedError(name)


def check(x):
    return torch.from_numpy(x) if isinstance(x, np.ndarray) else x


def huber_loss(e, d):
    a = (abs(e) <= d).float()
    b = (abs(e) > d).float()
    return a * e ** 2 / 2 + b * d * (abs(e) - d / 2)


def mse_loss(e):
    return e ** 2 / 2


class ValueNorm(nn.Module):
    """Normalize a vector of observations across the first norm_axes dimensions (running, debiased)."""

    def __init__(self, input_shape, norm_axes=1, beta=0.99999, per_element_update=False, epsilon=1e-5):
        super().__init__()
        self.input_shape, self.norm_axes, self.epsilon, self.beta = input_shape, norm_axes, epsilon, beta
        self.per_element_update = per_element_update
        self.running_mean = nn.Parameter(torch.zeros(input_shape), requires_grad=False)
        self.running_mean_sq = nn.Parameter(torch.zeros(input_shape), requires_grad=False)
        self.debiasing_term = nn.Parameter(torch.tensor(0.0), requires_grad=False)

    def running_mean_var(self):
        debiased_mean = self.running_mean / self.debiasing_term.clamp(min=self.epsilon)
        debiased_mean_sq = self.running_mean_sq / self.debiasing_term.clamp(min=self.epsilon)
        debiased_var = (debiased_mean_sq - debiased_mean ** 2).clamp(min=1e-2)
        return debiased_mean, debiased_var

    @torch.no_grad()
    def update(self, input_vector):
        input_vector = check(input_vector).to(self.running_mean.device)
        n = float(np.prod(input_vector.size()[:self.norm_axes]))
        batch_mean = input_vector.mean(dim=tuple(range(self.norm_axes)))
        batch_sq_mean = (input_vector ** 2).mean(dim=tuple(range(self.norm_axes)))
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            # sharded batch: global moments = all-reduced sums / all-reduced count
            packed = torch.cat([batch_mean.reshape(-1) * n, batch_sq_mean.reshape(-1) * n,
                                torch.full((1,), n, device=batch_mean.device)]).double()      # (no host tensor: graph-capturable)
            packed = allreduce_stats(packed)
            k = batch_mean.numel()
            batch_mean = (packed[:k] / packed[-1]).to(batch_mean.dtype).reshape(batch_mean.shape)
            batch_sq_mean = (packed[k:2 * k] / packed[-1]).to(batch_mean.dtype).reshape(batch_mean.shape)
        weight = self.beta ** n if self.per_element_update else self.beta
        self.running_mean.mul_(weight).add_(batch_mean * (1.0 - weight))
        self.running_mean_sq.mul_(weight).add_(batch_sq_mean * (1.0 - weight))
        self.debiasing_term.mul_(weight).add_(1.0 * (1.0 - weight))

    def normalize(self, input_vector):
        input_vector = check(input_vector).to(self.running_mean.device)
        mean, var = self.running_mean_var()
        return (input_vector - mean[(None,) * self.norm_axes]) / torch.sqrt(var)[(None,) * self.norm_axes]

    def denormalize(self, input_vector):
        """Returns a device tensor (the reference hops to numpy here; the device buffer does not need to)."""
        input_vector = check(input_vector).to(self.running_mean.device)
        mean, var = self.running_mean_var()
        return input_vector * torch.sqrt(var)[(None,) * self.norm_axes] + mean[(None,) * self.norm_axes]


class RNNLayer(nn.Module):
    """nn.GRU + LayerNorm with the reference's mask handling (mappo/utils/algorithm_utils/rnn.py:7-80)."""

    def __init__(self, inputs_dim, outputs_dim, recurrent_N, use_orthogonal):
        super().__init__()
        self._recurrent_N = recurrent_N
        self.rnn = GRU(inputs_dim, outputs_dim, num_layers=recurrent_N)
        for name, param in self.rnn.named_parameters():
            if "bias" in name:
                nn.init.constant_(param, 0)
            elif "weight" in name:
                (nn.init.orthogonal_ if use_orthogonal else nn.init.xavier_uniform_)(param)
        self.norm = LayerNorm(outputs_dim)

    def forward(self, x, hxs, masks):
        if x.size(0) == hxs.size(0):
            x, hxs = self.rnn(x.unsqueeze(0), (hxs * masks.repeat(1, self._recurrent_N).unsqueeze(-1)).transpose(0, 1).contiguous())
            x = x.squeeze(0)
            hxs = hxs.transpose(0, 1)
        else:
            # x is a (T, N, -1) tensor flattened to (T * N, -1): run the GRU over the maximal runs of steps in
            # which no env of the minibatch resets its hidden state
            N = hxs.size(0)
            T = int(x.size(0) / N)
            x = x.view(T, N, x.size(1))
            masks = masks.view(T, N)
            resets = ((masks[1:] == 0.0).any(dim=-1).nonzero().flatten() + 1).tolist()
            bounds = [0] + resets + [T]
            hxs = hxs.transpose(0, 1)
            outputs = []
            for start, end in zip(bounds[:-1], bounds[1:]):
                temp = (hxs * masks[start].view(1, -1, 1).repeat(self._recurrent_N, 1, 1)).contiguous()
                scores, hxs = self.rnn(x[start:end], temp)
                outputs.append(scores)
            x = torch.cat(outputs, dim=0).reshape(T * N, -1)
            hxs = hxs.transpose(0, 1)
        return self.norm(x), hxs


class Categorical(nn.Module):
    def __init__(self, num_inputs, num_outputs, use_orthogonal=True, gain=0.01):
        super().__init__()
        self.linear = _init_linear(nn.Linear(num_inputs, num_outputs), use_orthogonal, gain)

    def forward(self, x, available_actions=None):
        x = self.linear(x)
        if available_actions is not None:
            x[available_actions == 0] = -1e10
        return torch.distributions.Categorical(logits=x)


class ACTLayer(nn.Module):
    def __init__(self, action_space, inputs_dim, use_orthogonal, gain):
        super().__init__()
        if action_space.__class__.__name__ != "Discrete":
            raise NotImplementedError("ma-gym action spaces are Discrete")
        self.action_out = Categorical(inputs_dim, action_space.n, use_orthogonal, gain)

    def forward(self, x, available_actions=None, deterministic=False):
        dist = self.action_out(x, available_actions)
        actions = dist.probs.argmax(dim=-1, keepdim=True) if deterministic else dist.sample().unsqueeze(-1)
        return actions, dist.log_prob(actions.squeeze(-1)).unsqueeze(-1)

    def evaluate_actions(self, x, action, available_actions=None, active_masks=None):
        dist = self.action_out(x, available_actions)
        action_log_probs = dist.log_prob(action.squeeze(-1)).view(action.size(0), -1).sum(-1).unsqueeze(-1)
        if active_masks is not None:
            dist_entropy = (dist.entropy() * active_masks.squeeze(-1)).sum() / active_masks.sum()
        else:
            dist_entropy = dist.entropy().mean()
        return action_log_probs, dist_entropy


_KEYMAP = [("base.feature_norm", "rnn.feature_norm"), ("base.mlp.fc1", "rnn.mlp.fc1"), ("base.mlp.fc2", "rnn.mlp.fc2"),
           ("rnn.rnn", "rnn.rnn.rnn"), ("rnn.norm", "rnn.rnn.norm")]


class _FusedTrunk(nn.Module):
    """Shared by actor and critic: base + rnn + a Linear head, with the one-step fused forward."""

    def _setup(self, args, obs_shape, head_key, n_out, device):
        self.hidden_size = args.hidden_size
        self._recurrent_N = args.recurrent_N
        self.tpdv = dict(dtype=torch.float32, device=device)
        self._dev = torch.device(device)
        self._in, self._n_out, self._head_key = int(obs_shape[0]), n_out, head_key
        self._packed, self._key = None, None
        self._fusable = (self._dev.type == "cuda" and args.layer_N == 1 and args.recurrent_N == 1 and
                         args.use_feature_normalization and args.use_ReLU and args.hidden_size in (32, 64) and
                         self._in <= 96 and (args.use_recurrent_policy or args.use_naive_recurrent_policy))
        self._precision = getattr(args, "fused_precision", "tf32")
        self._fused_update = bool(getattr(args, "fused_update", True))

    def _chunk_fusable(self, obs, rnn_states):
        """The differentiated chunk pass runs as the fused trunk op: default network shape, tf32, CUDA, TF32 matmuls on."""
        return (self._fusable and self._precision == "tf32" and trunk_fits(self._in, self.hidden_size, self._n_out)
                and getattr(self, "_fused_update", True) and torch.backends.cuda.matmul.allow_tf32
                and obs.size(0) > rnn_states.size(0) and obs.size(0) % rnn_states.size(0) == 0)

    def _trunk_params(self):
        sd = dict(self.named_parameters())
        return dict(ln0_w=sd["base.feature_norm.weight"], ln0_b=sd["base.feature_norm.bias"],
                    w1=sd["base.mlp.fc1.0.weight"], b1=sd["base.mlp.fc1.0.bias"],
                    ln1_w=sd["base.mlp.fc1.2.weight"], ln1_b=sd["base.mlp.fc1.2.bias"],
                    w2=sd["base.mlp.fc2.0.0.weight"], b2=sd["base.mlp.fc2.0.0.bias"],
                    ln2_w=sd["base.mlp.fc2.0.2.weight"], ln2_b=sd["base.mlp.fc2.0.2.bias"],
                    wih=sd["rnn.rnn.weight_ih_l0"], whh=sd["rnn.rnn.weight_hh_l0"], bih=sd["rnn.rnn.bias_ih_l0"],
                    bhh=sd["rnn.rnn.bias_hh_l0"], ln3_w=sd["rnn.norm.weight"], ln3_b=sd["rnn.norm.bias"],
                    wo=sd[self._head_key + ".weight"], bo=sd[self._head_key + ".bias"])

    def _fused_chunks(self, obs, rnn_states, masks):
        """obs [L*N, D] (step-major), rnn_states [N, 1, H], masks [L*N, 1] -> head output [L*N, n_out]."""
        N = rnn_states.size(0)
        Lc = obs.size(0) // N
        if isinstance(obs, ChunkRows):                 # chunks addressed in place in the rollout buffer (all three)
            out = ppo_trunk(obs, rnn_states, masks, self._trunk_params())
        else:
            out = ppo_trunk(obs.view(Lc, N, -1), rnn_states[:, 0, :], masks.view(Lc, N), self._trunk_params())
        return out.view(Lc * N, -1)

    def _agentq_state(self):
        sd = self.state_dict()
        out = {}
        for k, v in sd.items():
            for a, b in _KEYMAP:
                if k.startswith(a + "."):
                    out[b + k[len(a):]] = v
                    break
        out["q.action_out.weight"] = sd[self._head_key + ".weight"]
        out["q.action_out.bias"] = sd[self._head_key + ".bias"]
        return out

    def _fused_step(self, obs, rnn_states, masks):
        """One step for all rows: returns (head output [rows, n_out], new rnn_states [rows, 1, H])."""
        tc = self._precision == "tf32" and _tc_fits(self._in, self.hidden_size, self._n_out)
        key = (tc,) + tuple(p._version for p in self.parameters()) + tuple(p.data_ptr() for p in self.parameters())
        if self._packed is None or key != self._key:
            pack = pack_agentq_tc if tc else pack_agentq
            self._packed, self._key = pack(self._agentq_state(), self._in, self.hidden_size, self._n_out, self._dev), key
        h0 = (rnn_states[:, 0, :] * masks).contiguous()
        out, h, _ = agentq_forward(self._packed, obs.contiguous(), h0, self._in, self.hidden_size, self._n_out, tensor_cores=tc)
        return out, h.unsqueeze(1)


class R_Actor(_FusedTrunk):
    def __init__(self, args, obs_space, action_space, device=torch.device("cpu")):
        super().__init__()
        self._use_policy_active_masks = args.use_policy_active_masks
        self._use_rnn = args.use_naive_recurrent_policy or args.use_recurrent_policy
        obs_shape = _shape(obs_space)
        self.base = MLPBase(args, obs_shape[0])
        if self._use_rnn:
            self.rnn = RNNLayer(args.hidden_size, args.hidden_size, args.recurrent_N, args.use_orthogonal)
        self.act = ACTLayer(action_space, args.hidden_size, args.use_orthogonal, args.gain)
        self.to(device)
        self._setup(args, obs_shape, "act.action_out.linear", action_space.n, device)

    def _logits_step(self, obs, rnn_states, masks):
        return self._fused_step(obs, rnn_states, masks)

    def forward(self, obs, rnn_states, masks, available_actions=None, deterministic=False):
        obs = check(obs).to(**self.tpdv)
        rnn_states = check(rnn_states).to(**self.tpdv)
        masks = check(masks).to(**self.tpdv)
        if self._fusable and not torch.is_grad_enabled() and obs.size(0) == rnn_states.size(0) and available_actions is None:
            logits, rnn_states = self._fused_step(obs, rnn_states, masks)
            dist = torch.distributions.Categorical(logits=logits)
            actions = dist.probs.argmax(dim=-1, keepdim=True) if deterministic else dist.sample().unsqueeze(-1)
            return actions, dist.log_prob(actions.squeeze(-1)).unsqueeze(-1), rnn_states
        actor_features = self.base(obs)
        if self._use_rnn:
            actor_features, rnn_states = self.rnn(actor_features, rnn_states, masks)
        actions, action_log_probs = self.act(actor_features, available_actions, deterministic)
        return actions, action_log_probs, rnn_states

    def evaluate_actions(self, obs, rnn_states, action, masks, available_actions=None, active_masks=None):
        obs = check(obs).to(**self.tpdv)
        rnn_states = check(rnn_states).to(**self.tpdv)
        action = check(action).to(**self.tpdv)
        masks = check(masks).to(**self.tpdv)
        if active_masks is not None:
            active_masks = check(active_masks).to(**self.tpdv)
        if available_actions is None and self._chunk_fusable(obs, rnn_states):
            logits = self._fused_chunks(obs, rnn_states, masks)
            if getattr(self, "_want_logits", False):
                return logits, None                  # R_MAPPO's fused loss kernel takes the logits themselves
            dist = torch.distributions.Categorical(logits=logits)
            action_log_probs = dist.log_prob(action.squeeze(-1)).view(action.size(0), -1).sum(-1).unsqueeze(-1)
            if active_masks is not None and self._use_policy_active_masks:
                dist_entropy = (dist.entropy() * active_masks.squeeze(-1)).sum() / active_masks.sum()
            else:
                dist_entropy = dist.entropy().mean()
            return action_log_probs, dist_entropy
        actor_features = self.base(obs)
        if self._use_rnn:
            actor_features, rnn_states = self.rnn(actor_features, rnn_states, masks)
        return self.act.evaluate_actions(actor_features, action, available_actions,
                                         active_masks=active_masks if self._use_policy_active_masks else None)


class R_Critic(_FusedTrunk):
    def __init__(self, args, cent_obs_space, device=torch.device("cpu")):
        super().__init__()
        if args.use_popart:
            raise NotImplementedError("PopArt is outside the hot path (use_popart defaults to False)")
        self._use_rnn = args.use_naive_recurrent_policy or args.use_recurrent_policy
        cent_obs_shape = _shape(cent_obs_space)
        self.base = MLPBase(args, cent_obs_shape[0])
        if self._use_rnn:
            self.rnn = RNNLayer(args.hidden_size, args.hidden_size, args.recurrent_N, args.use_orthogonal)
        self.v_out = _init_linear(nn.Linear(args.hidden_size, 1), args.use_orthogonal, 1)
        self.to(device)
        self._setup(args, cent_obs_shape, "v_out", 1, device)

    def forward(self, cent_obs, rnn_states, masks):
        cent_obs = check(cent_obs).to(**self.tpdv)
        rnn_states = check(rnn_states).to(**self.tpdv)
        masks = check(masks).to(**self.tpdv)
        if self._fusable and not torch.is_grad_enabled() and cent_obs.size(0) == rnn_states.size(0):
            return self._fused_step(cent_obs, rnn_states, masks)
        if torch.is_grad_enabled() and self._chunk_fusable(cent_obs, rnn_states):
            return self._fused_chunks(cent_obs, rnn_states, masks), None
        critic_features = self.base(cent_obs)
        if self._use_rnn:
            critic_features, rnn_states = self.rnn(critic_features, rnn_states, masks)
        return self.v_out(critic_features), rnn_states


class R_MAPPOPolicy:
    def __init__(self, args, obs_space, cent_obs_space, act_space, device):
        self.device = device
        self.actor_lr, self.critic_lr = args.actor_lr, args.critic_lr
        self.opti_eps, self.weight_decay = args.opti_eps, args.weight_decay
        self.obs_space, self.share_obs_space, self.act_space = obs_space, cent_obs_space, act_space
        self.actor = R_Actor(args, obs_space, act_space, device)
        self.critic = R_Critic(args, cent_obs_space, device)
        # FlatAdam on the GPU (optim.py: clip + Adam as two launches on flat buffers), torch.optim.Adam for host parameters
        self.actor_optimizer = make_adam(self.actor.parameters(), lr=self.actor_lr, eps=self.opti_eps,
                                         weight_decay=self.weight_decay)
        self.critic_optimizer = make_adam(self.critic.parameters(), lr=self.critic_lr, eps=self.opti_eps,
                                          weight_decay=self.weight_decay)

    def lr_decay(self, episode, episodes):
        for opt, lr0 in ((self.actor_optimizer, self.actor_lr), (self.critic_optimizer, self.critic_lr)):
            lr = lr0 - (lr0 * (episode / float(episodes)))
            for group in opt.param_groups:
                group["lr"] = lr

    def get_actions(self, obs, cent_obs, rnn_states_actor, rnn_states_critic, masks, available_actions=None,
                    deterministic=False):
        actions, action_log_probs, rnn_states_actor = self.actor(obs, rnn_states_actor, masks, available_actions=None,
                                                                 deterministic=False)
        values, rnn_states_critic = self.critic(cent_obs, rnn_states_critic, masks)
        return values, actions, action_log_probs, rnn_states_actor, rnn_states_critic

    def get_values(self, cent_obs, rnn_states_critic, masks):
        values, _ = self.critic(cent_obs, rnn_states_critic, masks)
        return values

    def evaluate_actions(self, cent_obs, obs, rnn_states_actor, rnn_states_critic, action, masks, available_actions=None,
                         active_masks=None):
        action_log_probs, dist_entropy = self.actor.evaluate_actions(obs, rnn_states_actor, action, masks,
                                                                     available_actions, active_masks)
        values, _ = self.critic(cent_obs, rnn_states_critic, masks)
        return values, action_log_probs, dist_entropy

    def act(self, obs, rnn_states_actor, masks, available_actions=None, deterministic=False):
        actions, _, rnn_states_actor = self.actor(obs, rnn_states_actor, masks, available_actions, deterministic)
        return actions, rnn_states_actor


class R_MAPPO:
    def __init__(self, args, policy, device):
        self._use_popart = args.use_popart
        self._use_valuenorm = args.use_valuenorm
        self._use_huber_loss = args.use_huber_loss
        self._use_max_grad_norm = args.use_max_grad_norm
        self._use_recurrent_policy = args.use_recurrent_policy
        self._use_clipped_value_loss = args.use_clipped_value_loss
        self._use_value_active_masks = args.use_value_active_masks
        self._use_naive_recurrent = args.use_naive_recurrent_policy
        self._use_policy_active_masks = args.use_policy_active_masks
        self.ppo_epoch, self.train_batch_size = args.ppo_epoch, args.train_batch_size
        self.data_chunk_length = args.data_chunk_length
        self.clip_param, self.huber_delta = args.clip_param, args.huber_delta
        self.entropy_coef, self.max_grad_norm = args.entropy_coef, args.max_grad_norm
        self.value_loss_coef = args.value_loss_coef
        self.device = device
        self.tpdv = dict(dtype=torch.float32, device=device)
        self.policy = policy
        assert (self._use_popart and self._use_valuenorm) is False
        if self._use_popart:
            raise NotImplementedError("PopArt is outside the hot path")
        self.value_normalizer = ValueNorm(input_shape=1).to(self.device) if self._use_valuenorm else None
        self._fused_loss = bool(getattr(args, "fused_update", True)) and torch.device(device).type == "cuda"
        self._buckets = {}
        # one PPO epoch (row map, both networks' fused passes, losses, backward, clip + Adam: ~70 launches) replayed as a CUDA graph
        self._train_graph = bool(getattr(args, "train_graph", True)) and self._fused_loss
        self._graph_state = None

    def cal_value_loss(self, values, value_preds_batch, return_batch, active_masks_batch):
        if self._fused_loss and values.is_cuda:
            norm = None
            if self._use_valuenorm:
                self.value_normalizer.update(return_batch.gather() if isinstance(return_batch, ChunkRows) else return_batch)
                mean, var = self.value_normalizer.running_mean_var()
                norm = torch.cat([mean.reshape(1), torch.rsqrt(var).reshape(1)]).to(torch.float32)
            return ppo_value_loss(values, value_preds_batch, return_batch,
                                  active_masks_batch if self._use_value_active_masks else None, norm, self.clip_param,
                                  self.huber_delta, self._use_huber_loss, self._use_clipped_value_loss)
        value_pred_clipped = value_preds_batch + (values - value_preds_batch).clamp(-self.clip_param, self.clip_param)
        if self._use_valuenorm:
            self.value_normalizer.update(return_batch)
            error_clipped = self.value_normalizer.normalize(return_batch) - value_pred_clipped
            error_original = self.value_normalizer.normalize(return_batch) - values
        else:
            error_clipped = return_batch - value_pred_clipped
            error_original = return_batch - values
        loss_fn = (lambda e: huber_loss(e, self.huber_delta)) if self._use_huber_loss else mse_loss
        value_loss_clipped, value_loss_original = loss_fn(error_clipped), loss_fn(error_original)
        value_loss = torch.max(value_loss_original, value_loss_clipped) if self._use_clipped_value_loss else value_loss_original
        if self._use_value_active_masks:
            return (value_loss * active_masks_batch).sum() / active_masks_batch.sum()
        return value_loss.mean()

    def ppo_update(self, sample, update_actor=True):
        (share_obs_batch, obs_batch, rnn_states_batch, rnn_states_critic_batch, actions_batch, value_preds_batch,
         return_batch, masks_batch, active_masks_batch, old_action_log_probs_batch, adv_targ, available_actions_batch) = sample
        to = lambda x: x if isinstance(x, ChunkRows) else check(x).to(**self.tpdv)
        old_action_log_probs_batch, adv_targ = to(old_action_log_probs_batch), to(adv_targ)
        value_preds_batch, return_batch, active_masks_batch = to(value_preds_batch), to(return_batch), to(active_masks_batch)
        actor = self.policy.actor
        if (self._fused_loss and available_actions_batch is None and hasattr(actor, "_chunk_fusable")
                and actor._chunk_fusable(to(obs_batch), to(rnn_states_batch))):
            return self._ppo_update_fused(sample, update_actor)
        if isinstance(obs_batch, ChunkRows):             # eager path: materialise the minibatch the generator only addressed
            (share_obs_batch, obs_batch, rnn_states_batch, rnn_states_critic_batch, actions_batch, value_preds_batch,
             return_batch, masks_batch, active_masks_batch, old_action_log_probs_batch, adv_targ) = \
                [x.gather() for x in sample[:11]]
        values, action_log_probs, dist_entropy = self.policy.evaluate_actions(
            share_obs_batch, obs_batch, rnn_states_batch, rnn_states_critic_batch, actions_batch, masks_batch,
            available_actions_batch, active_masks_batch)
        imp_weights = torch.exp(action_log_probs - old_action_log_probs_batch)
        surr1 = imp_weights * adv_targ
        surr2 = torch.clamp(imp_weights, 1.0 - self.clip_param, 1.0 + self.clip_param) * adv_targ
        if self._use_policy_active_masks:
            policy_action_loss = (-torch.sum(torch.min(surr1, surr2), dim=-1, keepdim=True) * active_masks_batch).sum() \
                / active_masks_batch.sum()
        else:
            policy_action_loss = -torch.sum(torch.min(surr1, surr2), dim=-1, keepdim=True).mean()
        policy_loss = policy_action_loss
        self._zero_grads(self.policy.actor, self.policy.actor_optimizer)
        if update_actor:
            (policy_loss - dist_entropy * self.entropy_coef).backward()
        allreduce_mean_grads(list(self.policy.actor.parameters()), bucket=self._buckets.get(id(self.policy.actor)))
        actor_grad_norm = clip_and_step(self.policy.actor_optimizer, list(self.policy.actor.parameters()),
                                         self.max_grad_norm if self._use_max_grad_norm else 0.0)

        value_loss = self.cal_value_loss(values, value_preds_batch, return_batch, active_masks_batch)
        self._zero_grads(self.policy.critic, self.policy.critic_optimizer)
        (value_loss * self.value_loss_coef).backward()
        allreduce_mean_grads(list(self.policy.critic.parameters()), bucket=self._buckets.get(id(self.policy.critic)))
        critic_grad_norm = clip_and_step(self.policy.critic_optimizer, list(self.policy.critic.parameters()),
                                         self.max_grad_norm if self._use_max_grad_norm else 0.0)
        return value_loss, critic_grad_norm, policy_loss, dist_entropy, actor_grad_norm, imp_weights

    def _zero_grads(self, module, optimizer):
        """Multi-GPU: the module's gradients live in one flat ``GradBucket`` (``p.grad`` are views; backward and the fused
        trunk op write straight into it, the exchange is one all-reduce on the flat buffer), zeroed here instead of
        ``optimizer.zero_grad()`` so that the views survive.  Single GPU: plain ``zero_grad``."""
        import torch.distributed as dist
        if isinstance(optimizer, FlatAdam):
            self._buckets[id(module)] = optimizer.bucket
            optimizer.zero_grad()
        elif dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            b = self._buckets.get(id(module))
            if b is None or not b.attached():
                b = self._buckets[id(module)] = GradBucket(list(module.parameters()))
            b.zero()
        else:
            optimizer.zero_grad()

    def _ppo_update_fused(self, sample, update_actor=True):
        """Same update with the fused trunk op and the loss kernels: logits -> (policy loss, entropy, ratio) in one pass."""
        (share_obs_batch, obs_batch, rnn_states_batch, rnn_states_critic_batch, actions_batch, value_preds_batch,
         return_batch, masks_batch, active_masks_batch, old_action_log_probs_batch, adv_targ, _) = sample
        to = lambda x: x if isinstance(x, ChunkRows) else check(x).to(**self.tpdv)
        actor, critic = self.policy.actor, self.policy.critic
        logits = actor._fused_chunks(to(obs_batch), to(rnn_states_batch), to(masks_batch))
        policy_loss, dist_entropy, ratio = ppo_actor_loss(
            logits, to(actions_batch), to(old_action_log_probs_batch), to(adv_targ),
            to(active_masks_batch) if self._use_policy_active_masks else None, self.clip_param)
        self._zero_grads(self.policy.actor, self.policy.actor_optimizer)
        if update_actor:
            with direct_param_grads():                   # the trunk op adds its 18 gradients into the bucket views itself
                (policy_loss - dist_entropy * self.entropy_coef).backward()
        allreduce_mean_grads(list(actor.parameters()), bucket=self._buckets.get(id(actor)))
        actor_grad_norm = clip_and_step(self.policy.actor_optimizer, list(actor.parameters()),
                                         self.max_grad_norm if self._use_max_grad_norm else 0.0)

        if critic._chunk_fusable(to(share_obs_batch), to(rnn_states_critic_batch)):
            values = critic._fused_chunks(to(share_obs_batch), to(rnn_states_critic_batch), to(masks_batch))
        else:
            g = lambda x: x.gather() if isinstance(x, ChunkRows) else x
            values, _ = critic(g(share_obs_batch), g(rnn_states_critic_batch), g(masks_batch))
        value_loss = self.cal_value_loss(values, to(value_preds_batch), to(return_batch), to(active_masks_batch))
        self._zero_grads(self.policy.critic, self.policy.critic_optimizer)
        with direct_param_grads():
            (value_loss * self.value_loss_coef).backward()
        allreduce_mean_grads(list(critic.parameters()), bucket=self._buckets.get(id(critic)))
        critic_grad_norm = clip_and_step(self.policy.critic_optimizer, list(critic.parameters()),
                                         self.max_grad_norm if self._use_max_grad_norm else 0.0)
        return value_loss, critic_grad_norm, policy_loss, dist_entropy, actor_grad_norm, ratio

    def train(self, buffer, update_actor=True, perms=None):
        """``perms`` (optional list, one permutation per PPO epoch) replaces the internal randperm (tests)."""
        if self._use_valuenorm:
            advantages = buffer.returns[:-1] - self.value_normalizer.denormalize(buffer.value_preds[:-1])
        else:
            advantages = buffer.returns[:-1] - buffer.value_preds[:-1]
        active = buffer.active_masks[:-1] != 0.0
        # nanmean / nanstd over active entries == masked moments; all-reduced so shards normalise identically
        a64 = advantages.double()
        stats = allreduce_stats(torch.stack([(a64 * active).sum(), (a64 * a64 * active).sum(), active.sum().double()]))
        mean_advantages = stats[0] / stats[2]
        std_advantages = torch.sqrt((stats[1] / stats[2] - mean_advantages ** 2).clamp(min=0))
        advantages = ((advantages - mean_advantages) / (std_advantages + 1e-5)).to(torch.float32)

        if self._graphable(buffer, perms):
            return self._train_graphed(buffer, advantages, update_actor)
        train_info = dict(value_loss=0, policy_loss=0, dist_entropy=0, actor_grad_norm=0, critic_grad_norm=0, ratio=0)
        tables = {}                    # normalised observation tables of this update (built on first use by the fused trunk)
        for epoch in range(self.ppo_epoch):
            perm = None if perms is None else perms[epoch]
            if self._use_recurrent_policy:
                gen = buffer.recurrent_generator(advantages, self.train_batch_size, self.data_chunk_length, perm=perm,
                                                 lazy_obs=self._fused_loss, tables=tables)
            elif self._use_naive_recurrent:
                gen = buffer.naive_recurrent_generator(advantages, self.train_batch_size, perm=perm)
            else:
                gen = buffer.feed_forward_generator(advantages, self.train_batch_size, perm=perm)
            for sample in gen:
                value_loss, critic_grad_norm, policy_loss, dist_entropy, actor_grad_norm, imp_weights = \
                    self.ppo_update(sample, update_actor)
                train_info["value_loss"] += value_loss.detach()
                train_info["policy_loss"] += policy_loss.detach()
                train_info["dist_entropy"] += dist_entropy.detach()
                train_info["actor_grad_norm"] += actor_grad_norm
                train_info["critic_grad_norm"] += critic_grad_norm
                train_info["ratio"] += imp_weights.detach().mean()
        num_updates = self.ppo_epoch * self.train_batch_size
        for k in train_info.keys():
            train_info[k] /= num_updates
        train_info["Sampling_Rewards"] = buffer.rewards.sum()
        return train_info

    # -- the PPO epochs as a replayed CUDA graph -----------------------------------------------------------------------------
    def _graphable(self, buffer, perms):
        """The launch-bound epoch loop (2 200 launches per update at the bench size, ~50 ms of host enqueue) is replayed from a
        graph when everything it touches is static storage: the fused chunk path for both networks, FlatAdam, the device
        permutation, no caller-supplied permutations.  With several GPUs the gradient / ValueNorm all-reduces (NCCL) are captured
        with the rest: every rank replays its own graph in lockstep."""
        if not (self._train_graph and perms is None and self._use_recurrent_policy and torch.is_tensor(buffer.obs) and buffer.obs.is_cuda):
            return False
        T, L = buffer.rewards.shape[0], self.data_chunk_length
        pol = self.policy
        fits = lambda net, x: hasattr(net, "_chunk_fusable") and net._fusable and net._precision == "tf32" and \
            trunk_fits(x.shape[-1], net.hidden_size, net._n_out) and getattr(net, "_fused_update", True) and L > 1
        return (T % L == 0 and torch.backends.cuda.matmul.allow_tf32 and fits(pol.actor, buffer.obs) and fits(pol.critic, buffer.share_obs)
                and isinstance(pol.actor_optimizer, FlatAdam) and isinstance(pol.critic_optimizer, FlatAdam))

    def _epoch(self, buffer, st, update_actor):
        gen = buffer.recurrent_generator(st["adv"], self.train_batch_size, self.data_chunk_length, perm=st["perm"], lazy_obs=True,
                                         tables=st["tables"])
        for sample in gen:
            value_loss, critic_grad_norm, policy_loss, dist_entropy, actor_grad_norm, ratio = self.ppo_update(sample, update_actor)
            st["acc"] += torch.stack([value_loss.detach(), policy_loss.detach(), dist_entropy.detach(), actor_grad_norm.reshape(()),
                                      critic_grad_norm.reshape(()), ratio.detach().mean()])

    def _train_graphed(self, buffer, advantages, update_actor):
        from ..nn.ppo_trunk import obs_table
        T, N, A = buffer.rewards.shape[0:3]
        dev = buffer.obs.device
        key = (buffer.obs.data_ptr(), buffer.share_obs.data_ptr(), tuple(buffer.obs.shape), tuple(buffer.share_obs.shape),
               bool(update_actor), self.ppo_epoch, self.train_batch_size, self.data_chunk_length)
        st = self._graph_state
        fresh = st is None or st["key"] != key
        if fresh:
            st = self._graph_state = dict(key=key, adv=torch.empty_like(advantages), tables={}, graph=None,
                                          perm=torch.empty(N * T * A // self.data_chunk_length, dtype=torch.int64, device=dev),
                                          acc=torch.zeros(6, dtype=torch.float32, device=dev))
        st["adv"].copy_(advantages)
        st["acc"].zero_()
        masks = buffer.masks.reshape(-1)
        for k, src in (("obs", buffer.obs), ("share_obs", buffer.share_obs)):       # this update's normalised fp16 tables, in place
            if k in st["tables"]:
                obs_table(src.reshape(-1, src.shape[-1]), masks, out=st["tables"][k])
        refresh_packs(st["tables"])                       # the interleaved loss-input tables (ops/ppo_loss.py), in place
        for opt in (self.policy.actor_optimizer, self.policy.critic_optimizer):
            opt.sync_lr()
        # the epochs' permutations (same generator calls in the same order as the eager loop) are drawn on a second stream while
        # the previous epochs replay: the radix sorts behind randperm are ~1.1 ms per update that nothing waits for any more
        main = torch.cuda.current_stream(dev)
        if "perm_stream" not in st:
            st["perm_stream"] = torch.cuda.Stream(device=dev)
            st["perms"] = torch.empty((self.ppo_epoch, st["perm"].numel()), dtype=torch.int64, device=dev)
            st["perm_events"] = [torch.cuda.Event() for _ in range(self.ppo_epoch)]
        ps = st["perm_stream"]
        ps.wait_stream(main)                              # the previous update's reads of ``perms`` are enqueued on main
        with torch.cuda.stream(ps):
            for epoch in range(self.ppo_epoch):
                st["perms"][epoch].copy_(torch.randperm(st["perm"].numel(), device=dev))
                st["perm_events"][epoch].record(ps)
        for epoch in range(self.ppo_epoch):
            main.wait_event(st["perm_events"][epoch])
            st["perm"].copy_(st["perms"][epoch])
            if st["graph"] is not None:
                st["graph"].replay()
                for opt in (self.policy.actor_optimizer, self.policy.critic_optimizer):
                    opt.bump_versions()                  # a replay runs no Python; the rollout's packed parameters key on versions
                continue
            # first epoch with this buffer: a real epoch on a side stream (builds the tables, warms the allocator), then the capture
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self._epoch(buffer, st, update_actor)
            main.wait_stream(side)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):      # NCCL's watchdog thread may query events meanwhile
                self._epoch(buffer, st, update_actor)
            st["graph"] = graph
        vals = st["acc"] / (self.ppo_epoch * self.train_batch_size)
        train_info = dict(zip(("value_loss", "policy_loss", "dist_entropy", "actor_grad_norm", "critic_grad_norm", "ratio"), vals.unbind(0)))
        train_info["Sampling_Rewards"] = buffer.rewards.sum()
        return train_info

    def prep_training(self):
        self.policy.actor.train()
        self.policy.critic.train()

    def prep_rollout(self):
        self.policy.actor.eval()
        self.policy.critic.eval()
