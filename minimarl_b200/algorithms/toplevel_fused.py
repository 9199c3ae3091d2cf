"""Python side of the fused kernels behind the top-level ``vdn/`` and ``qmix/`` packages (``csrc/toplevel_qnet.cu``,
``csrc/toplevel_mix.cu``): pointer tables of the reference-layout parameters, one call per kernel.

The kernels read ``nn.Parameter`` storage directly (a host array of device pointers, staged into shared memory inside the
kernel) and write gradients straight into ``p.grad``, so ``Q_Net`` / ``Mix_Net`` stay ordinary modules with the reference's
``state_dict`` keys, optimisers and ``load_state_dict`` target copies work unchanged, and nothing is re-packed after an
optimiser step.  Reference: ``vdn/_network.py:52-91``, ``vdn/_train.py:56-101``, ``qmix/_network.py:172-220``,
``qmix/_train.py:19-121``, ``vdn/main.py:139-177``.
"""
import ctypes as C

import torch

from .. import _lib

ENV_CHECKERS, ENV_SWITCH = 1, 2
MAX_D, MAX_S, MAX_NA = 96, 96, 8


def qnet_tensors(net):
    """The 20 parameter tensors of a 2-agent recurrent ``Q_Net`` in kernel order (agent 0: w1 b1 w2 b2 wih whh bih bhh wo bo,
    then agent 1)."""
    out = []
    for i in range(2):
        f, g, a = net.feature_net[i], net.gru_net[i], net.action_net[i][0]
        out += [f[0].weight, f[0].bias, f[2].weight, f[2].bias, g.weight_ih, g.weight_hh, g.bias_ih, g.bias_hh, a.weight, a.bias]
    return out


def mix_tensors(mix):
    b2 = mix.hyper_net_bias_2
    return [mix.gru.weight_ih, mix.gru.weight_hh, mix.gru.bias_ih, mix.gru.bias_hh, mix.hyper_net_weight_1.weight,
            mix.hyper_net_weight_1.bias, mix.hyper_net_weight_2.weight, mix.hyper_net_weight_2.bias,
            mix.hyper_net_bias_1.weight, mix.hyper_net_bias_1.bias, b2[0].weight, b2[0].bias, b2[2].weight, b2[2].bias]


def qnet_supported(net):
    """The fused kernels cover what both top-level packages run by default: two agents, recurrent, equal observation sizes."""
    try:
        if net.num_agents != 2 or not net.use_recurrent or len(net.gru_net) != 2:
            return False
        d = [s.shape[0] for s in net.observation_space]
        na = [s.n for s in net.action_space]
        p = next(net.parameters())
        return d[0] == d[1] and d[0] <= MAX_D and na[0] == na[1] and na[0] <= MAX_NA and p.is_cuda and p.dtype == torch.float32
    except (AttributeError, IndexError, StopIteration):
        return False


def mix_supported(mix):
    try:
        p = next(mix.parameters())
        return (mix.recurrent and mix.n_agents == 2 and mix.gru.weight_ih.shape[1] <= MAX_S and mix.hidden_dim == 32
                and mix.hx_size == 32 and p.is_cuda and p.dtype == torch.float32)
    except (AttributeError, StopIteration):
        return False


class _Table:
    """Host array of device pointers; keeps the tensors alive and re-reads the addresses when a tensor was replaced."""

    def __init__(self, tensors):
        self.tensors = list(tensors)
        for t in self.tensors:
            if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
                raise RuntimeError("minimarl_b200 fused kernels take contiguous float32 CUDA parameters (no CPU fallback)")
        self.array = (C.c_void_p * len(self.tensors))(*[t.data_ptr() for t in self.tensors])


def _params(module, tensors_fn):
    tab = module.__dict__.get("_mm_ptrs")
    ts = tensors_fn(module)
    if tab is None or any(a.data_ptr() != b for a, b in zip(ts, tab.array)):
        tab = _Table(ts)
        module.__dict__["_mm_ptrs"] = tab
    return tab


def _grads(module, tensors_fn):
    """Pointer table of the ``.grad`` tensors (created as zeros when missing; never replaced afterwards, so a captured CUDA
    graph keeps writing into the storage the optimiser reads)."""
    ts = tensors_fn(module)
    for p in ts:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    tab = module.__dict__.get("_mm_grad_ptrs")
    gs = [p.grad for p in ts]
    if tab is None or any(a.data_ptr() != b for a, b in zip(gs, tab.array)):
        tab = _Table(gs)
        module.__dict__["_mm_grad_ptrs"] = tab
    return tab


def _dims(net):
    return net.observation_space[0].shape[0], net.action_space[0].n


def rollout_episode(net, env, T, L, eps, seed=0, step_offset=0, forced_actions=None, out=None):
    """``T`` lockstep steps of every env of ``env`` (reset first) under the behaviour ``Q_Net`` in one launch.
    Returns a dict of chunk-layout tensors: states / next_states [C,N,L,A,D], actions / rewards [C,N,L,A], dones [C,N,L],
    behv [T,N], alive [T,N] (uint8), score [N]."""
    from ..envs.batched import BatchedCheckers
    D, NA = _dims(net)
    N, A, dev = env.n_envs, env.n_agents, env.device
    if T % L or A != 2 or D != env.obs_dim:
        raise RuntimeError("fused top-level rollout: T must be a multiple of chunk_size, 2 agents, matching obs size")
    Cn = T // L
    if out is None:
        z = dict(dtype=torch.float32, device=dev)
        out = dict(states=torch.empty((Cn, N, L, A, D), **z), next_states=torch.empty((Cn, N, L, A, D), **z),
                   actions=torch.empty((Cn, N, L, A), **z), rewards=torch.empty((Cn, N, L, A), **z),
                   dones=torch.empty((Cn, N, L), **z), behv=torch.empty((T, N), **z),
                   alive=torch.empty((T, N), dtype=torch.uint8, device=dev), score=torch.empty(N, **z))
    kind = ENV_CHECKERS if isinstance(env, BatchedCheckers) else ENV_SWITCH
    fa = None
    if forced_actions is not None:
        fa = forced_actions.to(device=dev, dtype=torch.int32).reshape(T, N, A).contiguous()
    tab = _params(net, qnet_tensors)
    _lib.check(_lib.lib().mm_toplevel_rollout(
        tab.array, kind, C.byref(env.cfg), _lib.ptr(out["states"]), _lib.ptr(out["next_states"]), _lib.ptr(out["actions"]),
        _lib.ptr(out["rewards"]), _lib.ptr(out["dones"]), _lib.ptr(out["behv"]), _lib.ptr(out["alive"]), _lib.ptr(out["score"]),
        _lib.ptr(env.state), _lib.ptr(fa), N, T, L, D, NA, float(eps), int(seed), int(step_offset), _lib.stream()))
    return out


def qnet_seq(net, obs, B, L, C_=1, row_index=None, dones=None, reset_on_done=False, save=False):
    """``Q_Net`` over recorded observations.  ``obs`` [C,rows,L,A,D] (or [rows,L,A,D]); ``row_index`` [B] int64 picks the rows
    (PER slots).  Returns q [C,B,L,A,NA] (and the activation record when ``save``)."""
    D, NA = _dims(net)
    dev = obs.device
    q = torch.empty((C_, B, L, 2, NA), dtype=torch.float32, device=dev)
    sv = torch.empty((2, C_ * L, B, _lib.lib().mm_toplevel_qnet_save_stride()), dtype=torch.float32, device=dev) if save else None
    chunk_stride = obs.stride(0) if (obs.dim() == 5 and C_ > 1) else 0
    done_stride = dones.stride(0) if (dones is not None and C_ > 1) else 0
    tab = _params(net, qnet_tensors)
    _lib.check(_lib.lib().mm_toplevel_qnet_seq(
        tab.array, _lib.ptr(obs), _lib.ptr(row_index), chunk_stride, _lib.ptr(dones), done_stride, int(bool(reset_on_done)),
        _lib.ptr(q), _lib.ptr(sv), B, L, C_, D, NA, _lib.stream()))
    return (q, sv) if save else q


def qnet_backward(net, obs, row_index, dones, save, dq, B, L, accumulate=False, reset_on_done=True):
    """BPTT of ``Q_Net`` through sampled chunks; the weight gradients land in ``p.grad`` of the 20 parameter tensors."""
    D, NA = _dims(net)
    lib = _lib.lib()
    tiles = (B + 31) // 32
    partials = torch.empty(2 * tiles * lib.mm_toplevel_qnet_param_count(D, NA), dtype=torch.float32, device=obs.device)
    _lib.check(lib.mm_toplevel_qnet_bwd(
        _params(net, qnet_tensors).array, _grads(net, qnet_tensors).array, _lib.ptr(obs), _lib.ptr(row_index), _lib.ptr(dones),
        int(bool(reset_on_done)), _lib.ptr(save), _lib.ptr(dq), _lib.ptr(partials), int(bool(accumulate)), B, L, D, NA,
        _lib.stream()))


def td_chunks(q_tgt, rewards, dones, behv, alive, gamma):
    """``cal_td_error`` per step, summed per chunk over the alive steps: td_sum [C,N], chunk_valid [C,N] (uint8)."""
    Cn, N, L, _, NA = q_tgt.shape
    td = torch.empty((Cn, N), dtype=torch.float32, device=q_tgt.device)
    valid = torch.empty((Cn, N), dtype=torch.uint8, device=q_tgt.device)
    _lib.check(_lib.lib().mm_toplevel_td(_lib.ptr(q_tgt), _lib.ptr(rewards), _lib.ptr(dones), _lib.ptr(behv), _lib.ptr(alive),
                                         _lib.ptr(td), _lib.ptr(valid), N, L, Cn, NA, float(gamma), _lib.stream()))
    return td, valid


def vdn_chunk_loss(q_beh, q_tgt, actions, rewards, dones, row_index, w, gamma):
    """se [B,L], dq [B,L,A,NA], td_last [B] for ``Target_Dqn.train``."""
    _, B, L, _, NA = q_beh.shape
    dev = q_beh.device
    se = torch.empty((B, L), dtype=torch.float32, device=dev)
    dq = torch.empty((B, L, 2, NA), dtype=torch.float32, device=dev)
    td = torch.empty(B, dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().mm_vdn_chunk_loss(_lib.ptr(q_beh), _lib.ptr(q_tgt), _lib.ptr(actions), _lib.ptr(rewards), _lib.ptr(dones),
                                            _lib.ptr(row_index), _lib.ptr(w), _lib.ptr(se), _lib.ptr(dq), _lib.ptr(td), B, L, NA,
                                            float(gamma), _lib.stream()))
    return se, dq, td


def mix_seq(mix, obs, row_index, dones, q_all, actions, B, L, save=False):
    """``Mix_Net`` over sampled chunks: Qtot [B,L] from the taken-action values (``actions`` given) or max_o (``None``)."""
    S = mix.gru.weight_ih.shape[1]
    NA = q_all.shape[-1]
    dev = obs.device
    qtot = torch.empty((B, L), dtype=torch.float32, device=dev)
    sv = torch.empty((L, B, _lib.lib().mm_toplevel_mix_save_stride()), dtype=torch.float32, device=dev) if save else None
    _lib.check(_lib.lib().mm_toplevel_mix_seq(_params(mix, mix_tensors).array, _lib.ptr(obs), _lib.ptr(row_index), _lib.ptr(dones),
                                              _lib.ptr(q_all), _lib.ptr(actions), _lib.ptr(qtot), _lib.ptr(sv), B, L, S, NA,
                                              _lib.stream()))
    return (qtot, sv) if save else qtot


def mix_backward(mix, obs, row_index, dones, q_all, actions, save, dqtot, B, L, accumulate=False):
    """BPTT of ``Mix_Net``: weight gradients into ``p.grad``; returns dq [B,L,A,NA] for ``qnet_backward``."""
    S = mix.gru.weight_ih.shape[1]
    NA = q_all.shape[-1]
    lib = _lib.lib()
    dq = torch.empty((B, L, 2, NA), dtype=torch.float32, device=obs.device)
    partials = torch.empty(((B + 31) // 32) * lib.mm_toplevel_mix_param_count(S), dtype=torch.float32, device=obs.device)
    _lib.check(lib.mm_toplevel_mix_bwd(
        _params(mix, mix_tensors).array, _grads(mix, mix_tensors).array, _lib.ptr(obs), _lib.ptr(row_index), _lib.ptr(dones),
        _lib.ptr(q_all), _lib.ptr(actions), _lib.ptr(save), _lib.ptr(dqtot), _lib.ptr(dq), _lib.ptr(partials),
        int(bool(accumulate)), B, L, S, NA, _lib.stream()))
    return dq


def qmix_chunk_loss(qtot, qtot_tgt, rewards, dones, row_index, w, gamma):
    B, L = qtot.shape
    dev = qtot.device
    se = torch.empty((B, L), dtype=torch.float32, device=dev)
    dqtot = torch.empty((B, L), dtype=torch.float32, device=dev)
    td = torch.empty(B, dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().mm_qmix_chunk_loss(_lib.ptr(qtot), _lib.ptr(qtot_tgt), _lib.ptr(rewards), _lib.ptr(dones),
                                             _lib.ptr(row_index), _lib.ptr(w), _lib.ptr(se), _lib.ptr(dqtot), _lib.ptr(td), B, L,
                                             float(gamma), _lib.stream()))
    return se, dqtot, td
