"""QMIX / VDN mixers with the reference's module tree, fused on the GPU when no gradient is needed.

``QMixer`` mirrors offpolicy/algorithms/qmix/algorithm/q_mixer.py:8-94 (state_dict keys ``hyper_w1.0.weight``,
``hyper_w1.2.weight``, ``hyper_w2.*``, ``hyper_b1.*``, ``hyper_b2.*``); ``VDNMixer`` mirrors
offpolicy/algorithms/vdn/algorithm/vdn_mixer.py:6-39.  The target mixer of a QMIX update always runs
under ``torch.no_grad()`` and therefore always takes the one-kernel path (``mm_qmixer_forward``).
"""
import ctypes as C

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib
from ._direct import grad_targets
from .linear import Linear

_NAMES = ["wcat", "bcat", "w1b", "b1b", "w2b", "b2b", "wb2", "bb2"]


def pack_qmixer(sd, SD, A, device, prefix=""):
    lib = _lib.lib()
    total = lib.mm_qmixer_param_count(SD, A)
    if total < 0:
        raise RuntimeError(f"fused mixer supports 2 or 4 agents; got A={A}")
    offs = (C.c_int32 * 8)()
    _lib.check(lib.mm_qmixer_param_offsets(SD, A, offs))
    o = dict(zip(_NAMES, list(offs)))
    g = lambda k: sd[prefix + k].detach().to(device=device, dtype=torch.float32)
    buf = torch.zeros(total, dtype=torch.float32, device=device)
    K1 = (SD + 3) // 4 * 4
    wcat = torch.zeros((K1, 256), dtype=torch.float32, device=device)
    bcat = torch.zeros(256, dtype=torch.float32, device=device)
    for col, key, width in [(0, "hyper_w1.0", 64), (64, "hyper_w2.0", 64), (128, "hyper_b2.0", 64), (192, "hyper_b1", 32)]:
        wcat[:SD, col:col + width] = g(key + ".weight").t()
        bcat[col:col + width] = g(key + ".bias")

    def put(name, t):
        t = t.reshape(-1)
        buf[o[name]:o[name] + t.numel()] = t

    put("wcat", wcat); put("bcat", bcat)
    put("w1b", g("hyper_w1.2.weight").t().contiguous()); put("b1b", g("hyper_w1.2.bias"))
    w2b = torch.zeros((64, 64), dtype=torch.float32, device=device)
    w2b[:, :32] = g("hyper_w2.2.weight").t()
    b2b = torch.zeros(64, dtype=torch.float32, device=device)
    b2b[:32] = g("hyper_w2.2.bias")
    put("w2b", w2b); put("b2b", b2b)
    put("wb2", g("hyper_b2.2.weight")); put("bb2", g("hyper_b2.2.bias"))
    return buf


_TC_NAMES = ["wcat", "w1b", "w2b", "bcat", "b1b", "b2b", "wb2", "bb2"]


def pack_qmixer_tc(sd, SD, A, device, prefix=""):
    """Parameter buffer of the tensor-core mixer (mm_qmixer_tc_forward): the PyTorch [out][in] matrices as they are, the four
    first layers stacked row-wise with their input columns zero-padded to 96.  None when the shape is not supported."""
    lib = _lib.lib()
    total = lib.mm_qmixer_tc_param_count(SD, A)
    if total < 0:
        return None
    offs = (C.c_int32 * 8)()
    _lib.check(lib.mm_qmixer_tc_param_offsets(SD, A, offs))
    o = dict(zip(_TC_NAMES, list(offs)))
    g = lambda k: sd[prefix + k].detach().to(device=device, dtype=torch.float32)
    buf = torch.zeros(total, dtype=torch.float32, device=device)
    wcat = torch.zeros((224, 96), dtype=torch.float32, device=device)
    bcat = torch.zeros(224, dtype=torch.float32, device=device)
    for r0, key, width in [(0, "hyper_w1.0", 64), (64, "hyper_w2.0", 64), (128, "hyper_b2.0", 64), (192, "hyper_b1", 32)]:
        wcat[r0:r0 + width, :SD] = g(key + ".weight")
        bcat[r0:r0 + width] = g(key + ".bias")

    def put(name, t):
        t = t.reshape(-1)
        buf[o[name]:o[name] + t.numel()] = t

    put("wcat", wcat); put("bcat", bcat)
    put("w1b", g("hyper_w1.2.weight")); put("b1b", g("hyper_w1.2.bias"))
    put("w2b", g("hyper_w2.2.weight")); put("b2b", g("hyper_w2.2.bias"))
    put("wb2", g("hyper_b2.2.weight")); put("bb2", g("hyper_b2.2.bias"))
    return buf


class _MixingFn(torch.autograd.Function):
    """hidden = elu(q @ w1 + b1); out = hidden . w2 + b2 per row, one warp per row (mm_qmix_mixing_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, q, w1, b1, w2, b2):
        M, A = q.shape
        q, w1, b1, w2, b2 = (t.contiguous() for t in (q, w1, b1, w2, b2))
        out = torch.empty(M, dtype=torch.float32, device=q.device)
        _lib.check(_lib.lib().mm_qmix_mixing_fwd(_lib.ptr(q), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2), _lib.ptr(b2),
                                                 _lib.ptr(out), M, A, _lib.stream()))
        ctx.save_for_backward(q, w1, b1, w2)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, w1, b1, w2 = ctx.saved_tensors
        M, A = q.shape
        dq, dw1, db1, dw2 = torch.empty_like(q), torch.empty_like(w1), torch.empty_like(b1), torch.empty_like(w2)
        db2 = torch.empty(M, dtype=torch.float32, device=q.device)
        _lib.check(_lib.lib().mm_qmix_mixing_bwd(_lib.ptr(dout.contiguous()), _lib.ptr(q), _lib.ptr(w1), _lib.ptr(b1), _lib.ptr(w2),
                                                 _lib.ptr(dq), _lib.ptr(dw1), _lib.ptr(db1), _lib.ptr(dw2), _lib.ptr(db2),
                                                 M, A, _lib.stream()))
        return dq, dw1, db1, dw2, db2


def _ptr_array(params):
    arr = (C.c_void_p * len(params))()
    for i, t in enumerate(params):
        if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
            raise RuntimeError("minimarl_b200 fused mixer takes contiguous float32 CUDA parameters (no CPU fallback)")
        arr[i] = t.data_ptr()
    return arr


_BWD_DIMS = {}


def _bwd_dims(SD):
    if SD not in _BWD_DIMS:
        d, off = (C.c_int32 * 4)(), (C.c_int32 * 15)()
        _lib.check(_lib.lib().mm_qmixer_bwd_dims(SD, d))
        _lib.check(_lib.lib().mm_qmixer_bwd_grad_layout(SD, off))
        _BWD_DIMS[SD] = (list(d), list(off))
    return _BWD_DIMS[SD]


class _QMixerLiveFn(torch.autograd.Function):
    """The differentiated mixer as two kernels: forward = ``mm_qmixer_tc_forward`` (nothing saved but the inputs), backward =
    ``mm_qmixer_bwd`` (recomputes a 128-row tile's activations, dq and every weight gradient on tcgen05) + ``_unfold``.
    params: the 14 tensors of ``QMixer.parameters()`` (hypernet_layers == 2) in module order."""

    @staticmethod
    def forward(ctx, q, states, *params):
        M, A = q.shape
        SD = states.shape[1]
        lib = _lib.lib()
        q, states = q.contiguous(), states.contiguous()
        ctx.param_objs = list(params)                   # the Parameters themselves (their .grad may take the gradients directly)
        params = [t.detach().contiguous() for t in params]
        packed = torch.empty(lib.mm_qmixer_tc_param_count(SD, A), dtype=torch.float32, device=q.device)
        _lib.check(lib.mm_qmixer_tc_pack(_ptr_array(params), _lib.ptr(packed), SD, A, _lib.stream()))
        out = torch.empty(M, dtype=torch.float32, device=q.device)
        _lib.check(lib.mm_qmixer_tc_forward(_lib.ptr(packed), _lib.ptr(q), _lib.ptr(states), _lib.ptr(out), M, SD, A, _lib.stream()))
        ctx.save_for_backward(q, states, *params)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, states = ctx.saved_tensors[:2]
        params = list(ctx.saved_tensors[2:])
        M, A = q.shape
        SD = states.shape[1]
        lib = _lib.lib()
        (acc_cols, vec_n, tile_rows, n_grads), off = _bwd_dims(SD)
        dout = dout.contiguous()
        z = dict(dtype=torch.float32, device=q.device)
        scale = torch.empty(2, **z)
        _lib.check(lib.mm_ppo_trunk2_scale(_lib.ptr(dout), M, _lib.ptr(scale), _lib.stream()))
        n_ctas = lib.mm_qmixer_bwd_ctas(M)
        acc, vec = torch.empty((n_ctas, tile_rows, acc_cols), **z), torch.empty((n_ctas, vec_n), **z)
        dq = torch.empty_like(q)
        _lib.check(lib.mm_qmixer_bwd(_ptr_array(params), _lib.ptr(q), _lib.ptr(states), _lib.ptr(dout), _lib.ptr(scale), _lib.ptr(dq),
                                     _lib.ptr(acc), _lib.ptr(vec), n_ctas, M, SD, A, _lib.stream()))
        targets = grad_targets(ctx.param_objs, ctx.needs_input_grad[2:])
        if targets is not None:
            gptrs = (C.c_void_p * len(targets))(*[g.data_ptr() for g in targets])
            _lib.check(lib.mm_qmixer_bwd_unfold_into(_lib.ptr(acc), _lib.ptr(vec), n_ctas, _lib.ptr(scale), gptrs, SD, _lib.stream()))
            return (dq if ctx.needs_input_grad[0] else None, None) + (None,) * len(params)
        grads = torch.empty(n_grads, **z)
        _lib.check(lib.mm_qmixer_bwd_unfold(_lib.ptr(acc), _lib.ptr(vec), n_ctas, _lib.ptr(scale), _lib.ptr(grads), SD, _lib.stream()))
        return (dq if ctx.needs_input_grad[0] else None, None) + tuple(
            grads[off[i]:off[i + 1]].view_as(params[i]) if ctx.needs_input_grad[2 + i] else None for i in range(len(params)))


def _init(m, use_orthogonal):
    (nn.init.orthogonal_ if use_orthogonal else nn.init.xavier_uniform_)(m.weight.data, gain=1)
    nn.init.constant_(m.bias.data, 0)
    return m


class QMixer(nn.Module):
    def __init__(self, args, num_agents, cent_obs_dim, device, multidiscrete_list=None):
        super().__init__()
        if multidiscrete_list:
            raise NotImplementedError("multi-discrete mixing is outside the hot path")
        self.device = torch.device(device)
        self.tpdv = dict(dtype=torch.float32, device=self.device)
        self.num_agents = num_agents
        self.cent_obs_dim = cent_obs_dim
        self.hidden_layer_dim = args.mixer_hidden_dim
        self.hypernet_hidden_dim = args.hypernet_hidden_dim
        self.num_mixer_q_inps = num_agents
        uo = args.use_orthogonal
        lin = lambda i, o: _init(Linear(i, o), uo)
        S, HH, E, A = cent_obs_dim, self.hypernet_hidden_dim, self.hidden_layer_dim, num_agents
        if args.hypernet_layers == 1:
            self.hyper_w1, self.hyper_w2 = lin(S, A * E), lin(S, E)
        elif args.hypernet_layers == 2:
            self.hyper_w1 = nn.Sequential(lin(S, HH), nn.ReLU(), lin(HH, A * E))
            self.hyper_w2 = nn.Sequential(lin(S, HH), nn.ReLU(), lin(HH, E))
        self.hyper_b1 = lin(S, E)
        self.hyper_b2 = nn.Sequential(lin(S, HH), nn.ReLU(), lin(HH, 1))
        self.to(self.device)
        self.fusable = args.hypernet_layers == 2 and HH == 64 and E == 32 and A in (2, 4)
        self._packed, self._packed_key = None, None
        # "tf32": the no-grad pass runs on the tensor cores (mm_qmixer_tc_forward, ~1e-3 relative on Q_tot); "fp32": FFMA kernel
        self.precision = getattr(args, "fused_precision", "tf32")
        self.tc_ok = self.fusable and _lib.lib().mm_qmixer_tc_param_count(S, A) > 0
        self._packed_tc, self._packed_tc_key = None, None

    def packed(self):
        key = tuple(p._version for p in self.parameters()) + tuple(p.data_ptr() for p in self.parameters())
        if self._packed is None or key != self._packed_key:
            self._packed = pack_qmixer(self.state_dict(), self.cent_obs_dim, self.num_agents, self.device)
            self._packed_key = key
        return self._packed

    def packed_tc(self):
        key = tuple(p._version for p in self.parameters()) + tuple(p.data_ptr() for p in self.parameters())
        if self._packed_tc is None or key != self._packed_tc_key:
            self._packed_tc = pack_qmixer_tc(self.state_dict(), self.cent_obs_dim, self.num_agents, self.device)
            self._packed_tc_key = key
        return self._packed_tc

    def forward(self, agent_q_inps, states):
        agent_q_inps = torch.as_tensor(agent_q_inps).to(**self.tpdv)
        states = torch.as_tensor(states).to(**self.tpdv)
        batch_size = agent_q_inps.size(1)
        needs_graph = torch.is_grad_enabled() and (agent_q_inps.requires_grad or
                                                   any(p.requires_grad for p in self.parameters()))
        if self.fusable and not needs_graph:
            s = states.reshape(-1, self.cent_obs_dim).contiguous()
            q = agent_q_inps.reshape(-1, self.num_agents).contiguous()
            out = torch.empty(s.shape[0], **self.tpdv)
            if self.precision == "tf32" and self.tc_ok:
                _lib.check(_lib.lib().mm_qmixer_tc_forward(_lib.ptr(self.packed_tc()), _lib.ptr(q), _lib.ptr(s), _lib.ptr(out),
                                                           s.shape[0], self.cent_obs_dim, self.num_agents, _lib.stream()))
            else:
                _lib.check(_lib.lib().mm_qmixer_forward(_lib.ptr(self.packed()), _lib.ptr(q), _lib.ptr(s), _lib.ptr(out),
                                                        s.shape[0], self.cent_obs_dim, self.num_agents, _lib.stream()))
            return out.view(-1, batch_size, 1, 1)
        states = states.view(-1, batch_size, self.cent_obs_dim).float()
        if (states.is_cuda and self.fusable and self.tc_ok and self.precision == "tf32" and self.cent_obs_dim <= 94
                and torch.backends.cuda.matmul.allow_tf32):
            # differentiated pass on the tensor cores: forward kernel + ONE backward kernel with the weight gradients inside
            out = _QMixerLiveFn.apply(agent_q_inps.reshape(-1, self.num_agents), states.reshape(-1, self.cent_obs_dim),
                                      *self.parameters())
            return out.view(-1, batch_size, 1, 1)
        if states.is_cuda and self.hidden_layer_dim == 32 and self.num_mixer_q_inps in (2, 4):
            # differentiated pass: hypernetworks through cuBLAS, the per-row mixing through one custom kernel
            s2 = states.reshape(-1, self.cent_obs_dim)
            if self.cent_obs_dim % 4:                                    # pad once for the four first layers (nn/linear.py)
                s2 = F.pad(s2, (0, (-self.cent_obs_dim) % 4))
            w1 = torch.abs(self.hyper_w1(s2)).view(-1, self.num_mixer_q_inps, self.hidden_layer_dim)
            w2 = torch.abs(self.hyper_w2(s2))
            out = _MixingFn.apply(agent_q_inps.reshape(-1, self.num_mixer_q_inps), w1, self.hyper_b1(s2), w2,
                                  self.hyper_b2(s2).reshape(-1))
            return out.view(-1, batch_size, 1, 1)
        agent_q_inps = agent_q_inps.view(-1, batch_size, 1, self.num_mixer_q_inps)
        w1 = torch.abs(self.hyper_w1(states)).view(-1, batch_size, self.num_mixer_q_inps, self.hidden_layer_dim)
        b1 = self.hyper_b1(states).view(-1, batch_size, 1, self.hidden_layer_dim)
        hidden_layer = F.elu(torch.matmul(agent_q_inps, w1) + b1)
        w2 = torch.abs(self.hyper_w2(states)).view(-1, batch_size, self.hidden_layer_dim, 1)
        b2 = self.hyper_b2(states).view(-1, batch_size, 1, 1)
        return (torch.matmul(hidden_layer, w2) + b2).view(-1, batch_size, 1, 1)


class VDNMixer(nn.Module):
    def __init__(self, args, num_agents, cent_obs_dim, device, multidiscrete_list=None):
        super().__init__()
        self.device = torch.device(device)
        self.num_agents = num_agents
        self.num_mixer_q_inps = num_agents

    def forward(self, agent_q_inps, states=None):
        agent_q_inps = torch.as_tensor(agent_q_inps).to(device=self.device, dtype=torch.float32)
        if torch.is_grad_enabled() and agent_q_inps.requires_grad:
            return agent_q_inps.sum(dim=-1).view(-1, 1, 1)
        q = agent_q_inps.reshape(-1, self.num_agents).contiguous()
        out = torch.empty(q.shape[0], dtype=torch.float32, device=self.device)
        _lib.check(_lib.lib().mm_vdn_mix(_lib.ptr(q), _lib.ptr(out), q.shape[0], self.num_agents, _lib.stream()))
        return out.view(-1, 1, 1)
