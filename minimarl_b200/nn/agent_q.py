"""Recurrent agent network with the reference's module tree and state_dict keys, fused on the GPU.

Mirrors ``AgentQFunction`` (offpolicy/algorithms/qmix/algorithm/agent_q_function.py:8-66) and its parts
``MLPLayer`` / ``MLPBase`` (offpolicy/utils/algorithm_utils/mlp.py:7-89), ``RNNLayer`` / ``RNNBase``
(rnn.py:4-47), ``ACTLayer`` (act.py:5-37): same attribute names, so ``state_dict`` keys
(``rnn.feature_norm.weight``, ``rnn.mlp.fc1.0.weight``, ``rnn.mlp.fc2.0.0.weight``,
``rnn.rnn.rnn.weight_ih_l0``, ``rnn.rnn.norm.weight``, ``q.action_out.weight`` ... including the unused
``rnn.mlp.fc_h``) load both ways, optimisers and target copies work unchanged.

``forward`` runs the whole chain over the whole sequence in ONE kernel (``mm_agentq_forward``) whenever
no gradient is needed -- rollout action selection, the target network and the double-Q greedy pass,
i.e. two of the three network passes of a QMIX update -- and falls back to the identical eager PyTorch
layer chain only when autograd must record the graph (the live-network pass that is differentiated).
"""
import copy
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from .linear import Linear

from .. import _lib
from .gru import GRU
from .layernorm import LayerNorm

_OFFSET_NAMES = ["ln0_w", "ln0_b", "w1", "b1", "ln1_w", "ln1_b", "w2", "b2", "ln2_w", "ln2_b", "wih", "whh",
                 "bih", "bhh", "ln3_w", "ln3_b", "wq", "bq"]


def agentq_layout(D, H, NA):
    lib = _lib.lib()
    total = lib.mm_agentq_param_count(D, H, NA)
    if total < 0:
        raise RuntimeError(f"fused agent network supports D <= 96, H in (32, 64); got D={D}, H={H}")
    offs = (C.c_int32 * 18)()
    _lib.check(lib.mm_agentq_param_offsets(D, H, NA, offs))
    return total, dict(zip(_OFFSET_NAMES, list(offs)))


def pack_agentq(sd, D, H, NA, device, prefix=""):
    """Flatten a reference-keyed state_dict into the kernel's packed fp32 buffer (matrices transposed)."""
    total, o = agentq_layout(D, H, NA)
    buf = torch.zeros(total, dtype=torch.float32, device=device)
    g = lambda k: sd[prefix + k].detach().to(device=device, dtype=torch.float32)

    def put(name, t):
        t = t.reshape(-1)
        buf[o[name]:o[name] + t.numel()] = t

    put("ln0_w", g("rnn.feature_norm.weight")); put("ln0_b", g("rnn.feature_norm.bias"))
    K1 = (D + 3) // 4 * 4
    w1 = torch.zeros((K1, H), dtype=torch.float32, device=device)
    w1[:D] = g("rnn.mlp.fc1.0.weight").t()
    put("w1", w1); put("b1", g("rnn.mlp.fc1.0.bias"))
    put("ln1_w", g("rnn.mlp.fc1.2.weight")); put("ln1_b", g("rnn.mlp.fc1.2.bias"))
    put("w2", g("rnn.mlp.fc2.0.0.weight").t().contiguous()); put("b2", g("rnn.mlp.fc2.0.0.bias"))
    put("ln2_w", g("rnn.mlp.fc2.0.2.weight")); put("ln2_b", g("rnn.mlp.fc2.0.2.bias"))
    put("wih", g("rnn.rnn.rnn.weight_ih_l0").t().contiguous()); put("whh", g("rnn.rnn.rnn.weight_hh_l0").t().contiguous())
    put("bih", g("rnn.rnn.rnn.bias_ih_l0")); put("bhh", g("rnn.rnn.rnn.bias_hh_l0"))
    put("ln3_w", g("rnn.rnn.norm.weight")); put("ln3_b", g("rnn.rnn.norm.bias"))
    NAp = (NA + 3) // 4 * 4
    wq = torch.zeros((H, NAp), dtype=torch.float32, device=device)
    wq[:, :NA] = g("q.action_out.weight").t()
    put("wq", wq)
    bq = torch.zeros(NAp, dtype=torch.float32, device=device)
    bq[:NA] = g("q.action_out.bias")
    put("bq", bq)
    return buf


def agentq_tc_layout(D, H, NA):
    lib = _lib.lib()
    total = lib.mm_agentq_tc_param_count(D, H, NA)
    if total < 0:
        return None, None
    offs = (C.c_int32 * 18)()
    _lib.check(lib.mm_agentq_tc_param_offsets(D, H, NA, offs))
    return total, dict(zip(_OFFSET_NAMES, list(offs)))


def pack_agentq_tc(sd, D, H, NA, device, prefix=""):
    """Parameter buffer of the tensor-core kernel: the PyTorch tensors as they are ([out][in] matrices), only
    the Linear(D,H) weight zero-padded to a multiple of 8 input columns."""
    total, o = agentq_tc_layout(D, H, NA)
    buf = torch.zeros(total, dtype=torch.float32, device=device)
    g = lambda k: sd[prefix + k].detach().to(device=device, dtype=torch.float32)

    def put(name, t):
        t = t.reshape(-1)
        buf[o[name]:o[name] + t.numel()] = t

    K1 = (D + 7) // 8 * 8
    w1 = torch.zeros((H, K1), dtype=torch.float32, device=device)
    w1[:, :D] = g("rnn.mlp.fc1.0.weight")
    put("w1", w1); put("w2", g("rnn.mlp.fc2.0.0.weight"))
    put("wih", g("rnn.rnn.rnn.weight_ih_l0")); put("whh", g("rnn.rnn.rnn.weight_hh_l0"))
    put("ln0_w", g("rnn.feature_norm.weight")); put("ln0_b", g("rnn.feature_norm.bias"))
    put("b1", g("rnn.mlp.fc1.0.bias")); put("ln1_w", g("rnn.mlp.fc1.2.weight")); put("ln1_b", g("rnn.mlp.fc1.2.bias"))
    put("b2", g("rnn.mlp.fc2.0.0.bias")); put("ln2_w", g("rnn.mlp.fc2.0.2.weight")); put("ln2_b", g("rnn.mlp.fc2.0.2.bias"))
    put("bih", g("rnn.rnn.rnn.bias_ih_l0")); put("bhh", g("rnn.rnn.rnn.bias_hh_l0"))
    put("ln3_w", g("rnn.rnn.norm.weight")); put("ln3_b", g("rnn.rnn.norm.bias"))
    put("wq", g("q.action_out.weight")); put("bq", g("q.action_out.bias"))
    return buf


def agentq_forward(packed, obs, h0, D, H, NA, want_q=True, want_rnn_out=False, tensor_cores=False):
    """obs [T1, R, D] (or [R, D]); h0 [R, H] or None.  Returns (q [T1,R,NA], h_final [R,H], rnn_out|None).
    ``tensor_cores`` selects mm_agentq_tc_forward (tcgen05, TF32) and needs a buffer from pack_agentq_tc."""
    seq = obs.dim() == 3
    if not seq:
        obs = obs.unsqueeze(0)
    obs = obs.contiguous()
    T1, R = obs.shape[0], obs.shape[1]
    z = dict(dtype=torch.float32, device=obs.device)
    q = torch.empty((T1, R, NA), **z) if want_q else None
    h_final = torch.empty((R, H), **z)
    rnn_out = torch.empty((T1, R, H), **z) if want_rnn_out else None
    if h0 is not None:
        h0 = h0.reshape(R, H).contiguous()
    fn = _lib.lib().mm_agentq_tc_forward if tensor_cores else _lib.lib().mm_agentq_forward
    _lib.check(fn(_lib.ptr(packed), _lib.ptr(obs), _lib.ptr(h0), _lib.ptr(q), _lib.ptr(h_final), _lib.ptr(rnn_out),
                  T1, R, D, H, NA, _lib.stream()))
    if not seq and q is not None:
        q = q[0]
    return q, h_final, rnn_out


def _tc_fits(D, H, NA):
    """Shared-memory budget of agentq_tc.cu (weights + activation tile + hidden tile + obs stage + vectors)."""
    total, o = agentq_tc_layout(D, H, NA)
    if total is None:
        return False
    K1 = (D + 7) // 8 * 8
    small = total - o["ln0_w"]
    nbytes = 4 * (H * K1 + 7 * H * H + 128 * max(K1, H) + 128 * H + (128 * D + 3) // 4 * 4 + 128 * 8 + small) + 128
    return nbytes <= 227 * 1024


# ---- module tree (reference names) ---------------------------------------------------------------
def _init_linear(m, use_orthogonal, gain):
    (nn.init.orthogonal_ if use_orthogonal else nn.init.xavier_uniform_)(m.weight.data, gain=gain)
    nn.init.constant_(m.bias.data, 0)
    return m


class MLPLayer(nn.Module):
    def __init__(self, input_dim, hidden_size, layer_N, use_orthogonal, use_ReLU):
        super().__init__()
        self._layer_N = layer_N
        act = nn.ReLU if use_ReLU else nn.Tanh
        gain = nn.init.calculate_gain("relu" if use_ReLU else "tanh")
        mk = lambda i, o: nn.Sequential(_init_linear(Linear(i, o), use_orthogonal, gain), act(), LayerNorm(o))
        self.fc1 = mk(input_dim, hidden_size)
        self.fc_h = mk(hidden_size, hidden_size)        # registered but unused, exactly like the reference (Q6)
        self.fc2 = nn.ModuleList([copy.deepcopy(self.fc_h) for _ in range(layer_N)])

    def forward(self, x):
        x = self.fc1(x)
        for i in range(self._layer_N):
            x = self.fc2[i](x)
        return x


class MLPBase(nn.Module):
    def __init__(self, args, inputs_dim):
        super().__init__()
        self._use_feature_normalization = args.use_feature_normalization
        self._use_orthogonal = args.use_orthogonal
        self._use_ReLU = args.use_ReLU
        self._layer_N = args.layer_N
        self.hidden_size = args.hidden_size
        if getattr(args, "use_conv1d", False):
            raise NotImplementedError("use_conv1d is outside the hot path (SURVEY.md section 2)")
        if self._use_feature_normalization:
            self.feature_norm = LayerNorm(inputs_dim)
        self.mlp = MLPLayer(inputs_dim, self.hidden_size, self._layer_N, self._use_orthogonal, self._use_ReLU)

    def forward(self, x):
        if self._use_feature_normalization:
            x = self.feature_norm(x)
        return self.mlp(x)


class RNNLayer(nn.Module):
    def __init__(self, inputs_dim, outputs_dim, recurrent_N, use_orthogonal):
        super().__init__()
        self.rnn = GRU(inputs_dim, outputs_dim, num_layers=recurrent_N)
        for name, param in self.rnn.named_parameters():
            if "bias" in name:
                nn.init.constant_(param, 0)
            elif "weight" in name:
                (nn.init.orthogonal_ if use_orthogonal else nn.init.xavier_uniform_)(param)
        self.norm = LayerNorm(outputs_dim)

    def forward(self, x, hxs):
        x, hxs = self.rnn(x, hxs)
        return self.norm(x), hxs[0, :, :]


class RNNBase(MLPBase):
    def __init__(self, args, inputs_dim):
        super().__init__(args, inputs_dim)
        self._recurrent_N = args.recurrent_N
        self.rnn = RNNLayer(self.hidden_size, self.hidden_size, self._recurrent_N, self._use_orthogonal)

    def _fused_mlp_ok(self, x):
        """The differentiated pass runs the row-parallel layers as the fused op (csrc/agent_mlp.cu): default network shape on
        CUDA with TF32 matmuls allowed (the switch the fp32 parity path turns off), sequence input."""
        from .agent_mlp import agent_mlp_fits
        return (x.is_cuda and x.dim() == 3 and x.dtype == torch.float32 and torch.is_grad_enabled() and self._use_feature_normalization
                and self._use_ReLU and self._recurrent_N == 1 and agent_mlp_fits(x.shape[-1], self.hidden_size, self._layer_N)
                and torch.backends.cuda.matmul.allow_tf32 and torch.backends.cudnn.allow_tf32
                and getattr(self, "fused_mlp", True))

    def hidden_seq(self, x, hxs):
        """The fused differentiated pass up to the GRU outputs (before ``rnn.norm``): h_seq [T, R, H]."""
        from .agent_mlp import agent_mlp_gi, gru_from_gi
        T, R, D = x.shape
        wb = lambda m: (m.weight, m.bias)
        g = self.rnn.rnn
        gi = agent_mlp_gi(x.reshape(T * R, D), wb(self.feature_norm), wb(self.mlp.fc1[0]), wb(self.mlp.fc1[2]),
                          wb(self.mlp.fc2[0][0]), wb(self.mlp.fc2[0][2]), g.weight_ih_l0, g.bias_ih_l0)
        return gru_from_gi(gi, hxs[0] if hxs is not None else None, g.weight_hh_l0, g.bias_hh_l0, T, R)

    def forward(self, x, hxs):
        if self._fused_mlp_ok(x):
            h_seq = self.hidden_seq(x, hxs)
            return self.rnn.norm(h_seq), h_seq[-1]
        if self._use_feature_normalization:
            x = self.feature_norm(x)
        x = self.mlp(x)
        return self.rnn(x, hxs)


class ACTLayer(nn.Module):
    def __init__(self, act_dim, hidden_size, use_orthogonal, gain):
        super().__init__()
        self.multi_discrete = False
        if isinstance(act_dim, np.ndarray):
            raise NotImplementedError("multi-discrete action heads are outside the hot path")
        self.action_out = _init_linear(Linear(hidden_size, act_dim), use_orthogonal, gain)

    def forward(self, x, no_sequence=False):
        out = self.action_out(x)
        return out[0, :, :] if no_sequence else out


class AgentQFunction(nn.Module):
    """Individual agent q network (RNN): same constructor and ``forward(obs, rnn_states)`` as the reference."""

    def __init__(self, args, input_dim, act_dim, device):
        super().__init__()
        self.hidden_size = args.hidden_size
        self._use_rnn_layer = args.use_rnn_layer
        self.input_dim, self.act_dim = int(input_dim), int(act_dim)
        self.device = torch.device(device)
        self.tpdv = dict(dtype=torch.float32, device=self.device)
        if not self._use_rnn_layer:
            raise NotImplementedError("the non-recurrent (mqmix) agent is outside the hot path")
        self.rnn = RNNBase(args, input_dim)
        self.q = ACTLayer(act_dim, self.hidden_size, args.use_orthogonal, gain=args.gain)
        self.to(self.device)
        self._packed, self._packed_key = None, None
        self._packed_tc, self._packed_tc_key = None, None
        self.fusable = (args.layer_N == 1 and args.recurrent_N == 1 and args.use_feature_normalization and
                        args.use_ReLU and self.hidden_size in (32, 64) and self.input_dim <= 96)
        # "tf32": no-grad passes run on the tensor cores (tcgen05, TF32 operands, ~1e-3 relative on Q);
        # "fp32": FFMA kernel, ~1e-5 of the reference.  The tensor-core tile needs D <= ~56 at H = 64.
        self.precision = getattr(args, "fused_precision", "tf32")
        self.tc_ok = self.fusable and _tc_fits(self.input_dim, self.hidden_size, self.act_dim)

    def packed(self):
        """Packed parameter buffer for the kernel, rebuilt only when a parameter changed."""
        key = tuple(p._version for p in self.parameters()) + tuple(p.data_ptr() for p in self.parameters())
        if self._packed is None or key != self._packed_key:
            self._packed = pack_agentq(self.state_dict(), self.input_dim, self.hidden_size, self.act_dim, self.device)
            self._packed_key = key
        return self._packed

    def packed_tc(self):
        key = tuple(p._version for p in self.parameters()) + tuple(p.data_ptr() for p in self.parameters())
        if self._packed_tc is None or key != self._packed_tc_key:
            self._packed_tc = pack_agentq_tc(self.state_dict(), self.input_dim, self.hidden_size, self.act_dim, self.device)
            self._packed_tc_key = key
        return self._packed_tc

    def forward(self, obs, rnn_states):
        obs = torch.as_tensor(obs).to(**self.tpdv) if not torch.is_tensor(obs) else obs.to(**self.tpdv)
        rnn_states = (torch.as_tensor(rnn_states) if not torch.is_tensor(rnn_states) else rnn_states).to(**self.tpdv)
        no_sequence = obs.dim() == 2
        needs_graph = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if self.fusable and not needs_graph:
            h0 = rnn_states[0] if rnn_states.dim() == 3 else rnn_states
            if self.precision == "tf32" and self.tc_ok:
                q, h_final, _ = agentq_forward(self.packed_tc(), obs, h0, self.input_dim, self.hidden_size,
                                               self.act_dim, tensor_cores=True)
            else:
                q, h_final, _ = agentq_forward(self.packed(), obs, h0, self.input_dim, self.hidden_size, self.act_dim)
            return q, h_final
        if no_sequence:
            obs = obs[None]
        if rnn_states.dim() == 2:
            rnn_states = rnn_states[None]
        from .agent_mlp import ln_head, ln_head_fits
        if not no_sequence and self.rnn._fused_mlp_ok(obs) and ln_head_fits(self.hidden_size, self.act_dim):
            # fused differentiated pass: row-parallel trunk, recurrence, then LayerNorm + head as one op
            h_seq = self.rnn.hidden_seq(obs, rnn_states.contiguous())
            norm, head = self.rnn.rnn.norm, self.q.action_out
            return ln_head(h_seq, norm.weight, norm.bias, head.weight, head.bias), h_seq[-1]
        rnn_outs, h_final = self.rnn(obs, rnn_states.contiguous())
        return self.q(rnn_outs, no_sequence), h_final
