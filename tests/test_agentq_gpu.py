"""GPU parity: fused agent network forward vs outputs recorded from the reference's AgentQFunction
(tests/golden/offpolicy_qmix.npz).  fp32 FFMA path: tolerance 2e-5 abs on q (|q| ~ 0.05), 2e-5 on h."""
import argparse
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
TOL = 2e-5


def _args(hidden=64, precision="fp32"):
    return argparse.Namespace(hidden_size=hidden, use_rnn_layer=True, use_orthogonal=True, gain=0.01, layer_N=1,
                              recurrent_N=1, use_feature_normalization=True, use_ReLU=True, use_conv1d=False,
                              fused_precision=precision)


def _load(net, g, prefix):
    sd = {k[len(prefix):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix)}
    missing = net.load_state_dict(sd, strict=True)      # same keys as the reference, incl. the unused fc_h
    return missing


def test_state_dict_keys_and_fused_forward_match_reference(cuda):
    from minimarl_b200.nn import AgentQFunction
    g = np.load(os.path.join(GOLDEN, "offpolicy_qmix.npz"))
    net = AgentQFunction(_args(), 47, 5, cuda)
    _load(net, g, "q.")
    assert sum(p.numel() for p in net.parameters()) == 37283
    obs = torch.from_numpy(g["obs"]).to(cuda)                        # [A, T+1, B, D]
    stacked = torch.cat(list(obs), dim=-2)                           # [T+1, A*B, D]
    with torch.no_grad():
        q, h = net(stacked, torch.zeros(1, stacked.shape[1], 64, device=cuda))
    assert q.shape == g["q_all"].shape and h.shape == g["h_final"].shape
    assert np.abs(q.cpu().numpy() - g["q_all"]).max() < TOL
    assert np.abs(h.cpu().numpy() - g["h_final"]).max() < TOL
    with torch.no_grad():                                            # no_sequence form used during rollouts
        q1, h1 = net(stacked[0], torch.zeros(stacked.shape[1], 64, device=cuda))
    assert q1.shape == g["q_step0"].shape
    assert np.abs(q1.cpu().numpy() - g["q_step0"]).max() < TOL
    assert np.abs(h1.cpu().numpy() - g["h_step0"]).max() < TOL


@pytest.mark.parametrize("hidden,rows,T1,D", [(64, 1000, 9, 47), (32, 777, 5, 47), (64, 130, 3, 94), (32, 64, 1, 3)])
def test_fused_equals_eager_layer_chain(cuda, hidden, rows, T1, D):
    """Ragged row counts, both hidden sizes, non-zero h0: kernel vs the eager torch modules (same weights)."""
    from minimarl_b200.nn import AgentQFunction
    torch.manual_seed(hidden + rows)
    net = AgentQFunction(_args(hidden), D, 5, cuda)
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
        net.q.action_out.weight.mul_(30)
    obs = torch.rand((T1, rows, D), device=cuda)
    h0 = torch.randn((rows, hidden), device=cuda) * 0.5
    import copy
    with torch.no_grad():
        q, h = net(obs, h0)
        # reference = the same modules in fp32 on the CPU (cuDNN's GRU would use TF32 on the GPU)
        rnn_cpu, head_cpu = copy.deepcopy(net.rnn).cpu(), copy.deepcopy(net.q).cpu()
        rnn_outs, h_ref = rnn_cpu(obs.cpu(), h0.cpu()[None].contiguous())
        q_ref = head_cpu(rnn_outs)
    assert (q.cpu() - q_ref).abs().max().item() < 5e-5
    assert (h.cpu() - h_ref).abs().max().item() < 5e-5


def test_autograd_pass_uses_the_same_parameters(cuda):
    from minimarl_b200.nn import AgentQFunction
    torch.backends.cudnn.allow_tf32 = False                          # keep the eager GRU in fp32 for the comparison
    net = AgentQFunction(_args(), 47, 5, cuda)
    obs = torch.rand((4, 32, 47), device=cuda)
    q, _ = net(obs, torch.zeros(1, 32, 64, device=cuda))             # grad enabled: eager path, differentiable
    q.sum().backward()
    assert net.rnn.mlp.fc1[0].weight.grad is not None
    assert net.rnn.mlp.fc_h[0].weight.grad is None                  # unused duplicate gets no grad (Q6)
    with torch.no_grad():
        q2, _ = net(obs, torch.zeros(1, 32, 64, device=cuda))
    assert (q - q2).abs().max().item() < 5e-5
    with torch.no_grad():                                            # an optimiser step invalidates the packed cache
        for p in net.parameters():
            p.add_(0.01)
        q3, _ = net(obs, torch.zeros(1, 32, 64, device=cuda))
    assert (q3 - q2).abs().max().item() > 1e-4


# ---- tensor-core path (tcgen05, TF32 operands, fp32 accumulate) ---------------------------------------------
# Tolerance: TF32 keeps 10 mantissa bits (2^-11 relative per operand); through 3 GEMM layers + LayerNorms and a
# T-step recurrence the Q-values stay within 3e-3 absolute (|q| <= ~1) / 2e-3 relative of the fp32 reference.
def test_tensor_core_forward_matches_reference_fixture(cuda):
    from minimarl_b200.nn import AgentQFunction
    g = np.load(os.path.join(GOLDEN, "offpolicy_qmix.npz"))
    net = AgentQFunction(_args(precision="tf32"), 47, 5, cuda)
    assert net.tc_ok
    _load(net, g, "q.")
    stacked = torch.cat(list(torch.from_numpy(g["obs"]).to(cuda)), dim=-2)
    with torch.no_grad():
        q, h = net(stacked, torch.zeros(1, stacked.shape[1], 64, device=cuda))
    assert np.abs(q.cpu().numpy() - g["q_all"]).max() < 1e-3          # |q| ~ 0.1 here (head gain 0.01)
    assert np.abs(h.cpu().numpy() - g["h_final"]).max() < 3e-3


@pytest.mark.parametrize("hidden,rows,T1,D", [(64, 1000, 9, 47), (32, 777, 5, 47), (64, 128, 40, 47), (32, 64, 1, 3)])
def test_tensor_core_equals_fp32_kernel(cuda, hidden, rows, T1, D):
    from minimarl_b200.nn import AgentQFunction
    from minimarl_b200.nn.agent_q import agentq_forward
    torch.manual_seed(hidden + rows)
    net = AgentQFunction(_args(hidden, "tf32"), D, 5, cuda)
    assert net.tc_ok
    with torch.no_grad():
        for p in net.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
        net.q.action_out.weight.mul_(30)
    obs = torch.rand((T1, rows, D), device=cuda)
    h0 = torch.randn((rows, hidden), device=cuda) * 0.5
    q32, h32, r32 = agentq_forward(net.packed(), obs, h0, D, hidden, 5, want_rnn_out=True)
    qtc, htc, rtc = agentq_forward(net.packed_tc(), obs, h0, D, hidden, 5, want_rnn_out=True, tensor_cores=True)
    scale = max(1.0, q32.abs().max().item())
    assert (qtc - q32).abs().max().item() < 4e-3 * scale
    assert (htc - h32).abs().max().item() < 4e-3
    assert (rtc - r32).abs().max().item() < 2e-2                      # LayerNorm output, O(1..3) values
    assert ((qtc - q32).norm() / q32.norm()).item() < 2e-3


def test_tensor_core_falls_back_when_tile_does_not_fit(cuda):
    from minimarl_b200.nn import AgentQFunction
    net = AgentQFunction(_args(64, "tf32"), 94, 5, cuda)             # full_observable Checkers: D = 94
    assert not net.tc_ok
    with torch.no_grad():
        q, _ = net(torch.rand((2, 8, 94), device=cuda), torch.zeros(8, 64, device=cuda))
    assert q.shape == (2, 8, 5)
