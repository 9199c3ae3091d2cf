"""GPU parity: the fused row-parallel half of the agent network's differentiated pass (mm_agent_mlp_fwd / _bwd: fp16 operands --
tf32's 11-bit significand -- on tcgen05, fp32 accumulation, weight gradients inside the backward kernel) against a float64
statement of the same layers: LayerNorm -> [Linear ReLU LayerNorm] x2 -> GRU input projection (mlp.py:7-89, rnn.py:4-47).
Stated tolerances: gi 1e-2 abs (values of order 1); gradients of the last LayerNorm and the input projection within 3e-3 of
their norm (measured 2e-4 .. 6e-4); gradients below it (behind one or two ReLU + LayerNorm backward passes, where a rounding
of the forward flips ReLUs and the LayerNorm projection amplifies operand rounding) within 5e-2 AND within twice the distance
the same layers have in eager PyTorch with TF32 matmuls on the same inputs (measured 0.5e-3 .. 3.0e-2 vs 0.6e-3 .. 2.7e-2)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ref(obs, p):
    ln = lambda x, w, b: F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)
    x = ln(obs, p["ln0_w"], p["ln0_b"])
    x = ln(torch.relu(x @ p["w1"].t() + p["b1"]), p["ln1_w"], p["ln1_b"])
    x = ln(torch.relu(x @ p["w2"].t() + p["b2"]), p["ln2_w"], p["ln2_b"])
    return x @ p["wih"].t() + p["bih"]


@pytest.mark.parametrize("rows,D,gscale", [(1000, 47, 1.0), (128, 48, 1e-6), (37, 30, 1.0), (20000, 47, 1e-4)])
def test_agent_mlp_matches_float64(cuda, rows, D, gscale):
    from minimarl_b200.nn.agent_mlp import agent_mlp_gi
    g = torch.Generator().manual_seed(rows + D)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    p = dict(ln0_w=1 + 0.2 * rn(D), ln0_b=0.2 * rn(D), w1=rn(64, D) / D ** 0.5, b1=0.2 * rn(64), ln1_w=1 + 0.2 * rn(64),
             ln1_b=0.2 * rn(64), w2=rn(64, 64) / 8, b2=0.2 * rn(64), ln2_w=1 + 0.2 * rn(64), ln2_b=0.2 * rn(64),
             wih=rn(192, 64) / 8, bih=0.2 * rn(192))
    obs = torch.rand((rows, D), generator=g, dtype=torch.float64)
    w = rn(rows, 192) * gscale
    pr = {k: v.clone().requires_grad_() for k, v in p.items()}
    ref = _ref(obs, pr)
    (ref * w).sum().backward()
    pc = {k: v.float().to(cuda).requires_grad_() for k, v in p.items()}
    out = agent_mlp_gi(obs.float().to(cuda), (pc["ln0_w"], pc["ln0_b"]), (pc["w1"], pc["b1"]), (pc["ln1_w"], pc["ln1_b"]),
                       (pc["w2"], pc["b2"]), (pc["ln2_w"], pc["ln2_b"]), pc["wih"], pc["bih"])
    (out * w.float().to(cuda)).sum().backward()
    assert float((out.detach().cpu().double() - ref.detach()).abs().max()) < 1e-2
    # the same layers in eager PyTorch with TF32 matmuls: the precision class the fused op claims
    pe = {k: v.float().to(cuda).requires_grad_() for k, v in p.items()}
    flag = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        (_ref(obs.float().to(cuda), pe) * w.float().to(cuda)).sum().backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = flag
    rel = lambda q, k: float((q[k].grad.cpu().double() - pr[k].grad).norm() / (pr[k].grad.norm() + 1e-300))
    for k in p:
        if k in ("ln2_w", "ln2_b", "wih", "bih"):
            assert rel(pc, k) < 3e-3, (k, rel(pc, k))
        else:
            assert rel(pc, k) < 5e-2 and rel(pc, k) < 2 * rel(pe, k) + 1e-3, (k, rel(pc, k), rel(pe, k))


def test_agent_network_live_pass_uses_the_fused_op_and_matches_eager(cuda):
    """AgentQFunction.forward under autograd: fused row-parallel layers + recurrence kernels vs the eager module chain (TF32
    matmuls) on the same weights: q 5e-3 abs, gradients 5e-2 of the norm (two TF32-class computations of the same function)."""
    import argparse
    from minimarl_b200.nn.agent_q import AgentQFunction
    args = argparse.Namespace(hidden_size=64, use_rnn_layer=True, layer_N=1, recurrent_N=1, use_feature_normalization=True,
                              use_ReLU=True, use_orthogonal=True, gain=0.01, use_conv1d=False, fused_precision="tf32")
    torch.manual_seed(3)
    net = AgentQFunction(args, 47, 5, cuda)
    with torch.no_grad():
        for q in net.parameters():
            if q.dim() == 1:
                q.add_(0.1 * torch.randn_like(q))
        net.q.action_out.weight.mul_(30.0)
    T, R = 9, 300
    obs = torch.rand((T, R, 47), device=cuda)
    h0 = torch.randn((1, R, 64), device=cuda) * 0.3
    w = torch.randn((T, R, 5), device=cuda)
    flags = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
    try:
        outs = []
        for fused in (True, False):
            net.rnn.fused_mlp = fused
            net.zero_grad()
            q, h = net(obs, h0)
            (q * w).sum().backward()
            outs.append((q.detach().clone(), h.detach().clone(), {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = flags
        net.rnn.fused_mlp = True
    (qa, ha, ga), (qb, hb, gb) = outs
    assert float((qa - qb).abs().max()) < 5e-3 and float((ha - hb).abs().max()) < 5e-3
    assert set(ga) == set(gb)
    for n in ga:
        rel = float((ga[n] - gb[n]).norm() / (gb[n].norm() + 1e-30))
        assert rel < 5e-2, (n, rel)


@pytest.mark.parametrize("M,H,NA", [(1, 64, 5), (77, 64, 5), (8 * 1184 * 8 + 13, 64, 5), (333, 32, 5), (100, 64, 8), (50, 64, 3)])
def test_ln_head_matches_float64(cuda, M, H, NA):
    """mm_ln_head_fwd / _bwd (LayerNorm + the Q head, fp32 arithmetic) against the same two layers in float64: outputs 2e-6 of
    their scale, dh and the four parameter gradients 1e-5 relative (fp32 sums over the rows, atomics across CTAs)."""
    from minimarl_b200.nn.agent_mlp import ln_head
    torch.manual_seed(M + NA)
    h = torch.randn(M, H) * 2 + 0.3
    ps = [1 + 0.2 * torch.randn(H), 0.2 * torch.randn(H), torch.randn(NA, H) / 8, 0.1 * torch.randn(NA)]
    dq = torch.randn(M, NA) * (torch.rand(M, NA) < 0.3)                     # mostly one action per row carries a gradient
    hm = h.to(cuda).requires_grad_(True)
    pm = [p.to(cuda).requires_grad_(True) for p in ps]
    q = ln_head(hm, *pm)
    q.backward(dq.to(cuda))
    hr = h.double().requires_grad_(True)
    pr = [p.double().requires_grad_(True) for p in ps]
    qr = torch.nn.functional.linear(torch.nn.functional.layer_norm(hr, (H,), pr[0], pr[1], 1e-5), pr[2], pr[3])
    qr.backward(dq.double())
    assert (q.detach().cpu().double() - qr.detach()).abs().max().item() < 2e-6 * max(1.0, qr.detach().abs().max().item())
    rel = lambda a, b: ((a.detach().cpu().double() - b).norm() / b.norm().clamp(min=1e-30)).item()
    assert rel(hm.grad, hr.grad) < 1e-5
    for a, b in zip(pm, pr):
        assert rel(a.grad, b.grad) < 1e-5, rel(a.grad, b.grad)
