"""GPU parity: the fused PPO trunk op (mm_ppo_trunk2_fwd / mm_ppo_trunk2_bwd: fp16 operands -- the 11-bit significand of
tf32 -- on tcgen05 with fp32 accumulation) against the eager fp32 R_Actor.evaluate_actions / R_Critic.forward on CPU (the
reference's layer chain, mask-segmented GRU).  Stated tolerances:
  outputs                                                    1e-2 abs
  gradients of the GRU, its LayerNorm, the head, LayerNorm 2  3e-3 of the gradient's norm (measured 2e-4 .. 1.1e-3)
  gradients below the second LayerNorm (fc1, fc2, LN 0 / 1)   8e-2 of the gradient's norm (measured 0.7e-2 .. 6.3e-2)
The lower layers sit behind two ReLU + LayerNorm backward passes, whose projection (dx - mean(dx) - xh mean(dx xh)) cancels the
component of dx along the normalised activation: what is left carries the operand rounding (2^-12) amplified by |dx| / |dx_perp|.
The eager PyTorch path with TF32 matmuls and the first-generation tf32 kernels show the same amplification on these inputs
(1.2e-2 .. 3.5e-2); the fp32 reference numbers of the update itself are pinned by tests/test_mappo_gpu.py (eager fp32 vs the
reference fixture, 2e-4)."""

# parameters whose gradient does not pass through a LayerNorm backward below the GRU: tight tolerance
_UPPER = ("rnn.", "act.", "v_out", "base.mlp.fc2.0.2")


def _grad_tol(name):
    return 3e-3 if name.startswith(_UPPER) else 8e-2
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _nets(cuda, seed=0):
    from minimarl_b200.algorithms import R_MAPPOPolicy
    from minimarl_b200.envs.spaces import Box, Discrete
    from test_mappo_gpu import _args, _g
    torch.manual_seed(seed)
    args = _args(_g(), "tf32", hidden_size=32)
    space = Box(np.zeros(47), np.ones(47))
    policy = R_MAPPOPolicy(args, space, space, Discrete(5), device=cuda)
    with torch.no_grad():
        policy.actor.act.action_out.linear.weight.mul_(40.0)
        for net in (policy.actor, policy.critic):
            for p in net.parameters():
                if p.dim() == 1:
                    p.add_(0.2 * torch.randn_like(p))
    return policy


def _cpu(net):
    c = copy.deepcopy(net).cpu()
    c._fusable = False
    c.tpdv = dict(dtype=torch.float32, device="cpu")
    return c


def _rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-12))


@pytest.mark.parametrize("L,N", [(8, 300), (4, 128), (10, 77)])
def test_fused_chunk_pass_matches_eager_actor_and_critic(cuda, L, N):
    policy = _nets(cuda, seed=L)
    g = torch.Generator().manual_seed(100 + N)
    obs = torch.rand((L * N, 47), generator=g)
    h0 = torch.randn((N, 1, 32), generator=g) * 0.5
    masks = (torch.rand((L * N, 1), generator=g) > 0.2).float()
    act = torch.randint(0, 5, (L * N, 1), generator=g).float()
    wl, wv = torch.randn((L * N, 1), generator=g), torch.randn((L * N, 1), generator=g)
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        # actor
        ref, fused = _cpu(policy.actor), policy.actor
        lp_r, ent_r = ref.evaluate_actions(obs, h0, act, masks, active_masks=torch.ones(L * N, 1))
        (lp_r * wl).sum().add(ent_r * 3.0).backward()
        dev = lambda x: x.to(cuda)
        assert fused._chunk_fusable(dev(obs), dev(h0))
        lp_f, ent_f = fused.evaluate_actions(dev(obs), dev(h0), dev(act), dev(masks), active_masks=torch.ones(L * N, 1, device=cuda))
        (lp_f * dev(wl)).sum().add(ent_f * 3.0).backward()
        assert float((lp_f.detach().cpu() - lp_r.detach()).abs().max()) < 1e-2 and abs(float(ent_f.detach()) - float(ent_r.detach())) < 1e-3
        for (k, pr), pf in zip(ref.named_parameters(), fused.parameters()):
            if pr.grad is None:                      # mlp.fc_h: the template layer the reference clones and never calls
                assert pf.grad is None or not pf.grad.any(), k     # (a zero slice of FlatAdam's bucket)
                continue
            assert _rel(pf.grad.cpu(), pr.grad) < _grad_tol(k), (k, _rel(pf.grad.cpu(), pr.grad))
        # critic
        ref, fused = _cpu(policy.critic), policy.critic
        v_r, _ = ref(obs, h0, masks)
        (v_r * wv).sum().backward()
        v_f, _ = fused(dev(obs), dev(h0), dev(masks))
        (v_f * dev(wv)).sum().backward()
        assert float((v_f.detach().cpu() - v_r.detach()).abs().max()) < 1e-2
        for (k, pr), pf in zip(ref.named_parameters(), fused.parameters()):
            if pr.grad is not None:
                assert _rel(pf.grad.cpu(), pr.grad) < _grad_tol(k), (k, _rel(pf.grad.cpu(), pr.grad))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = False


def test_fused_update_matches_stepwise_update(cuda):
    """R_MAPPO.train with the fused trunk vs the eager chunk pass on the same buffer and permutations."""
    from minimarl_b200 import BatchedCheckers
    from minimarl_b200.runners import MappoRunner
    from test_mappo_gpu import _args, _g
    N, T = 64, 16
    infos = {}
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        for fused in (True, False):
            torch.manual_seed(5)
            args = _args(_g(), "tf32", max_step=T, batch_size=N, hidden_size=32, ppo_epoch=2, train_batch_size=1,
                         data_chunk_length=8, max_episodes=1, use_eval=False, eval_interval=2, use_linear_lr_decay=False, fused_update=fused,
                         seed=9)
            env = BatchedCheckers(N, max_steps=12, auto_reset=True, device=cuda)
            r = MappoRunner({"args": args, "train_env": env, "eval_env": env, "num_agents": 2, "device": cuda})
            r.warmup(); r.rollout(); r.compute()
            perms = [torch.randperm(N * T * 2 // 8, generator=torch.Generator().manual_seed(e)).to(cuda) for e in range(2)]
            r.trainer.prep_training()
            infos[fused] = {k: float(v) for k, v in r.trainer.train(r.buffer, perms=perms).items()}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = False
    for k in ("value_loss", "policy_loss", "dist_entropy", "actor_grad_norm", "critic_grad_norm", "ratio"):
        a, b = infos[True][k], infos[False][k]
        assert abs(a - b) <= 2e-2 * max(abs(b), 1e-3) + 1e-4, (k, a, b)


def test_chunks_read_in_place_equal_gathered_chunks(cuda):
    """recurrent_generator(lazy_obs=True) hands the trunk op row addresses instead of a gathered copy: same bits."""
    from minimarl_b200.buffers import SharedReplayBuffer
    from minimarl_b200.envs.spaces import Box, Discrete
    from minimarl_b200.nn.ppo_trunk import ChunkRows
    from test_mappo_gpu import _args, _g
    policy = _nets(cuda, seed=2)
    T, N, Lc = 24, 37, 8
    args = _args(_g(), "tf32", hidden_size=32, max_step=T, batch_size=N)
    space = Box(np.zeros(47), np.ones(47))
    buf = SharedReplayBuffer(args, 2, space, space, Discrete(5), device=cuda)
    buf.obs.uniform_(); buf.share_obs.copy_(buf.obs); buf.rnn_states.normal_(); buf.masks.bernoulli_(0.8)
    adv = torch.randn_like(buf.rewards)
    perm = torch.randperm(T * N * 2 // Lc, generator=torch.Generator().manual_seed(1)).to(cuda)
    lazy = next(iter(buf.recurrent_generator(adv, 1, Lc, perm=perm, lazy_obs=True)))
    full = next(iter(buf.recurrent_generator(adv, 1, Lc, perm=perm)))
    assert all(isinstance(x, ChunkRows) for x in lazy[:11])
    for a, b in zip(lazy[:11], full[:11]):                 # every field: the row map addresses exactly the gathered minibatch
        assert torch.equal(a.gather(), b)
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        with torch.no_grad():
            out_lazy = policy.actor._fused_chunks(lazy[1], lazy[2], lazy[7])
            out_full = policy.actor._fused_chunks(full[1], full[2], full[7])
    finally:
        torch.backends.cuda.matmul.allow_tf32 = False
    assert torch.equal(out_lazy, out_full)
    # the loss kernels through the same row map
    from minimarl_b200.ops.ppo_loss import ppo_actor_loss, ppo_value_loss
    logits = out_full.detach().clone().requires_grad_()
    logits2 = out_full.detach().clone().requires_grad_()
    a = ppo_actor_loss(logits, full[4], full[9], full[10], full[8], 0.2)
    b = ppo_actor_loss(logits2, lazy[4], lazy[9], lazy[10], lazy[8], 0.2)
    (a[0] - 0.01 * a[1]).backward(); (b[0] - 0.01 * b[1]).backward()
    assert all(torch.equal(x, y) for x, y in zip(a, b)) and torch.equal(logits.grad, logits2.grad)
    v = torch.randn(out_full.shape[0], 1, device=cuda)
    v1, v2 = v.clone().requires_grad_(), v.clone().requires_grad_()
    norm = torch.tensor([0.3, 1.7], device=cuda)
    la = ppo_value_loss(v1, full[5], full[6], full[8], norm, 0.2, 10.0, True, True)
    lb = ppo_value_loss(v2, lazy[5], lazy[6], lazy[8], norm, 0.2, 10.0, True, True)
    la.backward(); lb.backward()
    assert torch.equal(la, lb) and torch.equal(v1.grad, v2.grad)


def _trunk_ref(obs, h0, masks, p):
    """Plain PyTorch statement of the trunk (float64 on CPU): LayerNorm -> [Linear ReLU LayerNorm] x2 -> GRU with the state
    multiplied by masks[t] before every step -> LayerNorm -> Linear."""
    import torch.nn.functional as F
    L, R, D = obs.shape
    ln = lambda x, w, b: F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)
    x = ln(obs, p["ln0_w"], p["ln0_b"])
    x = ln(torch.relu(x @ p["w1"].t() + p["b1"]), p["ln1_w"], p["ln1_b"])
    x = ln(torch.relu(x @ p["w2"].t() + p["b2"]), p["ln2_w"], p["ln2_b"])
    h, outs = h0, []
    for t in range(L):
        hp = h * masks[t].unsqueeze(-1)
        gi, gh = x[t] @ p["wih"].t() + p["bih"], hp @ p["whh"].t() + p["bhh"]
        r = torch.sigmoid(gi[:, :32] + gh[:, :32])
        z = torch.sigmoid(gi[:, 32:64] + gh[:, 32:64])
        n = torch.tanh(gi[:, 64:] + r * gh[:, 64:])
        h = n + z * (hp - n)
        outs.append(ln(h, p["ln3_w"], p["ln3_b"]) @ p["wo"].t() + p["bo"])
    return torch.stack(outs)


@pytest.mark.parametrize("L,R,D,NO", [(3, 200, 30, 8), (8, 129, 47, 1), (5, 128, 48, 6), (2, 1, 9, 3)])
def test_trunk_op_general_shapes_against_plain_pytorch(cuda, L, R, D, NO):
    """Input widths below 47 (zero padding) and the full 48 (masks gathered separately instead of riding in the observation
    table), every head width 1..8, ragged and single-row tiles; float64 statement of the trunk as the reference."""
    from minimarl_b200.nn.ppo_trunk import PARAM_ORDER, ppo_trunk
    g = torch.Generator().manual_seed(L * 1000 + R + D + NO)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    p = dict(ln0_w=1 + 0.2 * rn(D), ln0_b=0.2 * rn(D), w1=rn(32, D) / D ** 0.5, b1=0.2 * rn(32), ln1_w=1 + 0.2 * rn(32),
             ln1_b=0.2 * rn(32), w2=rn(32, 32) / 32 ** 0.5, b2=0.2 * rn(32), ln2_w=1 + 0.2 * rn(32), ln2_b=0.2 * rn(32),
             wih=rn(96, 32) / 32 ** 0.5, whh=rn(96, 32) / 32 ** 0.5, bih=0.2 * rn(96), bhh=0.2 * rn(96),
             ln3_w=1 + 0.2 * rn(32), ln3_b=0.2 * rn(32), wo=rn(NO, 32) / 32 ** 0.5, bo=0.2 * rn(NO))
    obs, h0 = torch.rand((L, R, D), generator=g, dtype=torch.float64), 0.5 * rn(R, 32)
    masks = (torch.rand((L, R), generator=g) > 0.25).double()
    w = rn(L, R, NO)
    pr = {k: v.clone().requires_grad_() for k, v in p.items()}
    ref = _trunk_ref(obs, h0, masks, pr)
    (ref * w).sum().backward()
    pc = {k: v.float().to(cuda).requires_grad_() for k, v in p.items()}
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        out = ppo_trunk(obs.float().to(cuda), h0.float().to(cuda), masks.float().to(cuda), pc)
        (out * w.float().to(cuda)).sum().backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = False
    assert float((out.detach().cpu().double() - ref.detach()).abs().max()) < 2e-2
    for k in PARAM_ORDER:
        rel = _rel(pc[k].grad.cpu().double(), pr[k].grad)
        assert rel < (3e-3 if k in ("ln2_w", "ln2_b", "wih", "whh", "bih", "bhh", "ln3_w", "ln3_b", "wo", "bo") else 5e-2), (k, rel)


def test_trunk_generations_agree(cuda):
    """The first-generation op (tf32 tiles, G^T X on cuBLAS) and the second (fp16 tiles, weight gradients in the kernel) on the
    same inputs: independent implementations of the same function, outputs 1e-2, gradients 3e-2 of the norm."""
    from minimarl_b200.nn.ppo_trunk import PARAM_ORDER, ppo_trunk, ppo_trunk_v1
    g = torch.Generator().manual_seed(11)
    rn = lambda *s: torch.randn(*s, generator=g)
    L, R, D, NO = 8, 700, 47, 5
    p = dict(ln0_w=1 + 0.2 * rn(D), ln0_b=0.2 * rn(D), w1=rn(32, D) / D ** 0.5, b1=0.2 * rn(32), ln1_w=1 + 0.2 * rn(32),
             ln1_b=0.2 * rn(32), w2=rn(32, 32) / 32 ** 0.5, b2=0.2 * rn(32), ln2_w=1 + 0.2 * rn(32), ln2_b=0.2 * rn(32),
             wih=rn(96, 32) / 32 ** 0.5, whh=rn(96, 32) / 32 ** 0.5, bih=0.2 * rn(96), bhh=0.2 * rn(96),
             ln3_w=1 + 0.2 * rn(32), ln3_b=0.2 * rn(32), wo=rn(NO, 32) / 32 ** 0.5, bo=0.2 * rn(NO))
    obs, h0 = torch.rand((L, R, D), generator=g).to(cuda), (0.5 * rn(R, 32)).to(cuda)
    masks = (torch.rand((L, R), generator=g) > 0.25).float().to(cuda)
    w = (rn(L, R, NO) * 1e-5).to(cuda)                    # PPO-sized head gradients: exercises the power-of-two gradient scale
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        outs, grads = [], []
        for fn in (ppo_trunk, ppo_trunk_v1):
            pc = {k: v.to(cuda).requires_grad_() for k, v in p.items()}
            out = fn(obs, h0, masks, pc)
            (out * w).sum().backward()
            outs.append(out.detach())
            grads.append({k: pc[k].grad for k in PARAM_ORDER})
    finally:
        torch.backends.cuda.matmul.allow_tf32 = False
    assert float((outs[0] - outs[1]).abs().max()) < 1e-2
    for k in PARAM_ORDER:
        assert torch.isfinite(grads[0][k]).all()
        assert _rel(grads[0][k], grads[1][k]) < 3e-2, (k, _rel(grads[0][k], grads[1][k]))
