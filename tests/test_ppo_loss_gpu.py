"""GPU parity: the PPO loss kernels against the eager formulas of R_MAPPO.ppo_update / cal_value_loss
(mappo/algorithms/ramppo_network.py:70-171) evaluated with autograd in float64 on the same inputs.
Tolerance (fp32 kernels vs float64 reference): scalars 2e-6 relative, per-row gradients 1e-6 absolute on O(1e-4) values."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _huber(e, d):
    a = (abs(e) <= d).double()
    b = (abs(e) > d).double()
    return a * e ** 2 / 2 + b * d * (abs(e) - d / 2)


@pytest.mark.parametrize("use_active", [True, False])
def test_actor_loss_and_gradient(cuda, use_active):
    from minimarl_b200.ops.ppo_loss import ppo_actor_loss
    g = torch.Generator().manual_seed(3)
    R, NA, clip, coef = 5000, 5, 0.2, 0.01
    logits = (torch.randn((R, NA), generator=g) * 2).requires_grad_()
    act = torch.randint(0, NA, (R, 1), generator=g).float()
    with torch.no_grad():
        logp_now = torch.log_softmax(logits, -1).gather(1, act.long())
    old = logp_now + torch.randn((R, 1), generator=g) * 0.3
    old[::7] = logp_now[::7]                       # imp == 1 exactly: both surrogates tie, inside the clip range
    adv = torch.randn((R, 1), generator=g)
    adv[::11] = 0.0
    active = (torch.rand((R, 1), generator=g) > 0.25).float() if use_active else None
    # reference, float64
    l64 = logits.detach().double().requires_grad_()
    dist = torch.distributions.Categorical(logits=l64)
    lp = dist.log_prob(act.squeeze(-1).long()).unsqueeze(-1)
    imp = torch.exp(lp - old.double())
    s1, s2 = imp * adv.double(), torch.clamp(imp, 1 - clip, 1 + clip) * adv.double()
    if use_active:
        pl = (-torch.sum(torch.min(s1, s2), dim=-1, keepdim=True) * active.double()).sum() / active.double().sum()
        ent = (dist.entropy() * active.double().squeeze(-1)).sum() / active.double().sum()
    else:
        pl = -torch.sum(torch.min(s1, s2), dim=-1, keepdim=True).mean()
        ent = dist.entropy().mean()
    (pl - ent * coef).backward()
    # kernels
    lg = logits.detach().to(cuda).requires_grad_()
    dev = lambda x: None if x is None else x.to(cuda)
    pl_k, ent_k, ratio_k = ppo_actor_loss(lg, dev(act), dev(old), dev(adv), dev(active), clip)
    (pl_k - ent_k * coef).backward()
    assert abs(float(pl_k) - float(pl)) < 2e-6 * max(1, abs(float(pl))) and abs(float(ent_k) - float(ent)) < 2e-6 * float(ent)
    assert abs(float(ratio_k) - float(imp.mean())) < 2e-6 * float(imp.mean())
    assert float((lg.grad.cpu().double() - l64.grad).abs().max()) < 1e-6
    assert float(l64.grad.abs().max()) > 1e-5


@pytest.mark.parametrize("use_huber,use_clipped,use_active,use_norm",
                         [(True, True, True, True), (False, True, False, True), (True, False, True, False),
                          (False, False, False, False)])
def test_value_loss_and_gradient(cuda, use_huber, use_clipped, use_active, use_norm):
    from minimarl_b200.ops.ppo_loss import ppo_value_loss
    g = torch.Generator().manual_seed(5)
    R, clip, delta = 7001, 0.2, 10.0
    values = (torch.randn((R, 1), generator=g) * 3)
    vp = values + torch.randn((R, 1), generator=g) * 0.3
    vp[::5] = values[::5]                          # unclipped and equal: both errors tie
    ret = torch.randn((R, 1), generator=g) * 8 + 1
    active = (torch.rand((R, 1), generator=g) > 0.3).float() if use_active else None
    mean, var = torch.tensor(0.7), torch.tensor(3.1)
    v64 = values.double().requires_grad_()
    r64 = (ret.double() - mean.double()) / torch.sqrt(var.double()) if use_norm else ret.double()
    vc = vp.double() + (v64 - vp.double()).clamp(-clip, clip)
    f = (lambda e: _huber(e, delta)) if use_huber else (lambda e: e ** 2 / 2)
    lo, lc = f(r64 - v64), f(r64 - vc)
    loss = torch.max(lo, lc) if use_clipped else lo
    ref = (loss * active.double()).sum() / active.double().sum() if use_active else loss.mean()
    (ref * 2.5).backward()
    dev = lambda x: None if x is None else x.to(cuda)
    vk = values.to(cuda).requires_grad_()
    norm = torch.stack([mean, torch.rsqrt(var)]).to(cuda) if use_norm else None
    out = ppo_value_loss(vk, dev(vp), dev(ret), dev(active), norm, clip, delta, use_huber, use_clipped)
    (out * 2.5).backward()
    assert abs(float(out) - float(ref)) < 3e-6 * abs(float(ref))
    assert float((vk.grad.cpu().double() - v64.grad).abs().max()) < 2e-6 * float(v64.grad.abs().max()) + 1e-9
