"""GPU parity: mixer / selection / TD kernels and one full QMIX update vs values recorded from the
reference's QMixer, VDNMixer and QMix.train_policy_on_batch (tests/golden/offpolicy_qmix.npz).
Tolerances (fp32): mixer outputs 2e-5 abs; loss 1e-5 rel; gradients 2e-5 abs + 1e-3 rel; priorities 1e-5."""
import argparse
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from minimarl_b200.envs.spaces import Box, Discrete

pytestmark = pytest.mark.gpu


def _args(**over):
    a = argparse.Namespace(
        hidden_size=64, use_rnn_layer=True, use_orthogonal=True, gain=0.01, layer_N=1, recurrent_N=1,
        use_feature_normalization=True, use_ReLU=True, use_conv1d=False, prev_act_inp=False, batch_size=6,
        mixer_hidden_dim=32, hypernet_hidden_dim=64, hypernet_layers=2, use_per=True, per_eps=1e-6, per_nu=0.9,
        use_huber_loss=False, use_popart=False, lr=5e-4, tau=0.005, opti_eps=1e-5, episode_length=12,
        use_same_share_obs=True, use_double_q=True, gamma=0.99, max_grad_norm=10.0, epsilon_start=1.0,
        epsilon_finish=0.05, epsilon_anneal_time=50000, seed=1, weight_decay=0, fused_precision="fp32")
    for k, v in over.items():
        setattr(a, k, v)
    return a


def _g():
    return np.load(os.path.join(GOLDEN, "offpolicy_qmix.npz"))


def _load(module, g, prefix):
    module.load_state_dict({k[len(prefix):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix)})


def test_qmixer_and_vdn_forward_match_reference(cuda):
    from minimarl_b200.nn import QMixer, VDNMixer
    g = _g()
    mixer = QMixer(_args(), 2, 94, cuda)
    _load(mixer, g, "mix.")
    assert sum(p.numel() for p in mixer.parameters()) == 27585
    q = torch.from_numpy(g["mixer_in_q"]).to(cuda)
    s = torch.from_numpy(g["share_obs"][:-1]).to(cuda)
    with torch.no_grad():
        out = mixer(q, s)                                               # fused kernel
    assert out.shape == g["mixer_out"].shape
    assert np.abs(out.cpu().numpy() - g["mixer_out"]).max() < 2e-5
    out_eager = mixer(q.clone().requires_grad_(True), s)                # autograd path, same maths
    assert np.abs(out_eager.detach().cpu().numpy() - g["mixer_out"]).max() < 2e-5
    with torch.no_grad():
        v = VDNMixer(_args(), 2, 94, cuda)(q, s)
    assert np.array_equal(v.cpu().numpy(), g["vdn_out"])


def test_mixer_ragged_rows(cuda):
    """Row counts that are not a multiple of the 64-row tile, more tiles than SMs."""
    from minimarl_b200.nn import QMixer
    torch.manual_seed(0)
    mixer = QMixer(_args(), 2, 94, cuda)
    with torch.no_grad():
        for p in mixer.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    for T, B in [(1, 1), (3, 50), (100, 133)]:
        q = torch.randn((T, B, 2), device=cuda)
        s = torch.rand((T, B, 94), device=cuda)
        with torch.no_grad():
            fused = mixer(q, s)
        eager = mixer(q.clone().requires_grad_(True), s).detach()
        assert (fused - eager).abs().max().item() < 5e-5 * max(1.0, eager.abs().max().item())


def test_tensor_core_mixer_matches_fp32_and_reference(cuda):
    """mm_qmixer_tc_forward (tcgen05, TF32 operands): the reference-recorded mixer output within 2e-3 of its scale, and the
    fp32 kernel on ragged row counts (1 row, partial tiles, more tiles than SMs)."""
    from minimarl_b200.nn import QMixer
    torch.backends.cuda.matmul.allow_tf32 = False
    g = _g()
    mixer = QMixer(_args(fused_precision="tf32"), 2, 94, cuda)
    assert mixer.tc_ok
    _load(mixer, g, "mix.")
    q, s = torch.from_numpy(g["mixer_in_q"]).to(cuda), torch.from_numpy(g["share_obs"][:-1]).to(cuda)
    with torch.no_grad():
        out = mixer(q, s)
    scale = np.abs(g["mixer_out"]).max()
    assert np.abs(out.cpu().numpy() - g["mixer_out"]).max() < 2e-3 * scale
    torch.manual_seed(1)
    ref = QMixer(_args(fused_precision="fp32"), 2, 94, cuda)
    with torch.no_grad():
        for p in ref.parameters():
            p.add_(0.05 * torch.randn_like(p))
    mixer.load_state_dict(ref.state_dict())
    for T, B in [(1, 1), (3, 50), (7, 128), (100, 333), (40, 4096)]:
        q = torch.randn((T, B, 2), device=cuda)
        s = torch.rand((T, B, 94), device=cuda)
        with torch.no_grad():
            a, b = mixer(q, s), ref(q, s)
        assert a.shape == b.shape
        # Q_tot is a sum of 32 products of O(1) magnitude plus b2: the TF32 error scales with those terms, not with the sum
        assert (a - b).abs().max().item() < 5e-3 * max(1.0, b.abs().max().item()), (T, B)


def test_full_qmix_update_matches_reference(cuda):
    from minimarl_b200.algorithms import QMix, QMixPolicy
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = _g()
    args = _args()
    T, B, A, D, nA = 12, 6, 2, 47, 5
    pconf = {"cent_obs_dim": A * D, "cent_act_dim": A * nA, "obs_space": Box(np.zeros(D), np.ones(D)),
             "share_obs_space": Box(np.zeros(A * D), np.ones(A * D)), "act_space": Discrete(nA)}
    policy = QMixPolicy(args, {"args": args, "device": cuda, "num_agents": A}, pconf)
    trainer = QMix(args, A, B, {"policy_0": policy}, lambda a: "policy_0", device=cuda, episode_length=T)
    _load(policy.q_network, g, "q.")
    _load(trainer.target_policies["policy_0"].q_network, g, "tq.")
    _load(trainer.mixer, g, "mix.")
    _load(trainer.target_mixer, g, "tmix.")
    batch = tuple({"policy_0": g[k]} for k in ["obs", "share_obs", "acts", "rewards", "dones", "dones_env"]) + \
        ({"policy_0": None}, g["weights"], np.arange(B))
    before = [p.detach().clone() for p in trainer.parameters]
    info, new_prio, idxes = trainer.train_policy_on_batch(batch)
    assert abs(info["loss"].item() - float(g["train_loss"])) < 1e-5 * max(1.0, abs(float(g["train_loss"])))
    assert abs(float(info["grad_norm"]) - float(g["train_grad_norm"])) < 1e-4 * float(g["train_grad_norm"])
    assert abs(info["Q_tot"].item() - float(g["train_Q_tot"])) < 1e-5
    assert np.allclose(new_prio.cpu().numpy(), g["train_new_priorities"], rtol=1e-5, atol=1e-6)
    assert np.array_equal(idxes, np.arange(B))
    names = [n for n, _ in policy.q_network.named_parameters()] + ["mix." + n for n, _ in trainer.mixer.named_parameters()]
    checked = 0
    for n, p in zip(names, trainer.parameters):
        key = "grad." + n
        if key in g.files:
            assert np.allclose(p.grad.cpu().numpy(), g[key], rtol=1e-3, atol=2e-5), n
            checked += 1
        else:
            assert p.grad is None or not p.grad.any(), n               # the unused fc_h duplicate (a zero slice of FlatAdam's bucket)
    assert checked >= 30
    assert any((a - b).abs().max().item() > 0 for a, b in zip(before, trainer.parameters))
    trainer.soft_target_updates()
    trainer.hard_target_updates()
    trainer.prep_rollout(); trainer.prep_training()


def test_full_qmix_update_tf32_matches_reference(cuda):
    """The benchmarked learner precision: TF32 operands (tensor-core target / greedy / mixer passes, TF32 cuBLAS in the live
    pass), fp32 accumulation.  Same fixture as the fp32 test; stated tolerances: loss 3e-3 rel, grad norm 2e-3 rel, Q_tot 2e-3,
    priorities 2e-3 of the largest, every gradient tensor 1e-2 of its norm (measured: 6e-4, 3e-4, 3e-4, 4e-4, <= 2.6e-3)."""
    from minimarl_b200.algorithms import QMix, QMixPolicy
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        g = _g()
        args = _args(fused_precision="tf32")
        T, B, A, D, nA = 12, 6, 2, 47, 5
        pconf = {"cent_obs_dim": A * D, "cent_act_dim": A * nA, "obs_space": Box(np.zeros(D), np.ones(D)),
                 "share_obs_space": Box(np.zeros(A * D), np.ones(A * D)), "act_space": Discrete(nA)}
        policy = QMixPolicy(args, {"args": args, "device": cuda, "num_agents": A}, pconf)
        trainer = QMix(args, A, B, {"policy_0": policy}, lambda a: "policy_0", device=cuda, episode_length=T)
        _load(policy.q_network, g, "q.")
        _load(trainer.target_policies["policy_0"].q_network, g, "tq.")
        _load(trainer.mixer, g, "mix.")
        _load(trainer.target_mixer, g, "tmix.")
        batch = tuple({"policy_0": g[k]} for k in ["obs", "share_obs", "acts", "rewards", "dones", "dones_env"]) + \
            ({"policy_0": None}, g["weights"], np.arange(B))
        info, new_prio, _ = trainer.train_policy_on_batch(batch)
        assert abs(info["loss"].item() - float(g["train_loss"])) < 3e-3 * abs(float(g["train_loss"]))
        assert abs(float(info["grad_norm"]) - float(g["train_grad_norm"])) < 2e-3 * float(g["train_grad_norm"])
        assert abs(info["Q_tot"].item() - float(g["train_Q_tot"])) < 2e-3
        want = g["train_new_priorities"]
        assert np.abs(new_prio.cpu().numpy() - want).max() < 2e-3 * np.abs(want).max()
        names = [n for n, _ in policy.q_network.named_parameters()] + ["mix." + n for n, _ in trainer.mixer.named_parameters()]
        checked = 0
        for n, p in zip(names, trainer.parameters):
            key = "grad." + n
            if key in g.files:
                err = np.linalg.norm(p.grad.cpu().numpy() - g[key]) / np.linalg.norm(g[key])
                assert err < 1e-2, (n, err)
                checked += 1
        assert checked >= 30
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32


def test_q_select_and_td_kernels_against_torch(cuda):
    from minimarl_b200.ops import q_select, qmix_td_loss
    T, B, A, NA = 20, 37, 2, 5
    gen = torch.Generator(device=cuda).manual_seed(0)
    q_all = torch.randn((T + 1, A * B, NA), device=cuda, generator=gen)
    tq_all = torch.randn((T + 1, A * B, NA), device=cuda, generator=gen)
    idx = torch.randint(0, NA, (T, A * B), device=cuda, generator=gen)
    acts = torch.nn.functional.one_hot(idx, NA).float()
    q_taken, next_q, act_idx, greedy = q_select(q_all, tq_all, acts, T, B, A, double_q=True, want_greedy=True)
    assert torch.equal(act_idx, idx) and torch.equal(greedy, q_all.argmax(-1))
    want = torch.gather(q_all[:-1], 2, idx.unsqueeze(-1)).squeeze(-1).view(T, A, B).permute(0, 2, 1)
    assert torch.equal(q_taken, want)
    want_next = torch.gather(tq_all, 2, greedy.unsqueeze(-1)).squeeze(-1)[1:].view(T, A, B).permute(0, 2, 1)
    assert torch.equal(next_q, want_next)
    _, next_max, _, _ = q_select(q_all, tq_all, acts, T, B, A, double_q=False)
    assert torch.equal(next_max, tq_all.max(-1)[0][1:].view(T, A, B).permute(0, 2, 1))
    # Q11: the reference stores the action INDEX broadcast over the one-hot dim -> all ties -> index 0
    tied = idx.unsqueeze(-1).expand(T, A * B, NA).float().contiguous()
    _, _, tied_idx, _ = q_select(q_all, None, tied, T, B, A)
    assert int(tied_idx.abs().max()) == 0

    q_tot = torch.randn((T, B, 1), device=cuda, generator=gen, requires_grad=True)
    nq = torch.randn((T, B, 1), device=cuda, generator=gen)
    rew = torch.randn((T, B, 1), device=cuda, generator=gen)
    done = (torch.rand((T, B, 1), device=cuda, generator=gen) < 0.1).float()
    w = torch.rand(B, device=cuda, generator=gen)
    loss, prio, qmean, denom = qmix_td_loss(q_tot, nq, rew, done, w, 0.99, 0.9, 1e-6)
    loss.backward()
    q2 = q_tot.detach().clone().requires_grad_(True)
    bad = torch.cat((torch.zeros(1, B, 1, device=cuda), done[:-1]))
    err = (q2 - (rew + (1 - done) * 0.99 * nq)) * (1 - bad)
    ref = ((err ** 2).sum(0).flatten() * w).sum() / (1 - bad).sum()
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-5 * abs(ref.item())
    assert torch.allclose(q_tot.grad, q2.grad, rtol=1e-4, atol=1e-7)
    td = err.abs().detach()
    assert torch.allclose(prio, (0.1 * td.mean(0) + 0.9 * td.max(0)[0]).flatten() + 1e-6, rtol=1e-5)
    assert abs(qmean.item() - (q2.detach() * (1 - bad)).mean().item()) < 1e-6


def test_graphed_update_replays_the_eager_update(cuda):
    """GraphedUpdate (PER draw + gather, train_policy_on_batch on both streams, priority write-back, soft target update captured
    as ONE CUDA graph) against the same four updates launched eagerly: same PER draws, so the same batches -- parameters after four
    Adam steps within 2e-3 of their norm (TF32-class kernels with atomics in the reductions: not bit-identical), priority trees
    within 1e-3 of the total mass."""
    import copy
    from minimarl_b200.algorithms import QMix, QMixPolicy
    from minimarl_b200.algorithms.qmix import GraphedUpdate
    from minimarl_b200.buffers import PrioritizedRecReplayBuffer
    flags = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
    try:
        T, A, D, NA, S, B = 10, 2, 47, 5, 96, 32
        args = _args(fused_precision="tf32", episode_length=T, batch_size=B)
        info = {"policy_0": {"obs_space": Box(np.zeros(D), np.ones(D)), "share_obs_space": Box(np.zeros(A * D), np.ones(A * D)),
                             "act_space": Discrete(NA)}}
        pconf = {"cent_obs_dim": A * D, "cent_act_dim": A * NA, "obs_space": info["policy_0"]["obs_space"],
                 "share_obs_space": info["policy_0"]["share_obs_space"], "act_space": info["policy_0"]["act_space"]}
        results = []
        for graphed in (False, True):
            torch.manual_seed(7)
            buf = PrioritizedRecReplayBuffer(0.6, info, {"policy_0": [0, 1]}, S, T, True, False, device=cuda)
            pb = buf.policy_buffers["policy_0"]
            g = torch.Generator(device=cuda).manual_seed(3)
            pb.obs.copy_(torch.rand(pb.obs.shape, device=cuda, generator=g))
            pb.share_obs.copy_(pb.obs.reshape(T + 1, S, A * D))
            pb.acts.copy_(torch.nn.functional.one_hot(torch.randint(0, NA, (T, S, A), device=cuda, generator=g), NA).float())
            pb.rewards.copy_(torch.randn(pb.rewards.shape, device=cuda, generator=g))
            pb.dones.zero_(); pb.dones_env.zero_(); pb.dones_env[-1] = 1
            pb.filled_i, pb.current_i = S, 0
            buf.update_priorities(torch.arange(S, device=cuda), torch.rand(S, device=cuda, generator=g) * 0.99 + 0.01, "policy_0")
            buf.seed(11)
            policy = QMixPolicy(args, {"args": args, "device": cuda, "num_agents": A}, pconf)
            trainer = QMix(args, A, B, {"policy_0": policy}, lambda a: "policy_0", device=cuda, episode_length=T)
            losses = []
            if graphed:
                gu = GraphedUpdate(trainer, buf, B)
                for _ in range(4):
                    losses.append(float(gu(0.4)["loss"]))
            else:
                for _ in range(4):
                    u = torch.rand(B, dtype=torch.float64, device=cuda, generator=buf._gen)
                    sample = buf.sample(B, beta=0.4, p_id="policy_0", uniforms=u)
                    info_, prio, idx = trainer.train_policy_on_batch(sample)
                    buf.update_priorities(idx, prio, "policy_0")
                    trainer.soft_target_updates()
                    losses.append(float(info_["loss"]))
            # the rollout / target kernels read PACKED copies of the parameters, cached on the tensors' version counters: the
            # optimiser kernel writes through raw pointers and a graph replay runs no Python, so both must bump them
            from minimarl_b200.nn.agent_q import pack_agentq_tc
            for net in (policy.q_network, trainer.target_policies["policy_0"].q_network):
                assert net.tc_ok
                fresh = pack_agentq_tc(net.state_dict(), net.input_dim, net.hidden_size, net.act_dim, cuda)
                assert torch.equal(net.packed_tc(), fresh), "stale packed parameters"
            results.append((losses, [p.detach().clone() for p in trainer.parameters],
                            buf._it_sums["policy_0"]._value.clone()))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = flags
    (la, pa, ta), (lb, pb_, tb) = results
    assert np.allclose(la, lb, rtol=2e-3), (la, lb)
    for x, y in zip(pa, pb_):
        assert float((x - y).norm()) <= 2e-3 * float(y.norm()) + 1e-6
    assert float((ta - tb).abs().max()) <= 1e-3 * float(tb[1])


def _mixer_f64(P, q, sd):
    """q_mixer.py:68-94 in float64; also returns every pre-activation that sits under a ReLU / abs kink."""
    M, SD = sd.shape
    lin = lambda x, n: x @ P[n + ".weight"].t() + P[n + ".bias"]
    h1a, h2a, hba = lin(sd, "hyper_w1.0"), lin(sd, "hyper_w2.0"), lin(sd, "hyper_b2.0")
    w1raw, w2raw = lin(torch.relu(h1a), "hyper_w1.2"), lin(torch.relu(h2a), "hyper_w2.2")
    b1 = lin(sd, "hyper_b1").view(M, 1, 32)
    b2 = lin(torch.relu(hba), "hyper_b2.2").view(M, 1, 1)
    pre = torch.matmul(q.view(M, 1, 2), w1raw.abs().view(M, 2, 32)) + b1
    out = (torch.matmul(torch.nn.functional.elu(pre), w2raw.abs().view(M, 32, 1)) + b2).reshape(-1)
    return out, torch.cat([h1a, h2a, hba, w1raw, w2raw], dim=1)


@pytest.mark.parametrize("M,SD", [(1, 94), (300, 94), (128 * 150 + 77, 94), (4099, 46)])
def test_live_mixer_kernels_match_float64(cuda, M, SD):
    """The differentiated mixer (forward kernel + mm_qmixer_bwd with the weight gradients inside) against the same function in
    float64 on the CPU (q_mixer.py:68-94 under autograd).  Operands carry an 11-bit significand (TF32 forward, fp16 tiles in the
    backward): Q_tot to 2e-3 of its scale, dq and every parameter gradient to 4e-3 relative (Frobenius).  The gradient of
    ReLU / abs jumps at 0, so any rounding of the forward picks a side there: the rows are drawn so that no pre-activation lies
    within 2e-3 of a kink (rounding moves them by ~3e-4); the statistical effect of the kinks at full size is covered by the
    update-level TF32 test above."""
    from minimarl_b200.nn import QMixer
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.manual_seed(M + SD)
    mixer = QMixer(_args(fused_precision="tf32"), 2, SD, cuda)
    with torch.no_grad():
        for p in mixer.parameters():
            p.add_(torch.randn_like(p) * 0.05)                            # biases are initialised to zero
    ref = {k: v.detach().double().cpu() for k, v in mixer.state_dict().items()}
    q, s = torch.randn(6 * M + 64, 1, 2) * 3, torch.rand(6 * M + 64, 1, SD)
    with torch.no_grad():
        _, kinks = _mixer_f64(ref, q.double(), s.double().reshape(-1, SD))
    keep = (kinks.abs().min(dim=1).values > 2e-3).nonzero().reshape(-1)[:M]
    assert keep.numel() == M
    q, s = q[keep].contiguous(), s[keep].contiguous()
    w = torch.randn(M) * 1e-4                                             # loss-mean sized upstream gradients
    q_m = q.to(cuda).requires_grad_(True)
    out = mixer(q_m, s.to(cuda))
    (out.reshape(-1) * w.to(cuda)).sum().backward()

    P = {k: v.clone().requires_grad_(True) for k, v in ref.items()}
    q_r = q.double().requires_grad_(True)
    out_r, _ = _mixer_f64(P, q_r, s.double().reshape(M, SD))
    (out_r * w.double()).sum().backward()

    scale = out_r.detach().abs().max().item()
    assert (out.detach().reshape(-1).cpu().double() - out_r.detach()).abs().max().item() < 2e-3 * max(scale, 1.0)
    rel = lambda a, b: ((a.double().cpu() - b).norm() / b.norm().clamp(min=1e-30)).item()
    assert rel(q_m.grad.reshape(M, 2), q_r.grad.reshape(M, 2)) < 4e-3
    for name, p in mixer.named_parameters():
        assert p.grad is not None and rel(p.grad, P[name].grad) < 4e-3, (name, rel(p.grad, P[name].grad))
