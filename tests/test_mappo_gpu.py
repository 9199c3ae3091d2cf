"""GPU parity: MAPPO policy / trainer mirror vs values recorded from the reference's R_MAPPOPolicy and
R_MAPPO.train (tests/golden/mappo_train.npz).  Tolerances: fp32 fused step 3e-5 abs; tensor-core step 4e-3;
train_info 2e-4 relative (sums of fp32 reductions in a different order); parameters after 4 Adam steps 5e-4 abs
(Adam turns a sign flip of a ~0 gradient element into a +-lr step)."""
import argparse
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from minimarl_b200.envs.spaces import Box, Discrete

pytestmark = pytest.mark.gpu


def _g():
    return np.load(os.path.join(GOLDEN, "mappo_train.npz"))


def _args(g, precision="fp32", **over):
    a = argparse.Namespace(
        hidden_size=int(g["hidden_size"]), layer_N=1, recurrent_N=1, use_feature_normalization=True, use_ReLU=True,
        use_orthogonal=True, gain=0.01, use_conv1d=False, use_recurrent_policy=True, use_naive_recurrent_policy=False,
        use_policy_active_masks=True, use_value_active_masks=True, use_popart=False, use_valuenorm=True,
        use_huber_loss=True, use_max_grad_norm=True, use_clipped_value_loss=True, use_gae=True,
        ppo_epoch=int(g["ppo_epoch"]), train_batch_size=int(g["train_batch_size"]),
        data_chunk_length=int(g["data_chunk_length"]), clip_param=float(g["clip_param"]),
        entropy_coef=float(g["entropy_coef"]), value_loss_coef=float(g["value_loss_coef"]),
        max_grad_norm=float(g["max_grad_norm"]), huber_delta=float(g["huber_delta"]), actor_lr=float(g["actor_lr"]),
        critic_lr=float(g["critic_lr"]), opti_eps=float(g["opti_eps"]), weight_decay=0, gamma=float(g["gamma"]),
        gae_lambda=float(g["gae_lambda"]), max_step=10, batch_size=6, fused_precision=precision)
    for k, v in over.items():
        setattr(a, k, v)
    return a


def _policy(cuda, g, precision="fp32"):
    from minimarl_b200.algorithms import R_MAPPOPolicy
    args = _args(g, precision)
    space = Box(np.zeros(47), np.ones(47))
    policy = R_MAPPOPolicy(args, space, space, Discrete(5), device=cuda)
    policy.actor.load_state_dict({k[6:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("actor.")})
    policy.critic.load_state_dict({k[7:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("critic.")})
    return args, policy


@pytest.mark.parametrize("precision,tol", [("fp32", 3e-5), ("tf32", 4e-3)])
def test_rollout_step_matches_reference(cuda, precision, tol):
    g = _g()
    args, policy = _policy(cuda, g, precision)
    assert sum(p.numel() for p in policy.actor.parameters()) == 10499
    assert sum(p.numel() for p in policy.critic.parameters()) == 10367
    obs, ha, hc, m = (torch.from_numpy(g[k]).to(cuda) for k in ["step_obs", "step_h_actor", "step_h_critic", "step_masks"])
    with torch.no_grad():
        values, hc1 = policy.critic(obs, hc, m)                          # fused kernel (one launch)
        logits, ha1 = policy.actor._fused_step(obs, ha, m)
        det, ha2 = policy.act(obs, ha, m, deterministic=True)
        v, a, lp, ha3, hc3 = policy.get_actions(obs, obs, ha, hc, m)
    assert np.abs(values.cpu().numpy() - g["step_values"]).max() < tol
    assert np.abs(hc1.cpu().numpy() - g["step_h_critic_out"]).max() < tol
    assert np.abs(logits.cpu().numpy() - g["step_logits"]).max() < tol
    assert np.abs(ha1.cpu().numpy() - g["step_h_actor_out"]).max() < tol
    if precision == "fp32":
        assert np.array_equal(det.cpu().numpy(), g["step_det_actions"])
    assert v.shape == (12, 1) and a.shape == (12, 1) and lp.shape == (12, 1) and ha3.shape == (12, 1, 32)
    want_lp = torch.log_softmax(logits, -1).gather(1, a)
    assert torch.allclose(lp, want_lp, atol=1e-5)


def test_full_mappo_train_matches_reference(cuda):
    from minimarl_b200.algorithms import R_MAPPO
    from minimarl_b200.buffers import SharedReplayBuffer
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = _g()
    args, policy = _policy(cuda, g)
    trainer = R_MAPPO(args, policy, device=cuda)
    space = Box(np.zeros(47), np.ones(47))
    buf = SharedReplayBuffer(args, 2, space, space, Discrete(5), device=cuda)
    for k in ["obs", "share_obs", "rnn_states", "rnn_states_critic", "actions", "masks", "rewards", "value_preds",
              "action_log_probs", "active_masks"]:
        getattr(buf, k).copy_(torch.from_numpy(g["buf_" + k]))
    buf.compute_returns(g["next_value"], trainer.value_normalizer)        # GAE kernel with ValueNorm denormalisation
    assert np.array_equal(buf.returns[:-1].cpu().numpy(), g["buf_returns"][:-1])
    info = trainer.train(buf, perms=list(g["perms"]))
    for k in ["value_loss", "policy_loss", "dist_entropy", "actor_grad_norm", "critic_grad_norm", "ratio", "Sampling_Rewards"]:
        want = float(g["info_" + k])
        got = float(info[k])
        assert abs(got - want) < 2e-4 * max(1.0, abs(want)), (k, got, want)
    for prefix, module in [("actor_after.", policy.actor), ("critic_after.", policy.critic), ("vn_after.", trainer.value_normalizer)]:
        for name, p in module.state_dict().items():
            assert np.allclose(p.cpu().numpy(), g[prefix + name], atol=5e-4, rtol=1e-4), prefix + name


def test_full_mappo_train_fused_tf32_matches_reference(cuda):
    """R_MAPPO.train through the FUSED update (tcgen05 trunk forward / backward, loss kernels, chunks read in place; TF32
    operands) against the reference-recorded run: train_info 5e-4 relative, parameters after the 4 Adam steps 2e-3 abs
    (measured 6e-5 and 3e-4; lr = 5e-4-scale steps turn a sign flip of a ~0 gradient into a 2 lr difference)."""
    from minimarl_b200.algorithms import R_MAPPO
    from minimarl_b200.buffers import SharedReplayBuffer
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        g = _g()
        args, policy = _policy(cuda, g, "tf32")
        trainer = R_MAPPO(args, policy, device=cuda)
        space = Box(np.zeros(47), np.ones(47))
        buf = SharedReplayBuffer(args, 2, space, space, Discrete(5), device=cuda)
        for k in ["obs", "share_obs", "rnn_states", "rnn_states_critic", "actions", "masks", "rewards", "value_preds",
                  "action_log_probs", "active_masks"]:
            getattr(buf, k).copy_(torch.from_numpy(g["buf_" + k]))
        buf.compute_returns(g["next_value"], trainer.value_normalizer)
        calls = []
        fused = trainer._ppo_update_fused
        trainer._ppo_update_fused = lambda *a, **k: (calls.append(1), fused(*a, **k))[1]
        info = trainer.train(buf, perms=list(g["perms"]))
        assert len(calls) == int(g["ppo_epoch"]) * int(g["train_batch_size"]), "the fused update did not run"
        for k in ["value_loss", "policy_loss", "dist_entropy", "actor_grad_norm", "critic_grad_norm", "ratio", "Sampling_Rewards"]:
            want, got = float(g["info_" + k]), float(info[k])
            assert abs(got - want) < 5e-4 * max(1.0, abs(want)), (k, got, want)
        for prefix, module in [("actor_after.", policy.actor), ("critic_after.", policy.critic), ("vn_after.", trainer.value_normalizer)]:
            for name, p in module.state_dict().items():
                assert np.allclose(p.cpu().numpy(), g[prefix + name], atol=2e-3, rtol=1e-3), prefix + name
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32


def test_lockstep_mappo_rollout_and_update(cuda):
    """Runner-shaped loop on the device: collect (fused actor / critic step) -> env kernel writes the buffer slot
    in place -> insert -> compute_returns (GAE kernel) -> train.  Checks shapes, finiteness and that stepping
    N envs at once equals what the buffer holds (obs of step t+1 is the env's obs)."""
    from minimarl_b200 import BatchedCheckers
    from minimarl_b200.algorithms import R_MAPPO, R_MAPPOPolicy
    from minimarl_b200.buffers import SharedReplayBuffer
    g = _g()
    N, T = 96, 16
    args = _args(g, "tf32", max_step=T, batch_size=N, ppo_epoch=2, train_batch_size=1, data_chunk_length=8)
    space = Box(np.zeros(47), np.ones(47))
    policy = R_MAPPOPolicy(args, space, space, Discrete(5), device=cuda)
    trainer = R_MAPPO(args, policy, device=cuda)
    buf = SharedReplayBuffer(args, 2, space, space, Discrete(5), device=cuda)
    env = BatchedCheckers(N, max_steps=12, auto_reset=True, device=cuda)
    env.reset(out=buf.obs[0])
    buf.share_obs[0].copy_(buf.obs[0])
    dones = torch.zeros((N, 2), dtype=torch.uint8, device=cuda)
    trainer.prep_rollout()
    with torch.no_grad():
        for t in range(T):
            v, a, lp, ha, hc = policy.get_actions(buf.obs[t].reshape(N * 2, 47), buf.share_obs[t].reshape(N * 2, 47),
                                                  buf.rnn_states[t].reshape(N * 2, 1, 32),
                                                  buf.rnn_states_critic[t].reshape(N * 2, 1, 32), buf.masks[t].reshape(N * 2, 1))
            env.step_into(a.to(torch.int32).reshape(N, 2), buf.obs[t + 1], buf.rewards[t].view(N, 2), dones)
            done_env = dones.bool().all(1)
            masks = (~done_env).float().view(N, 1, 1).expand(N, 2, 1)
            ha = ha.reshape(N, 2, 1, 32) * masks.unsqueeze(-1)
            hc = hc.reshape(N, 2, 1, 32) * masks.unsqueeze(-1)
            buf.insert(buf.obs[t + 1], buf.obs[t + 1], ha, hc, a.reshape(N, 2, 1).float(), lp.reshape(N, 2, 1),
                       v.reshape(N, 2, 1), buf.rewards[t], masks)
        next_v = policy.get_values(buf.share_obs[-1].reshape(N * 2, 47), buf.rnn_states_critic[-1].reshape(N * 2, 1, 32),
                                   buf.masks[-1].reshape(N * 2, 1))
    buf.compute_returns(next_v.reshape(N, 2, 1), trainer.value_normalizer)
    trainer.prep_training()
    info = trainer.train(buf)
    assert all(torch.isfinite(torch.as_tensor(v)).all() for v in info.values())
    assert (buf.masks[1:] == 0).any()            # some episodes ended (max_steps = 12) and were auto-reset
    buf.after_update()


def test_graphed_ppo_epochs_equal_the_eager_loop(cuda):
    """R_MAPPO.train with the epoch replayed as a CUDA graph (the default on one GPU) against the same loop issued launch by
    launch (train_graph=False): same device permutations, same kernels, so parameters, ValueNorm state and the reported
    statistics agree after two updates (the second one is all replays) up to the summation order of the dynamically scheduled
    weight-gradient tiles (2e-5 on parameters that moved by ~3e-3, 1e-3 relative on the statistics) -- including a
    learning-rate change between them, which the graph follows through FlatAdam's device-side lr."""
    from minimarl_b200.algorithms import R_MAPPO, R_MAPPOPolicy
    from minimarl_b200.buffers import SharedReplayBuffer
    g = _g()
    N, T = 64, 16
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
    try:
        results = []
        for graph in (False, True):
            args = _args(g, "tf32", max_step=T, batch_size=N, ppo_epoch=3, train_batch_size=1, data_chunk_length=8, train_graph=graph)
            space = Box(np.zeros(47), np.ones(47))
            torch.manual_seed(11)
            policy = R_MAPPOPolicy(args, space, space, Discrete(5), device=cuda)
            trainer = R_MAPPO(args, policy, device=cuda)
            buf = SharedReplayBuffer(args, 2, space, space, Discrete(5), device=cuda)
            infos = []
            for it in range(2):
                gen = torch.Generator(device=cuda).manual_seed(100 + it)
                r = lambda x: torch.rand(x.shape, device=cuda, generator=gen)
                buf.obs.copy_(r(buf.obs)); buf.share_obs.copy_(buf.obs)
                buf.rnn_states.copy_(r(buf.rnn_states) - 0.5); buf.rnn_states_critic.copy_(r(buf.rnn_states_critic) - 0.5)
                buf.actions.copy_((r(buf.actions) * 5).floor().clamp(max=4)); buf.action_log_probs.copy_(-1.6 + 0.1 * r(buf.action_log_probs))
                buf.value_preds.copy_(r(buf.value_preds)); buf.rewards.copy_(r(buf.rewards) - 0.3)
                buf.masks.copy_((r(buf.masks) > 0.1).float()); buf.active_masks.fill_(1.0)
                buf.compute_returns(buf.value_preds[-1].clone(), trainer.value_normalizer)
                policy.lr_decay(it, 4)
                trainer.prep_training()
                info = trainer.train(buf)
                infos.append({k: float(v) for k, v in info.items()})
            assert (trainer._graph_state is not None) == graph
            params = [p.detach().clone() for p in list(policy.actor.parameters()) + list(policy.critic.parameters())]
            results.append((infos, params, trainer.value_normalizer.running_mean.detach().clone()))
        (i0, p0, v0), (i1, p1, v1) = results
        for a, b in zip(i0, i1):
            assert all(abs(a[k] - b[k]) <= 1e-3 * max(abs(a[k]), 1e-3) for k in a), (a, b)
        assert all((a - b).abs().max().item() < 2e-5 for a, b in zip(p0, p1)) and torch.allclose(v0, v1, rtol=1e-5)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
