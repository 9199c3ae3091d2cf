"""GPU parity: the fused MAPPO rollout kernel (mm_mappo_rollout_tc) against (a) the env oracle replaying the
actions it sampled -- obs / rewards / masks bit-exact -- and (b) CPU fp32 copies of R_Actor / R_Critic evaluated on
the buffer it wrote: value_preds, action_log_probs and the next hidden states within the tensor-core (tf32)
tolerance 1e-2 abs (observed ~2e-3), plus the sampling law (inverse CDF on softmax) against the exact probabilities."""
import copy

import numpy as np
import pytest
import torch

from oracle.magym_c import CheckersBatchRef

pytestmark = pytest.mark.gpu

TOL = 1e-2


def _setup(cuda, N, T, max_steps=100, seed=3):
    from minimarl_b200 import BatchedCheckers
    from minimarl_b200.algorithms import R_MAPPOPolicy
    from minimarl_b200.buffers import SharedReplayBuffer
    from minimarl_b200.envs.spaces import Box, Discrete
    from test_mappo_gpu import _args, _g
    torch.manual_seed(seed)
    args = _args(_g(), "tf32", max_step=T, batch_size=N, hidden_size=32)
    space = Box(np.zeros(47), np.ones(47))
    policy = R_MAPPOPolicy(args, space, space, Discrete(5), device=cuda)
    with torch.no_grad():                         # make the nets non-trivial: default head gain 0.01 gives ~uniform logits
        policy.actor.act.action_out.linear.weight.mul_(60.0)
        for p in list(policy.actor.parameters()) + list(policy.critic.parameters()):
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    env = BatchedCheckers(N, max_steps=max_steps, auto_reset=True, device=cuda)
    buf = SharedReplayBuffer(args, 2, space, space, Discrete(5), device=cuda)
    return args, policy, env, buf


def test_fused_rollout_replays_through_env_oracle_and_cpu_networks(cuda):
    from minimarl_b200.rollout import mappo_rollout, mappo_rollout_fusable, pack_mappo_rollout
    N, T, MS = 301, 40, 13                          # ragged tile (602 rows), several auto-resets inside the rollout
    args, policy, env, buf = _setup(cuda, N, T, MS)
    assert mappo_rollout_fusable(policy, env)
    params = pack_mappo_rollout(policy.actor, policy.critic)
    mappo_rollout(params, env, buf, T, seed=11, step_offset=5, reset=True)
    torch.cuda.synchronize()
    g = {k: getattr(buf, k).cpu() for k in ["obs", "share_obs", "rnn_states", "rnn_states_critic", "actions",
                                           "action_log_probs", "value_preds", "rewards", "masks"]}
    assert torch.equal(g["obs"], g["share_obs"])
    # (a) env: replay the sampled actions through the oracle
    ref = CheckersBatchRef(N, False, MS, -0.01, auto_reset=True)
    o = ref.reset()
    assert np.array_equal(g["obs"][0].numpy(), o)
    acts = g["actions"].numpy().astype(np.int32)
    assert acts.min() >= 0 and acts.max() <= 4 and len(np.unique(acts)) == 5
    n_done = 0
    for t in range(T):
        o, r, d = ref.step(acts[t, :, :, 0])
        assert np.array_equal(g["obs"][t + 1].numpy(), o), t
        assert np.array_equal(g["rewards"][t, :, :, 0].numpy(), r), t
        keep = 1.0 - d.all(axis=1).astype(np.float32)
        assert np.array_equal(g["masks"][t + 1, :, :, 0].numpy(), np.stack([keep, keep], 1)), t
        n_done += int(d.all(axis=1).sum())
    assert n_done >= 3 * N                          # T // MS resets per env
    # state handed back to the env object continues the same trajectories
    a_next = np.random.default_rng(0).integers(0, 5, size=(N, 2), dtype=np.int32)
    obs2, rew2, done2 = env.step(torch.from_numpy(a_next).to(cuda))
    o, r, d = ref.step(a_next)
    assert np.array_equal(obs2.cpu().numpy(), o) and np.array_equal(rew2.cpu().numpy(), r)
    # (b) networks: CPU fp32 modules on the recorded inputs
    actor, critic = copy.deepcopy(policy.actor).cpu(), copy.deepcopy(policy.critic).cpu()
    actor._fusable = critic._fusable = False
    actor.tpdv = critic.tpdv = dict(dtype=torch.float32, device="cpu")
    R = N * 2
    worst = 0.0
    with torch.no_grad():
        for t in range(T):
            obs, m = g["obs"][t].reshape(R, 47), g["masks"][t].reshape(R, 1)
            ha, hc = g["rnn_states"][t].reshape(R, 1, 32), g["rnn_states_critic"][t].reshape(R, 1, 32)
            lp, _ = actor.evaluate_actions(obs, ha, g["actions"][t].reshape(R, 1), m)
            feats, ha1 = actor.rnn(actor.base(obs), ha, m)
            v, hc1 = critic(obs, hc, m)
            keep = g["masks"][t + 1].reshape(R, 1, 1)
            for got, want in ((g["action_log_probs"][t].reshape(R, 1), lp), (g["value_preds"][t].reshape(R, 1), v),
                              (g["rnn_states"][t + 1].reshape(R, 1, 32), ha1 * keep),
                              (g["rnn_states_critic"][t + 1].reshape(R, 1, 32), hc1 * keep)):
                worst = max(worst, float((got - want).abs().max()))
    assert worst < TOL, worst


def test_fused_rollout_sampling_law_and_determinism(cuda):
    from minimarl_b200.rollout import mappo_rollout, pack_mappo_rollout
    N, T = 16384, 2
    args, policy, env, buf = _setup(cuda, N, T)
    params = pack_mappo_rollout(policy.actor, policy.critic)
    mappo_rollout(params, env, buf, T, seed=1, reset=True)
    a1, lp1 = buf.actions.clone(), buf.action_log_probs.clone()
    # step 0: every env shows the same two observations, so the action counts are multinomial with known probabilities
    actor = copy.deepcopy(policy.actor).cpu()
    actor._fusable = False
    actor.tpdv = dict(dtype=torch.float32, device="cpu")
    with torch.no_grad():
        obs = buf.obs[0, 0].cpu()
        feats, _ = actor.rnn(actor.base(obs), torch.zeros(2, 1, 32), torch.ones(2, 1))
        probs = torch.softmax(actor.act.action_out.linear(feats), -1).numpy()
    assert probs.max() < 0.95 and probs.min() > 1e-3          # a non-degenerate law to test against
    for ag in range(2):
        counts = np.bincount(a1[0, :, ag, 0].cpu().numpy().astype(int), minlength=5)
        sigma = np.sqrt(N * probs[ag] * (1 - probs[ag]))
        assert np.all(np.abs(counts - N * probs[ag]) < 5 * sigma + 0.01 * N), (counts, N * probs[ag])
        # the recorded log-prob is the log of that law
        lp = lp1[0, :, ag, 0].cpu().numpy()
        assert np.abs(lp - np.log(probs[ag])[a1[0, :, ag, 0].cpu().numpy().astype(int)]).max() < TOL
    # same seed / offset -> identical; another offset -> another draw
    buf.rnn_states.zero_(); buf.rnn_states_critic.zero_(); buf.masks.fill_(1.0)
    mappo_rollout(params, env, buf, T, seed=1, reset=True)
    assert torch.equal(buf.actions, a1) and torch.equal(buf.action_log_probs, lp1)
    mappo_rollout(params, env, buf, T, seed=1, step_offset=100, reset=True)
    assert not torch.equal(buf.actions, a1)


def test_runner_fused_and_stepwise_paths(cuda):
    from minimarl_b200 import BatchedCheckers
    from minimarl_b200.runners import MappoRunner
    from test_mappo_gpu import _args, _g
    N, T = 96, 24
    for fused in (True, False):
        args = _args(_g(), "tf32", max_step=T, batch_size=N, hidden_size=32, ppo_epoch=2, train_batch_size=1,
                     data_chunk_length=8, max_episodes=3, use_eval=False, eval_interval=2, use_linear_lr_decay=False,
                     fused_rollout=fused)
        env = BatchedCheckers(N, max_steps=20, auto_reset=True, device=cuda)
        runner = MappoRunner({"args": args, "train_env": env, "eval_env": env, "num_agents": 2, "device": cuda})
        assert runner.fused_rollout == fused
        infos = runner.run()
        assert len(infos) == 3 and all(np.isfinite(float(v)) for v in infos[-1].values())
        b = runner.buffer
        # buffer invariants either path must leave behind: masks are 0 exactly where an episode ended, hidden states too
        ended = b.masks[1:] == 0
        assert ended.any() and bool((b.rnn_states[1:][ended.squeeze(-1)] == 0).all())
        assert bool(((b.actions >= 0) & (b.actions <= 4)).all()) and bool((b.action_log_probs <= 0).all())
