"""GPU parity: fused policy + env rollout kernel (mm_qmix_rollout_tc).

* eps = 0: identical, step for step, to the un-fused path (tensor-core agent forward with T1 = 1, arg-max,
  BatchedCheckers.step) -- same arithmetic, so obs / rewards / dones / actions / hidden state are bit-equal.
* any eps: the recorded actions replayed through the CPU oracle give bit-equal obs / dones and the team reward
  float32(double(r0) + double(r1)), with stepping frozen after the episode ends (collect_rollout semantics).
"""
import argparse

import numpy as np
import pytest
import torch

from oracle.magym_c import CheckersBatchRef

pytestmark = pytest.mark.gpu


def _net(cuda, hidden=64, seed=0):
    from minimarl_b200.nn import AgentQFunction
    torch.manual_seed(seed)
    args = argparse.Namespace(hidden_size=hidden, use_rnn_layer=True, use_orthogonal=True, gain=0.01, layer_N=1,
                              recurrent_N=1, use_feature_normalization=True, use_ReLU=True, use_conv1d=False,
                              fused_precision="tf32")
    net = AgentQFunction(args, 47, 5, cuda)
    with torch.no_grad():
        net.q.action_out.weight.mul_(100)          # spread the Q-values so that arg-max is not a near tie
        net.q.action_out.bias.normal_(0, 0.05)
    return net


def _replay_through_oracle(env, out, T, N, max_steps, step_cost=-0.01):
    ref = CheckersBatchRef(N, False, max_steps, step_cost)
    obs_ref = ref.reset()
    acts = out["act_idx"].cpu().numpy()
    assert np.array_equal(out["obs"][0].cpu().numpy(), obs_ref)
    finished = np.zeros(N, dtype=bool)
    table = env.reward_table
    for t in range(T):
        o, r, d = ref.step(acts[t])
        # frozen envs keep their previous obs, get zero reward and stay done
        obs_ref = np.where(finished[:, None, None], obs_ref, o)
        team = np.array([np.float32(table[float(a)] + table[float(b)]) for a, b in r], dtype=np.float32)
        team = np.where(finished, np.float32(0), team)
        done = np.where(finished[:, None], True, d)
        assert np.array_equal(out["obs"][t + 1].cpu().numpy(), obs_ref), f"obs differ at step {t}"
        assert np.array_equal(out["rewards"][t].cpu().numpy(), np.stack([team, team], 1)), f"rewards differ at step {t}"
        assert np.array_equal(out["dones"][t].cpu().numpy() > 0, done), f"dones differ at step {t}"
        assert np.array_equal(out["dones_env"][t].cpu().numpy() > 0, done.all(1))
        # the C oracle keeps stepping finished envs; un-do that by never looking at them again
        finished |= d.all(1)


@pytest.mark.parametrize("hidden,n_envs", [(64, 200), (32, 64)])
def test_greedy_fused_rollout_equals_stepwise_path(cuda, hidden, n_envs):
    from minimarl_b200 import BatchedCheckers
    from minimarl_b200.nn.agent_q import agentq_forward
    from minimarl_b200.rollout import qmix_rollout
    T, max_steps = 30, 25
    net = _net(cuda, hidden)
    env = BatchedCheckers(n_envs, max_steps=max_steps, device=cuda)
    q_out = torch.empty((T, n_envs * 2, 5), device=cuda)
    out = qmix_rollout(net, env, T, eps=0.0, q_out=q_out, freeze_after_done=False, common_reward=False)
    # un-fused: same tensor-core kernel with T1 = 1, arg-max, env.step
    env2 = BatchedCheckers(n_envs, max_steps=max_steps, device=cuda)
    obs = env2.reset()
    h = torch.zeros((n_envs * 2, hidden), device=cuda)
    for t in range(T):
        q, h, _ = agentq_forward(net.packed_tc(), obs.reshape(n_envs * 2, 47), h, 47, hidden, 5, tensor_cores=True)
        assert torch.equal(q, q_out[t]), f"q differs at step {t}"
        act = q.argmax(-1).to(torch.int32).reshape(n_envs, 2)
        assert torch.equal(act, out["act_idx"][t])
        obs, rew, done = env2.step(act)
        assert torch.equal(obs, out["obs"][t + 1]) and torch.equal(rew, out["rewards"][t])
        assert torch.equal(done.float(), out["dones"][t])
    assert torch.equal(h, out["h_final"]) and torch.equal(env.state, env2.state)


@pytest.mark.parametrize("eps,per_env", [(0.0, False), (0.3, False), (1.0, False), (0.5, True)])
def test_fused_rollout_replays_through_oracle(cuda, eps, per_env):
    from minimarl_b200 import BatchedCheckers
    from minimarl_b200.rollout import qmix_rollout
    n_envs, T, max_steps = 333, 60, 40            # odd env count: ragged last tile, unaligned obs slabs
    net = _net(cuda, 64, seed=3)
    env = BatchedCheckers(n_envs, max_steps=max_steps, device=cuda)
    q_out = torch.empty((T, n_envs * 2, 5), device=cuda)
    out = qmix_rollout(net, env, T, eps=eps, seed=11, q_out=q_out, eps_per_env=per_env)
    acts = out["act_idx"].cpu().numpy()
    assert acts.min() >= 0 and acts.max() <= 4
    _replay_through_oracle(env, out, T, n_envs, max_steps)
    greedy = q_out.argmax(-1).reshape(T, n_envs, 2).cpu().numpy()
    non_greedy = (acts != greedy).mean()
    assert abs(non_greedy - eps * 0.8) < 0.03, non_greedy          # a random action hits the greedy one 1 time in 5
    if eps == 1.0:
        counts = np.bincount(acts.reshape(-1), minlength=5) / acts.size
        assert np.all(np.abs(counts - 0.2) < 0.02)
    if per_env:                                                     # vdn: both agents of an env explore together
        explored = acts != greedy
        both_or_none = (explored[..., 0] == explored[..., 1]).mean()
        assert both_or_none > 0.8
    # a second rollout continues from the stored env state and draws fresh numbers
    out2 = qmix_rollout(net, env, 5, eps=eps, seed=11, step_offset=T, reset=False, obs=torch.cat([out["obs"][-1:], torch.zeros((5, n_envs, 2, 47), device=cuda)]))
    assert out2["obs"].shape[0] == 6


def test_episode_collector_writes_replay_slots_in_place(cuda):
    from minimarl_b200 import BatchedCheckers
    from minimarl_b200.algorithms import QMixPolicy
    from minimarl_b200.buffers import PrioritizedRecReplayBuffer
    from minimarl_b200.envs.spaces import Box, Discrete
    from minimarl_b200.rollout import EpisodeCollector
    from test_qmix_gpu import _args
    T, N, S, A, D, NA = 20, 64, 256, 2, 47, 5
    args = _args(episode_length=T, batch_size=32, fused_precision="tf32")
    info = {"policy_0": {"obs_space": Box(np.zeros(D), np.ones(D)), "share_obs_space": Box(np.zeros(A * D), np.ones(A * D)),
                         "act_space": Discrete(NA)}}
    pconf = dict(info["policy_0"], cent_obs_dim=A * D, cent_act_dim=A * NA)
    policy = QMixPolicy(args, {"args": args, "device": cuda, "num_agents": A}, pconf)
    buf = PrioritizedRecReplayBuffer(0.6, info, {"policy_0": [0, 1]}, S, T, True, False, device=cuda)
    env = BatchedCheckers(N, max_steps=T, device=cuda)
    col = EpisodeCollector(env, policy, buf, T, seed=5)
    for k in range(3):
        info_ = col.collect_rollout(explore=True, warmup=(k == 0))
    pb = buf.policy_buffers["policy_0"]
    assert len(buf) == 3 * N and pb.current_i == 3 * N and col.total_env_steps == 3 * N * T
    assert torch.isfinite(info_["average_episode_rewards"])
    assert torch.equal(pb.share_obs[:, :3 * N], pb.obs[:, :3 * N].reshape(T + 1, 3 * N, A * D))
    assert torch.all(pb.acts[:, :3 * N].sum(-1) == 1)                     # proper one-hot actions
    assert torch.all(pb.dones_env[-1, :3 * N] == 1)                      # horizon reached
    assert torch.all(pb.obs[:, 3 * N:] == 0) and torch.all(pb.dones[:, 3 * N:] == 1)   # untouched slots keep defaults
    assert torch.equal(pb.rewards[:, :3 * N, 0], pb.rewards[:, :3 * N, 1])  # team reward: every agent gets the sum
    # sampled batch feeds the trainer end to end
    from minimarl_b200.algorithms import QMix
    trainer = QMix(args, A, 32, {"policy_0": policy}, lambda a: "policy_0", device=cuda, episode_length=T)
    sample = buf.sample(32, beta=0.4, p_id="policy_0")
    tinfo, prio, idx = trainer.train_policy_on_batch(sample)
    buf.update_priorities(idx, prio, "policy_0")
    assert torch.isfinite(tinfo["loss"])
    # faithful mode reproduces the reference's storage quirks (Q11)
    col_f = EpisodeCollector(env, policy, buf, T, seed=6, faithful=True)
    col_f.collect_rollout(explore=True)
    sl = slice(3 * N, 4 * N)
    assert torch.all(pb.obs[T, sl] == 0)
    assert torch.all(pb.acts[:, sl].min(-1)[0] == pb.acts[:, sl].max(-1)[0])
