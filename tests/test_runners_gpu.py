"""GPU: the runner-level drop-in surface (offpolicy RecRunner / MAPPO MAGYM_Runner / vdn main loop) drives the
device pipeline end to end for a few iterations; counters follow the reference's bookkeeping."""
import argparse

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_offpolicy_runner_qmix_and_vdn(cuda):
    from minimarl_b200 import BatchedCheckers
    from minimarl_b200.runners import OffPolicyRunner
    from test_qmix_gpu import _args
    for algo in ("qmix", "vdn"):
        N, T = 64, 25
        args = _args(algorithm_name=algo, episode_length=T, batch_size=32, buffer_size=512, use_eval=True, use_soft_update=True,
                     num_env_steps=10 * N * T, train_interval_episode=1, hard_update_interval_episode=200, log_interval=N * T,
                     eval_interval=4 * N * T, per_alpha=0.6, per_beta_start=0.4, use_reward_normalization=False,
                     num_random_episodes=5, fused_precision="tf32")
        env = BatchedCheckers(N, max_steps=T, device=cuda)
        runner = OffPolicyRunner({"args": args, "train_env": env, "eval_env": env, "num_agents": 2, "device": cuda})
        assert len(runner.buffer) == N                      # warm-up: max(batch_size, 5) // N + 1 batches of N episodes
        steps = 0
        for _ in range(6):
            steps = runner.run()
        # warm-up steps count towards total_env_steps, as in the reference (magym_runner.py:110-111)
        assert steps == 7 * N * T and runner.num_episodes_collected >= 6 * N and runner.total_train_steps == 6
        assert np.isfinite(runner.last_log["loss"]) and np.isfinite(runner.last_log["average_episode_rewards"])
        assert torch.isfinite(runner.eval())


def test_mappo_runner(cuda):
    from minimarl_b200 import BatchedCheckers
    from minimarl_b200.runners import MappoRunner
    from test_mappo_gpu import _args, _g
    N, T = 64, 16
    args = _args(_g(), "tf32", max_step=T, batch_size=N, ppo_epoch=2, train_batch_size=1, data_chunk_length=8,
                 max_episodes=3, use_eval=True, eval_interval=2, use_linear_lr_decay=False)
    env = BatchedCheckers(N, max_steps=100, auto_reset=True, device=cuda)
    eval_env = BatchedCheckers(16, max_steps=30, device=cuda)
    runner = MappoRunner({"args": args, "train_env": env, "eval_env": eval_env, "num_agents": 2, "device": cuda})
    infos = runner.run()
    assert len(infos) == 3
    for k in ["value_loss", "policy_loss", "dist_entropy", "actor_grad_norm", "critic_grad_norm", "ratio"]:
        assert np.isfinite(float(infos[-1][k])), k
    assert "Test_Rewards" in infos[0]


def test_vdn_main_loop(cuda):
    from minimarl_b200.runners import run_vdn
    args = argparse.Namespace(use_recurrent=True, buffer_limit=256, batch_size=16, update_iter=2, chunk_size=10, gamma=0.99,
                              eps=1e-6, alpha=0.4, beta=0.4, step_weight=0.99, use_step_weight=True, update_alpha_beta=True,
                              max_episodes=100, grad_clip_norm=5, lr=1e-3, seed=42, max_step=20, step_cost=-0.01,
                              max_epsilon=0.8, min_epsilon=0.05, epsilon_anneal_episode=50, update_target_interval=2)
    history, net = run_vdn(args, n_envs=16, episodes=3, device=cuda)
    assert len(history) == 3 and np.isfinite(history[-1]["train_loss"]) and history[-1]["alpha"] > 0.4


def test_qmix_main_loop_on_lockstep_switch(cuda):
    """qmix/main.py loop on lockstep Switch2 envs (the qmix default env): rollout -> batched chunk inserts (more chunks
    than slots, so the in-batch eviction runs) -> Train_dqn with the recurrent mixer."""
    from minimarl_b200.runners import run_qmix
    args = argparse.Namespace(use_recurrent=True, buffer_limit=100, batch_size=32, update_iter=2, chunk_size=10, gamma=0.99,
                              eps=1e-6, alpha=0.4, beta=0.4, use_step_weight=False, update_alpha_beta=True,
                              max_episodes=100, grad_clip_norm=5, lr=1e-3, seed=42, max_step=20, step_cost=-0.1,
                              max_epsilon=0.9, min_epsilon=0.05, epsilon_anneal_episode=50, update_target_interval=2)
    timing = {}
    history, net = run_qmix(args, n_envs=64, episodes=3, device=cuda, timing=timing)
    assert len(history) == 3 and np.isfinite(history[-1]["train_loss"]) and history[-1]["beta"] > 0.4
    assert timing["chunks_in_buffer"] == 100 and timing["env_steps"] == 3 * 20 * 64


def test_graphed_training_iteration_matches_eager(cuda):
    """One Train_dqn iteration captured as a CUDA graph (PER draw with device-side alpha / beta, forward, backward, clipping,
    capturable Adam, priority write-back) replayed N times leaves the same weights, tree and schedule as N eager iterations
    started from the same state with the same random stream."""
    import copy
    from minimarl_b200 import BatchedSwitch
    from minimarl_b200.algorithms import Mix_Net, Prioritized_Experience_Replay, Q_Net, Train_dqn
    args = argparse.Namespace(use_recurrent=True, buffer_limit=256, batch_size=64, update_iter=4, chunk_size=10, gamma=0.99,
                              eps=1e-6, alpha=0.6, beta=0.4, use_step_weight=False, update_alpha_beta=True, max_episodes=50,
                              grad_clip_norm=5, lr=1e-3, seed=1)
    env = BatchedSwitch(8, n_agents=2, max_steps=50, device=cuda)
    torch.manual_seed(5)
    nets = [Q_Net(env.observation_space, env.action_space, args).to(cuda), Mix_Net(env.observation_space, args).to(cuda),
            Q_Net(env.observation_space, env.action_space, args).to(cuda), Mix_Net(env.observation_space, args).to(cuda)]
    g = torch.Generator(device=cuda).manual_seed(3)
    K, L = 256, 10
    chunks = [torch.rand((K, L, 2, 3), device=cuda, generator=g), torch.randint(0, 5, (K, L, 2), device=cuda, generator=g).float(),
              torch.randn((K, L, 2), device=cuda, generator=g), torch.rand((K, L, 2, 3), device=cuda, generator=g),
              (torch.rand((K, L), device=cuda, generator=g) < 0.05).float()]
    td = torch.rand(K, device=cuda, generator=g) * 3

    def run(graph):
        a = copy.deepcopy(args)
        a.train_graph = graph
        q, mix, tq, tmix = (copy.deepcopy(n) for n in nets)
        per = Prioritized_Experience_Replay(a, n_agents=2, obs_dim=3, device=cuda)
        per.collect_samples_batch(chunks, td, True)
        per.enable_graph_mode()                      # both runs draw from the default generator with device-side exponents
        opt = torch.optim.Adam(list(q.parameters()) + list(mix.parameters()), lr=a.lr, capturable=True)
        trainer = Train_dqn(a, cuda)
        torch.manual_seed(11)
        for _ in range(2):
            loss = trainer.train(per, q, mix, tq, tmix, opt, 0.5)
        return loss, q, mix, per

    l0, q0, m0, p0 = run(False)
    l1, q1, m1, p1 = run(True)
    assert p0.alpha == p1.alpha and p0.beta == p1.beta
    assert torch.allclose(l0, l1, rtol=1e-4), (l0, l1)
    for a_, b_ in zip(list(q0.parameters()) + list(m0.parameters()), list(q1.parameters()) + list(m1.parameters())):
        assert torch.allclose(a_, b_, rtol=1e-3, atol=1e-4)
    assert torch.allclose(p0.sum_tree.priority_tree, p1.sum_tree.priority_tree, rtol=1e-4, atol=1e-9)
