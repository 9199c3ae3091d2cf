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
