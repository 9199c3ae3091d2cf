"""Functional check of the runners end to end (not parity): the policies the ROLLOUT kernels execute must follow the learner's
updates.  Guards the coupling the parity tests do not see -- packed-parameter caches keyed on ``p._version``, graph replays,
raw-pointer optimiser steps -- by asserting that short runs of the reference's loops actually learn Checkers-v0:
greedy return of the random-initialised policies is about -2 (the step costs)."""
import argparse

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_qmix_runner_learns_checkers(cuda):
    from minimarl_b200 import BatchedCheckers
    from minimarl_b200.runners import OffPolicyRunner
    from test_qmix_gpu import _args
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
    try:
        torch.manual_seed(0)
        N, T = 512, 100
        args = _args(algorithm_name="qmix", episode_length=T, batch_size=256, buffer_size=8192, use_eval=False, use_soft_update=True,
                     num_env_steps=400 * N * T, train_interval_episode=1, hard_update_interval_episode=200, log_interval=10 ** 12,
                     eval_interval=10 ** 12, per_alpha=0.6, per_beta_start=0.4, use_reward_normalization=False,
                     num_random_episodes=5, fused_precision="tf32", epsilon_anneal_time=100 * N * T, lr=5e-4)
        env = BatchedCheckers(N, max_steps=T, device=cuda)
        runner = OffPolicyRunner({"args": args, "train_env": env, "eval_env": env, "num_agents": 2, "device": cuda})
        first = float(runner.eval())
        best = first
        for it in range(200):
            runner.run()
            if it % 40 == 39:
                best = max(best, float(runner.eval()))
        assert first < 0 and best > 3.0, (first, best)           # examples/learn_qmix.py reaches 8.0 by ~100 updates
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32


def test_mappo_runner_learns_checkers(cuda):
    from minimarl_b200 import BatchedCheckers
    from minimarl_b200.runners import MappoRunner
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
    try:
        torch.manual_seed(1)
        N, T, iters = 2048, 104, 40
        args = argparse.Namespace(
            hidden_size=32, layer_N=1, recurrent_N=1, use_feature_normalization=True, use_ReLU=True, use_orthogonal=True,
            gain=0.01, use_conv1d=False, use_recurrent_policy=True, use_naive_recurrent_policy=False,
            use_policy_active_masks=True, use_value_active_masks=True, use_popart=False, use_valuenorm=True,
            use_huber_loss=True, use_max_grad_norm=True, use_clipped_value_loss=True, use_gae=True, ppo_epoch=10,
            train_batch_size=1, data_chunk_length=8, clip_param=0.2, entropy_coef=0.01, value_loss_coef=1.0,
            max_grad_norm=10.0, huber_delta=10.0, actor_lr=5e-4, critic_lr=5e-4, opti_eps=1e-5, weight_decay=0, gamma=0.99,
            gae_lambda=0.95, max_step=T, batch_size=N, fused_precision="tf32", max_episodes=iters, use_eval=False,
            eval_interval=10, use_linear_lr_decay=False, seed=1)
        env = BatchedCheckers(N, max_steps=100, auto_reset=True, device=cuda)
        eval_env = BatchedCheckers(256, max_steps=100, device=cuda)
        runner = MappoRunner({"args": args, "train_env": env, "eval_env": eval_env, "num_agents": 2, "device": cuda})
        runner.warmup()
        first = float(runner.eval())
        for it in range(iters):
            runner.run_episode(it)
        last = float(runner.eval())
        assert runner.trainer._graph_state is not None                  # the graphed epochs were the path under test
        assert last > first + 5.0, (first, last)                 # examples/learn_mappo.py: -2 -> ~40 in 60 iterations
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
