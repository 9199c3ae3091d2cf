"""MAPPO on 4 096 lockstep Checkers-v0 boards through the fused rollout / GAE / PPO-update kernels: the greedy team return
(what mappo's `eval` logs as Test_Rewards) has to climb from the random-policy level (about -2: the step costs) towards the
apples.  Functional check of the whole mappo/ hot path, not a benchmark."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minimarl_b200 import BatchedCheckers
from minimarl_b200.runners import MappoRunner

dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
N, T, ITERS = 4096, 104, int(sys.argv[1]) if len(sys.argv) > 1 else 60
args = argparse.Namespace(
    hidden_size=32, layer_N=1, recurrent_N=1, use_feature_normalization=True, use_ReLU=True, use_orthogonal=True,
    gain=0.01, use_conv1d=False, use_recurrent_policy=True, use_naive_recurrent_policy=False,
    use_policy_active_masks=True, use_value_active_masks=True, use_popart=False, use_valuenorm=True,
    use_huber_loss=True, use_max_grad_norm=True, use_clipped_value_loss=True, use_gae=True, ppo_epoch=10,
    train_batch_size=1, data_chunk_length=8, clip_param=0.2, entropy_coef=0.01, value_loss_coef=1.0,
    max_grad_norm=10.0, huber_delta=10.0, actor_lr=5e-4, critic_lr=5e-4, opti_eps=1e-5, weight_decay=0, gamma=0.99,
    gae_lambda=0.95, max_step=T, batch_size=N, fused_precision="tf32", max_episodes=ITERS, use_eval=False,
    eval_interval=10, use_linear_lr_decay=False, seed=1)
env = BatchedCheckers(N, max_steps=100, auto_reset=True, device=dev)
eval_env = BatchedCheckers(512, max_steps=100, device=dev)
runner = MappoRunner({"args": args, "train_env": env, "eval_env": eval_env, "num_agents": 2, "device": dev})
assert runner.fused_rollout
runner.warmup()
t0 = time.time()
first = last = None
for it in range(ITERS):
    infos = runner.run_episode(it)
    if it % 10 == 0 or it == ITERS - 1:
        ev = float(runner.eval())
        first = ev if first is None else first
        last = ev
        print(f"iter {it:4d} env_steps {(it + 1) * N * T:9d} greedy_eval_return {ev:8.3f} "
              f"mean_step_reward {float(runner.buffer.rewards.mean()):8.4f} value_loss {float(infos['value_loss']):.4f} "
              f"entropy {float(infos['dist_entropy']):.3f} wall {time.time() - t0:6.1f}s", flush=True)
print("improved" if last > first + 1.0 else "NOT improved", first, "->", last)
