"""Driver for the r1d captures: one MAPPO iteration at BASELINE configs[2] size (16384 envs x 128 steps, hidden 32) with
ppo_epoch 2 so that the ncu launch list stays short.  Run plain first, then under ncu (see profiles/README.md)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minimarl_b200 import BatchedCheckers  # noqa: E402
from minimarl_b200.runners import MappoRunner  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
dev = torch.device("cuda:0")
N, T = 16384, 128
args = argparse.Namespace(
    hidden_size=32, layer_N=1, recurrent_N=1, use_feature_normalization=True, use_ReLU=True, use_orthogonal=True, gain=0.01,
    use_conv1d=False, use_recurrent_policy=True, use_naive_recurrent_policy=False, use_policy_active_masks=True,
    use_value_active_masks=True, use_popart=False, use_valuenorm=True, use_huber_loss=True, use_max_grad_norm=True,
    use_clipped_value_loss=True, use_gae=True, ppo_epoch=2, train_batch_size=1, data_chunk_length=8, clip_param=0.2,
    entropy_coef=0.01, value_loss_coef=1.0, max_grad_norm=10.0, huber_delta=10.0, actor_lr=5e-4, critic_lr=5e-4,
    opti_eps=1e-5, weight_decay=0, gamma=0.99, gae_lambda=0.95, max_step=T, batch_size=N, fused_precision="tf32",
    max_episodes=10, use_eval=False, eval_interval=10, use_linear_lr_decay=False, seed=1)
env = BatchedCheckers(N, max_steps=100, auto_reset=True, device=dev)
r = MappoRunner({"args": args, "train_env": env, "eval_env": env, "num_agents": 2, "device": dev})
r.warmup()
for it in range(2):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record(); r.rollout(); e[1].record(); r.compute(); e[2].record(); r.train(); e[3].record()
    torch.cuda.synchronize()
    print(f"iteration {it}: rollout {e[0].elapsed_time(e[1]):.2f} ms, gae+bootstrap {e[1].elapsed_time(e[2]):.2f} ms, "
          f"ppo update (2 epochs) {e[2].elapsed_time(e[3]):.2f} ms", flush=True)
