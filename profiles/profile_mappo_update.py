"""Host enqueue time vs wall clock of R_MAPPO.train at the bench size (16 384 envs x 128 steps, 15 epochs, mappo/_config.py
defaults), and the kernels of one update ranked by device time (torch profiler)."""
import argparse as ap
import sys
import time

import torch

sys.path.insert(0, ".")
from minimarl_b200 import BatchedCheckers  # noqa: E402
from minimarl_b200.runners import MappoRunner  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
n_envs, T = 16384, 128
args = ap.Namespace(
    hidden_size=32, layer_N=1, recurrent_N=1, use_feature_normalization=True, use_ReLU=True, use_orthogonal=True,
    gain=0.01, use_conv1d=False, use_recurrent_policy=True, use_naive_recurrent_policy=False,
    use_policy_active_masks=True, use_value_active_masks=True, use_popart=False, use_valuenorm=True,
    use_huber_loss=True, use_max_grad_norm=True, use_clipped_value_loss=True, use_gae=True, ppo_epoch=15,
    train_batch_size=1, data_chunk_length=8, clip_param=0.2, entropy_coef=0.01, value_loss_coef=0.5,
    max_grad_norm=0.5, huber_delta=10.0, actor_lr=1e-4, critic_lr=1e-4, opti_eps=1e-5, weight_decay=0, gamma=0.99,
    gae_lambda=0.95, max_step=T, batch_size=n_envs, fused_precision="tf32", max_episodes=1000, use_eval=False,
    eval_interval=10, use_linear_lr_decay=False, seed=1)
env = BatchedCheckers(n_envs, max_steps=100, auto_reset=True, device=dev)
r = MappoRunner({"args": args, "train_env": env, "eval_env": env, "num_agents": 2, "device": dev})
r.warmup()
r.run_episode(0)
r.rollout(); r.compute()
torch.cuda.synchronize()
t0 = time.perf_counter()
r.train()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"R_MAPPO.train: host enqueue {1e3 * (t1 - t0):.1f} ms, wall {1e3 * (t2 - t0):.1f} ms")
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    r.train()
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(e.device_time_total for e in ev)
print(f"device time of all kernels: {tot / 1e3:.1f} ms per update, {sum(e.count for e in ev)} launches")
for e in sorted(ev, key=lambda e: -e.device_time_total)[:24]:
    print(f"{e.device_time_total / 1e3:8.2f} ms  x{e.count:4d}  {e.key[:110]}")
