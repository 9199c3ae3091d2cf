"""Kernels of the qmix/main.py loop on 4 096 lockstep Switch2 envs (bench.py `qmix_toplevel`: one 100-step episode of every env +
one Train_dqn.train of 10 iterations) ranked by device time (torch profiler), eager launches (train_graph off)."""
import argparse as ap
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minimarl_b200.runners import run_qmix  # noqa: E402

args = ap.Namespace(use_recurrent=True, buffer_limit=8192, batch_size=4096, update_iter=10, chunk_size=10, gamma=0.99, eps=1e-6,
                    alpha=0.8, beta=0.2, use_step_weight=False, update_alpha_beta=True, max_episodes=100000, grad_clip_norm=5,
                    lr=1e-3, seed=42, max_step=100, step_cost=-0.01, max_epsilon=0.9, min_epsilon=0.05,
                    epsilon_anneal_episode=60000, update_target_interval=20, rollout_graph=True, train_graph=False,
                    test_interval=0, test_episodes=0)
run_qmix(args, 4096, 3, device="cuda")                    # warm-up (fills the buffer with the eps = 1 episodes first)
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    args.warm_up = False
    timing = {}
    run_qmix(args, 4096, 3, device="cuda", timing=timing)
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
tot = sum(e.device_time_total for e in ev) / 3
print(f"device time of all kernels: {tot / 1e3:.2f} ms per (episode + train), {sum(e.count for e in ev) // 3} launches; "
      f"rollout {1e3 * timing['rollout_s'] / 3:.2f} ms, train {1e3 * timing['train_s'] / 3:.2f} ms (wall, eager)")
for e in sorted(ev, key=lambda e: -e.device_time_total)[:30]:
    print(f"{e.device_time_total / 3e3:8.3f} ms  x{e.count // 3:4d}  {e.key[:110]}")
