"""The vdn/main.py loop (runners.run_vdn) on 512 lockstep Checkers-v0 boards with full observations: rollout chunks and training
iterations replayed as CUDA graphs, chunk PER on the device.  The reference's own run of this loop (one env, CPU, 15 000 episodes,
1 h 43 min) ends at a test score of 8.00 (vdn/logs/vdn-1710766189.log); this prints the epsilon-greedy training score per episode
of all boards, which has to climb from the random level (about -2 -> food minus step costs) as epsilon anneals."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minimarl_b200.runners import run_vdn

EPISODES = int(sys.argv[1]) if len(sys.argv) > 1 else 120
args = argparse.Namespace(use_recurrent=True, buffer_limit=10000, batch_size=1024, update_iter=10, chunk_size=10, gamma=0.99,
                          eps=1e-6, alpha=0.6, beta=0.4, step_weight=0.99, use_step_weight=False, update_alpha_beta=True,
                          max_episodes=EPISODES, grad_clip_norm=5, lr=1e-3, seed=42, max_step=100, step_cost=-0.01,
                          max_epsilon=0.9, min_epsilon=0.05, epsilon_anneal_episode=int(0.6 * EPISODES),
                          update_target_interval=5, rollout_graph=True, train_graph=True)
t0 = time.time()


def log(rec):
    if rec["episode"] % 10 == 0 or rec["episode"] == EPISODES - 1:
        print(f"episode {rec['episode']:4d} eps {rec['eps']:.2f} train_score {rec['train_score']:8.3f} loss {rec['train_loss']:.4f} "
              f"wall {time.time() - t0:6.1f}s", flush=True)


history, net = run_vdn(args, n_envs=512, episodes=EPISODES, device="cuda", log=log)
first, last = history[0]["train_score"], sum(h["train_score"] for h in history[-5:]) / 5
print("improved" if last > first + 2.0 else "NOT improved", first, "->", last)
