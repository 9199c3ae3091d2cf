"""The qmix/main.py loop (runners.run_qmix) on 1 024 lockstep Switch2-v0 corridors (the qmix/_config.py default env) on the fused
top-level kernels: episode rollout in one launch, set-semantics chunk PER, Train_dqn with the recurrent Mix_Net as fused forward /
BPTT kernels replayed as a CUDA graph.  Switch2 pays +5 per agent that reaches the far side of the corridor and a step cost
otherwise; a random policy rarely gets both agents through, so the training score has to climb as epsilon anneals."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minimarl_b200.runners import run_qmix
from minimarl_b200.utils import MetricSink

EPISODES = int(sys.argv[1]) if len(sys.argv) > 1 else 150
ENV = sys.argv[2] if len(sys.argv) > 2 else "Switch2-v0"          # or "Checkers-v0" (partial observations)
args = argparse.Namespace(use_recurrent=True, buffer_limit=8192, batch_size=1024, update_iter=10, chunk_size=10, gamma=0.99,
                          eps=1e-6, alpha=0.6, beta=0.4, use_step_weight=False, update_alpha_beta=True, max_episodes=EPISODES,
                          grad_clip_norm=5, lr=1e-3, seed=42, max_step=50 if ENV.startswith("Switch") else 100, step_cost=-0.1 if ENV.startswith("Switch") else -0.01, max_epsilon=0.9, min_epsilon=0.05,
                          epsilon_anneal_episode=int(0.6 * EPISODES), update_target_interval=5, train_graph=True)
t0 = time.time()
sink = MetricSink(use_wandb=False, use_logging=False, every=10)


def log(rec):
    if rec["episode"] % 10 == 0 or rec["episode"] == EPISODES - 1:
        print(f"episode {rec['episode']:4d} eps {rec['eps']:.2f} train_score {rec['train_score']:8.3f} loss {rec['train_loss']:.4f} "
              f"wall {time.time() - t0:6.1f}s", flush=True)


timing = {}
history, net = run_qmix(args, n_envs=1024, episodes=EPISODES, device="cuda", log=log, timing=timing, sink=sink, env_id=ENV)
first, last = history[0]["train_score"], sum(h["train_score"] for h in history[-5:]) / 5
print("fused path:", timing["fused"], "| sink records:", len(sink.history))
print("improved" if last > first + 2.0 else "NOT improved", first, "->", last)
