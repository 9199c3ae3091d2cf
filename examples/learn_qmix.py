import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from minimarl_b200 import BatchedCheckers
from minimarl_b200.runners import OffPolicyRunner, MappoRunner
from test_qmix_gpu import _args
dev = torch.device('cuda')
N, T = 512, 100
args = _args(algorithm_name="qmix", episode_length=T, batch_size=256, buffer_size=8192, use_eval=False, use_soft_update=True,
             num_env_steps=400 * N * T, train_interval_episode=1, hard_update_interval_episode=200, log_interval=10**12,
             eval_interval=10**12, per_alpha=0.6, per_beta_start=0.4, use_reward_normalization=False, num_random_episodes=5,
             fused_precision="tf32", epsilon_anneal_time=100 * N * T, lr=5e-4)
env = BatchedCheckers(N, max_steps=T, device=dev)
runner = OffPolicyRunner({"args": args, "train_env": env, "eval_env": env, "num_agents": 2, "device": dev})
t0 = time.time()
for it in range(300):
    runner.run()
    if it % 25 == 0 or it == 299:
        ev = float(runner.eval())
        print(f"iter {it:4d} env_steps {runner.total_env_steps:9d} train_return {float(runner.env_infos['average_episode_rewards']):8.3f} greedy_eval_return {ev:8.3f} loss {float(runner.train_infos[-1]['loss']):.4f} wall {time.time()-t0:6.1f}s", flush=True)
