"""Where one QMIX update (bench.py learner section: batch 4 096 episodes x T = 100, hidden 64, PER, double-Q) spends its time:
wall clock per update, host enqueue time per update, and the kernels of one update ranked by device time (torch profiler)."""
import argparse as ap
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from minimarl_b200.algorithms import QMix, QMixPolicy  # noqa: E402
from minimarl_b200.buffers import PrioritizedRecReplayBuffer  # noqa: E402
from minimarl_b200.envs.spaces import Box, Discrete  # noqa: E402

dev = torch.device("cuda:0")
T, A, D, NA, S, B = 100, 2, 47, 5, 10000, 4096
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
args = ap.Namespace(
    hidden_size=64, use_rnn_layer=True, use_orthogonal=True, gain=0.01, layer_N=1, recurrent_N=1,
    use_feature_normalization=True, use_ReLU=True, use_conv1d=False, prev_act_inp=False, batch_size=B,
    mixer_hidden_dim=32, hypernet_hidden_dim=64, hypernet_layers=2, use_per=True, per_eps=1e-6, per_nu=0.9,
    use_huber_loss=False, use_popart=False, lr=5e-4, tau=0.005, opti_eps=1e-5, episode_length=T,
    use_same_share_obs=True, use_double_q=True, gamma=0.99, max_grad_norm=10.0, epsilon_start=1.0,
    epsilon_finish=0.05, epsilon_anneal_time=50000, seed=1, fused_precision="tf32")
info = {"policy_0": {"obs_space": Box(np.zeros(D), np.ones(D)), "share_obs_space": Box(np.zeros(A * D), np.ones(A * D)),
                     "act_space": Discrete(NA)}}
buf = PrioritizedRecReplayBuffer(0.6, info, {"policy_0": [0, 1]}, S, T, True, False, device=dev)
pb = buf.policy_buffers["policy_0"]
g = torch.Generator(device=dev).manual_seed(42)
pb.obs.copy_((torch.rand(pb.obs.shape, device=dev, generator=g) * 100).round() / 100)
pb.share_obs.copy_(pb.obs.reshape(T + 1, S, A * D))
pb.acts.copy_(torch.nn.functional.one_hot(torch.randint(0, NA, (T, S, A), device=dev, generator=g), NA).float())
pb.rewards.copy_(torch.randn(pb.rewards.shape, device=dev, generator=g))
pb.dones.zero_(); pb.dones_env.zero_(); pb.dones_env[-1] = 1
pb.filled_i, pb.current_i = S, 0
buf.update_priorities(torch.arange(S, device=dev), torch.rand(S, device=dev, generator=g) * 0.99 + 0.01, "policy_0")
pconf = {"cent_obs_dim": A * D, "cent_act_dim": A * NA, "obs_space": info["policy_0"]["obs_space"],
         "share_obs_space": info["policy_0"]["share_obs_space"], "act_space": info["policy_0"]["act_space"]}
policy = QMixPolicy(args, {"args": args, "device": dev, "num_agents": A}, pconf)
trainer = QMix(args, A, B, {"policy_0": policy}, lambda a: "policy_0", device=dev, episode_length=T)
if len(sys.argv) > 1 and sys.argv[1] == "nofuse":
    policy.q_network.rnn.fused_mlp = False


def update():
    sample = buf.sample(B, beta=0.4, p_id="policy_0")
    info_, prio, idx = trainer.train_policy_on_batch(sample)
    buf.update_priorities(idx, prio, "policy_0")
    trainer.soft_target_updates()
    return info_


for _ in range(4):
    update()
torch.cuda.synchronize()
n = 20
t0 = time.perf_counter()
for _ in range(n):
    update()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / n:.2f} ms per update, wall {1e3 * (t2 - t0) / n:.2f} ms per update")
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        update()
    torch.cuda.synchronize()
ev = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
tot = sum(e.device_time_total for e in ev) / 3
print(f"device time of all kernels: {tot / 1e3:.2f} ms per update, {sum(e.count for e in ev) // 3} launches per update")
for e in sorted(ev, key=lambda e: -e.device_time_total)[:32]:
    print(f"{e.device_time_total / 3e3:8.3f} ms  x{e.count // 3:4d}  {e.key[:110]}")
