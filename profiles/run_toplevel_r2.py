"""Launches the fused top-level kernels once each (after warm-up) for ncu: rollout, target sequence, train passes."""
import argparse as ap
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from minimarl_b200 import BatchedCheckers, BatchedSwitch
from minimarl_b200.algorithms import Mix_Net, Prioritized_Experience_Replay, Q_Net
from minimarl_b200.algorithms import toplevel_fused as tf

dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "switch2"
N, T, L, B = 4096, 100, 10, 4096
args = ap.Namespace(use_recurrent=True, buffer_limit=8192, batch_size=B, update_iter=10, chunk_size=L, gamma=0.99, eps=1e-6,
                    alpha=0.8, beta=0.2, use_step_weight=False, update_alpha_beta=True, max_episodes=100000, grad_clip_norm=5,
                    lr=1e-3, seed=42, max_step=T, fast_tree=True)
if name == "switch2":
    env = BatchedSwitch(N, n_agents=2, max_steps=T, step_cost=-0.01, device=dev)
else:
    env = BatchedCheckers(N, full_observable=(name == "checkers_full"), max_steps=T, step_cost=-0.01, device=dev)
beh = Q_Net(env.observation_space, env.action_space, args).to(dev)
tgt = Q_Net(env.observation_space, env.action_space, args).to(dev)
per = Prioritized_Experience_Replay(args, n_agents=2, obs_dim=env.obs_dim, device=dev)
out = None
for _ in range(3):
    out = tf.rollout_episode(beh, env, T, L, 0.5, seed=1, out=out)
    q_tgt = tf.qnet_seq(tgt, out["next_states"], N, L, C_=T // L)
    td, valid = tf.td_chunks(q_tgt, out["rewards"], out["dones"], out["behv"], out["alive"], 0.99)
    for c in range(T // L):
        per.collect_samples_batch([out["states"][c], out["actions"][c], out["rewards"][c], out["next_states"][c], out["dones"][c]],
                                  td[c], warm_up=False, valid=valid[c])
    data, node, w = per.sample_indices(B)
    q_beh, save = tf.qnet_seq(beh, per.states, B, L, row_index=data, dones=per.dones, reset_on_done=True, save=True)
    tf.qnet_backward(beh, per.states, data, per.dones, save, torch.randn_like(q_beh[0]), B, L)
    if name != "checkers_full":
        mix = Mix_Net(env.observation_space, args).to(dev)
        qtot, msave = tf.mix_seq(mix, per.states, data, per.dones, q_beh, per.actions, B, L, save=True)
        tf.mix_backward(mix, per.states, data, per.dones, q_beh, per.actions, msave, qtot, B, L)
    per.update(node, torch.rand(B, device=dev))
torch.cuda.synchronize()
print("done", name)
