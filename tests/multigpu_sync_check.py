"""Multi-GPU learner check (not collected by pytest; run under torchrun on >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/multigpu_sync_check.py

Every rank builds the same QMIX learner (same seed) on DIFFERENT replay data, so the local gradients differ and the parameters
stay identical across ranks only if the flat gradient bucket is all-reduced before FlatAdam steps -- checked after eager updates
and after CUDA-graph replays (the NCCL all-reduce is captured inside the graph).  Same for R_MAPPO.train on different buffers.
Prints one line per check from rank 0; exit code 1 on a mismatch."""
import argparse as ap
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def in_sync(params):
    flat = torch.cat([p.detach().reshape(-1).double() for p in params])
    mine = torch.stack([flat.sum(), flat.abs().sum(), (flat * flat).sum()])
    every = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(every, mine)
    return bool(all(torch.equal(e, every[0]) for e in every) and torch.isfinite(mine).all())


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from minimarl_b200.algorithms import QMix, QMixPolicy, R_MAPPO, R_MAPPOPolicy
    from minimarl_b200.algorithms.qmix import GraphedUpdate
    from minimarl_b200.buffers import PrioritizedRecReplayBuffer, SharedReplayBuffer
    from minimarl_b200.envs.spaces import Box, Discrete
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True
    ok = True

    def report(name, good):
        nonlocal ok
        ok = ok and good
        if rank == 0:
            print(f"{name}: {'in sync' if good else 'MISMATCH'}", flush=True)

    # ---- QMIX ------------------------------------------------------------------------------------------------------------
    T, A, D, NA, S, B = 20, 2, 47, 5, 600, 256
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
    g = torch.Generator(device=dev).manual_seed(1000 + rank)                   # different data on every rank
    pb.obs.copy_(torch.rand(pb.obs.shape, device=dev, generator=g))
    pb.share_obs.copy_(pb.obs.reshape(T + 1, S, A * D))
    pb.acts.copy_(torch.nn.functional.one_hot(torch.randint(0, NA, (T, S, A), device=dev, generator=g), NA).float())
    pb.rewards.copy_(torch.randn(pb.rewards.shape, device=dev, generator=g))
    pb.dones.zero_(); pb.dones_env.zero_(); pb.dones_env[-1] = 1
    pb.filled_i, pb.current_i = S, 0
    buf.update_priorities(torch.arange(S, device=dev), torch.rand(S, device=dev, generator=g) * 0.99 + 0.01, "policy_0")
    pconf = {"cent_obs_dim": A * D, "cent_act_dim": A * NA, "obs_space": info["policy_0"]["obs_space"],
             "share_obs_space": info["policy_0"]["share_obs_space"], "act_space": info["policy_0"]["act_space"]}
    torch.manual_seed(7)
    policy = QMixPolicy(args, {"args": args, "device": dev, "num_agents": A}, pconf)
    trainer = QMix(args, A, B, {"policy_0": policy}, lambda a: "policy_0", device=dev, episode_length=T)
    report("qmix: initial parameters", in_sync(trainer.parameters))
    before = torch.cat([p.detach().reshape(-1) for p in trainer.parameters]).clone()
    for _ in range(3):
        sample = buf.sample(B, beta=0.4, p_id="policy_0")
        _, prio, idx = trainer.train_policy_on_batch(sample)
        buf.update_priorities(idx, prio, "policy_0")
        trainer.soft_target_updates()
    moved = float((torch.cat([p.detach().reshape(-1) for p in trainer.parameters]) - before).abs().max())
    report(f"qmix: after 3 eager updates (parameters moved by {moved:.2e})", in_sync(trainer.parameters) and moved > 0)
    graphed = GraphedUpdate(trainer, buf, B)
    for _ in range(5):
        graphed(0.4)
    report("qmix: after capture + 4 graph replays", in_sync(trainer.parameters))

    # ---- MAPPO -----------------------------------------------------------------------------------------------------------
    N, Tm = 64, 16
    margs = ap.Namespace(
        hidden_size=32, layer_N=1, recurrent_N=1, use_feature_normalization=True, use_ReLU=True, use_orthogonal=True,
        gain=0.01, use_conv1d=False, use_recurrent_policy=True, use_naive_recurrent_policy=False,
        use_policy_active_masks=True, use_value_active_masks=True, use_popart=False, use_valuenorm=True,
        use_huber_loss=True, use_max_grad_norm=True, use_clipped_value_loss=True, use_gae=True, ppo_epoch=3,
        train_batch_size=1, data_chunk_length=8, clip_param=0.2, entropy_coef=0.01, value_loss_coef=0.5,
        max_grad_norm=0.5, huber_delta=10.0, actor_lr=1e-4, critic_lr=1e-4, opti_eps=1e-5, weight_decay=0, gamma=0.99,
        gae_lambda=0.95, max_step=Tm, batch_size=N, fused_precision="tf32")
    space = Box(np.zeros(47), np.ones(47))
    torch.manual_seed(11)
    mpol = R_MAPPOPolicy(margs, space, space, Discrete(5), device=dev)
    mtr = R_MAPPO(margs, mpol, device=dev)
    mbuf = SharedReplayBuffer(margs, 2, space, space, Discrete(5), device=dev)
    mparams = list(mpol.actor.parameters()) + list(mpol.critic.parameters())
    report("mappo: initial parameters", in_sync(mparams))
    for it in range(2):
        r = lambda x: torch.rand(x.shape, device=dev, generator=g)
        mbuf.obs.copy_(r(mbuf.obs)); mbuf.share_obs.copy_(mbuf.obs)
        mbuf.rnn_states.copy_(r(mbuf.rnn_states) - 0.5); mbuf.rnn_states_critic.copy_(r(mbuf.rnn_states_critic) - 0.5)
        mbuf.actions.copy_((r(mbuf.actions) * 5).floor().clamp(max=4)); mbuf.action_log_probs.copy_(-1.6 + 0.1 * r(mbuf.action_log_probs))
        mbuf.value_preds.copy_(r(mbuf.value_preds)); mbuf.rewards.copy_(r(mbuf.rewards) - 0.3)
        mbuf.masks.copy_((r(mbuf.masks) > 0.1).float()); mbuf.active_masks.fill_(1.0)
        mbuf.compute_returns(mbuf.value_preds[-1].clone(), mtr.value_normalizer)
        mtr.prep_training()
        mtr.train(mbuf)
        report(f"mappo: after train() #{it + 1} ({'graph replays' if mtr._graph_state is not None else 'eager'})",
               in_sync(mparams + [mtr.value_normalizer.running_mean]))
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if ok else 1)          # (tearing the communicator down under live captured graphs can block)


if __name__ == "__main__":
    main()
