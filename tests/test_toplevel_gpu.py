"""GPU parity: top-level vdn/ and qmix/ mirror (Q_Net, Mix_Net, chunk PER on the device, Target_Dqn) vs values
recorded from the reference modules (tests/golden/toplevel.npz).  Tree contents and sampled indices are exact;
network outputs 2e-5 abs (fp32, cuBLAS vs CPU); loss 1e-4 relative."""
import argparse
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from minimarl_b200.envs.spaces import Box, Discrete

pytestmark = pytest.mark.gpu


def _g():
    return np.load(os.path.join(GOLDEN, "toplevel.npz"))


def _args(g):
    return argparse.Namespace(use_recurrent=True, buffer_limit=50, batch_size=8, update_iter=2, chunk_size=10,
                              gamma=float(g["arg_gamma"]), eps=float(g["arg_eps"]), alpha=float(g["arg_alpha"]),
                              beta=float(g["arg_beta"]), step_weight=float(g["arg_step_weight"]), use_step_weight=True,
                              update_alpha_beta=True, max_episodes=int(g["arg_max_episodes"]),
                              grad_clip_norm=float(g["arg_grad_clip_norm"]), lr=float(g["arg_lr"]), seed=0)


def _nets(cuda, g, args):
    from minimarl_b200.algorithms import Q_Net
    obs_space = [Box(np.zeros(94), np.ones(94)) for _ in range(2)]
    act_space = [Discrete(5) for _ in range(2)]
    beh, tgt = Q_Net(obs_space, act_space, args).to(cuda), Q_Net(obs_space, act_space, args).to(cuda)
    beh.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("beh.")})
    tgt.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("tgt.")})
    return beh, tgt


def test_qnet_mixnet_and_td_error_match_reference(cuda):
    from minimarl_b200.algorithms import Mix_Net, cal_td_error
    torch.backends.cuda.matmul.allow_tf32 = False
    g = _g()
    args = _args(g)
    beh, tgt = _nets(cuda, g, args)
    assert sum(p.numel() for p in beh.parameters()) == 29322
    obs, hid = torch.from_numpy(g["fwd_obs"]).to(cuda), torch.from_numpy(g["fwd_hidden"]).to(cuda)
    with torch.no_grad():
        q, h = beh(obs, hid)
    assert np.abs(q.cpu().numpy() - g["fwd_q"]).max() < 2e-5 and np.abs(h.cpu().numpy() - g["fwd_h"]).max() < 2e-5
    td = cal_td_error(torch.from_numpy(g["td_action"]).to(cuda).unsqueeze(0), [[-0.01, 9.99]], [0], q[:1],
                      torch.from_numpy(g["td_target_q"]).to(cuda), args.gamma)
    assert abs(td.item() - float(g["td_error"])) < 2e-5
    a, nh, qq = beh.sample_action(obs, hid, epsilon=0.0)
    assert torch.equal(a, q.argmax(2).float()) and a.dtype == torch.float32
    a, _, _ = beh.sample_action(obs.repeat(200, 1, 1), hid.repeat(200, 1, 1), epsilon=1.0)
    assert a.min() >= 0 and a.max() <= 4
    mix = Mix_Net([Box(np.zeros(3), np.ones(3)) for _ in range(2)], args).to(cuda)
    mix.load_state_dict({k[4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("mix.") and not k.startswith("mix_")})
    with torch.no_grad():
        y, mh = mix(torch.from_numpy(g["mix_q"]).to(cuda), torch.from_numpy(g["mix_obs"]).to(cuda), torch.from_numpy(g["mix_hidden"]).to(cuda))
    assert np.abs(y.cpu().numpy() - g["mix_out"]).max() < 2e-5 and np.abs(mh.cpu().numpy() - g["mix_h"]).max() < 2e-5


def test_chunk_per_and_dqn_train_match_reference(cuda):
    from minimarl_b200.algorithms import Prioritized_Experience_Replay, Target_Dqn
    torch.backends.cuda.matmul.allow_tf32 = False
    g = _g()
    args = _args(g)
    per = Prioritized_Experience_Replay(args, n_agents=2, obs_dim=94, device=cuda)
    n_add = g["chunk_td"].shape[0]
    for i in range(n_add):                     # 70 adds into 50 slots: the last 20 evict the arg-min leaf
        n = per.collect_sample([g["chunk_states"][i], g["chunk_actions"][i], g["chunk_rewards"][i],
                                g["chunk_next_states"][i], g["chunk_dones"][i]], float(g["chunk_td"][i]), warm_up=True)
    assert n == 50
    assert np.array_equal(per.sum_tree.priority_tree.cpu().numpy(), g["tree_after_add"])
    u = g["uniforms"]
    s = per.sample(8, 10, uniforms=u[0])
    assert np.array_equal(s[5].cpu().numpy(), g["sample_idx"])
    for got, key in zip(s[:5], ["states", "actions", "rewards", "next_states", "dones"]):
        assert np.array_equal(got.cpu().numpy(), g["sample_" + key]), key
    assert np.allclose(s[6].cpu().numpy(), g["sample_w"], rtol=1e-6)
    assert np.array_equal(per.sum_tree.priority_tree.cpu().numpy(), g["tree_after_sample"])   # whole-tree decay x0.99
    assert per.alpha == float(g["alpha_after_sample"]) and per.beta == float(g["beta_after_sample"])
    beh, tgt = _nets(cuda, g, args)
    trainer = Target_Dqn(per, beh, tgt, args, cuda)
    opt = torch.optim.Adam(beh.parameters(), lr=args.lr)
    loss = trainer.train(target_network=tgt, optimizer=opt, epsilon=0.5, uniforms=u[1:])
    assert abs(loss.item() - float(g["train_loss"])) < 1e-4 * float(g["train_loss"])
    assert np.allclose(per.sum_tree.priority_tree.cpu().numpy(), g["tree_after_train"], rtol=1e-4, atol=1e-7)
    for name, p in beh.state_dict().items():
        assert np.allclose(p.cpu().numpy(), g["beh_after." + name], atol=2e-3, rtol=1e-3), name   # 2 Adam steps, lr 1e-3


def test_qmix_toplevel_train_runs_with_lockstep_envs(cuda):
    """4 096-env config shape (here 256 envs): lockstep rollout -> device chunk PER -> Train_dqn with Mix_Net."""
    from minimarl_b200 import BatchedSwitch
    from minimarl_b200.algorithms import Mix_Net, Prioritized_Experience_Replay, Q_Net, Train_dqn, cal_td_error
    g = _g()
    args = _args(g)
    args.buffer_limit, args.batch_size, args.use_step_weight = 512, 32, False
    N, L = 256, args.chunk_size
    env = BatchedSwitch(N, n_agents=2, max_steps=100, step_cost=-0.01, device=cuda)     # qmix default env
    spaces = env.observation_space
    q, tq = Q_Net(spaces, env.action_space, args).to(cuda), Q_Net(spaces, env.action_space, args).to(cuda)
    mix, tmix = Mix_Net(spaces, args).to(cuda), Mix_Net(spaces, args).to(cuda)
    per = Prioritized_Experience_Replay(args, n_agents=2, obs_dim=3, device=cuda)
    obs = env.reset()
    hid, thid = q.init_hidden(N), tq.init_hidden(N)
    chunk = {k: [] for k in ["s", "a", "r", "s2", "d"]}
    td_sum = torch.zeros(N, device=cuda)
    with torch.no_grad():
        for t in range(L):
            a, hid2, bq = q.sample_action(obs, hid, epsilon=0.5)
            obs2, rew, done = env.step(a)
            tgt_q, thid = tq(obs2, thid)
            d = done.all(1).float()
            td_sum += cal_td_error(a, rew, d, bq, tgt_q, args.gamma)
            for k, v in zip(["s", "a", "r", "s2", "d"], [obs, a, rew, obs2, d]):
                chunk[k].append(v)
            obs, hid = obs2, hid2
    stack = {k: torch.stack(v, 1) for k, v in chunk.items()}
    for e in range(64):                                                  # one chunk per env, like the reference loop
        per.collect_sample([stack["s"][e], stack["a"][e], stack["r"][e], stack["s2"][e], stack["d"][e]], td_sum[e].item(), True)
    opt = torch.optim.Adam(list(q.parameters()) + list(mix.parameters()), lr=1e-3)
    loss = Train_dqn(args, cuda).train(per, q, mix, tq, tmix, opt, 0.5)
    assert torch.isfinite(loss)


def test_eval_adaptor_matches_reference_test_execute(cuda):
    """``Test.execute`` (vdn/_test.py) recorded from the reference on the oracle env (tests/golden/vdn_test.npz):
    greedy score exact (sums of the same float rewards), TD loss 1e-4 relative (fp32 network on GPU vs CPU)."""
    import argparse
    from minimarl_b200 import BatchedCheckers
    from minimarl_b200.algorithms import Q_Net, Test
    g = np.load(os.path.join(GOLDEN, "vdn_test.npz"))
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for tag in ("s21_", "s22_"):
            args = argparse.Namespace(use_recurrent=bool(g["use_recurrent"]), batch_size=8, chunk_size=10, update_iter=2,
                                      test_episodes=int(g["test_episodes"]), gamma=float(g["gamma"]), grad_clip_norm=5)
            env = BatchedCheckers(int(g["test_episodes"]), full_observable=True, max_steps=int(g[tag + "max_step"]),
                                  step_cost=float(g["step_cost"]), device=cuda)
            nets = []
            for name in ("beh.", "tgt."):
                net = Q_Net(env.observation_space, env.action_space, args).to(cuda)
                net.load_state_dict({k[len(tag + name):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(tag + name)})
                nets.append(net)
            score, loss = Test(env, args, cuda).execute(nets[0], nets[1])
            assert abs(score - float(g[tag + "score"])) < 1e-5, (score, float(g[tag + "score"]))
            assert abs(loss - float(g[tag + "loss"])) < 1e-4 * float(g[tag + "loss"]), (loss, float(g[tag + "loss"]))
        # one board, two episodes: two passes give the same averages
        env1 = BatchedCheckers(1, full_observable=True, max_steps=int(g["s22_max_step"]), step_cost=float(g["step_cost"]), device=cuda)
        s1, l1 = Test(env1, args, cuda).execute(nets[0], nets[1])
        assert abs(s1 - score) < 1e-6 and abs(l1 - loss) < 1e-5 * loss
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32


def test_render_shim(cuda):
    import minimarl_b200
    env = minimarl_b200.make("ma_gym:Checkers-v0", max_steps=10, step_cost=-0.01, full_observable=False)
    env.reset()
    text = env.render(mode="ansi")
    rows = text.split("\n")
    assert len(rows) == 3 and rows[0].split() == ["R", "Y", "R", "Y", "R", "Y", "1", "."] and rows[2].split()[6] == "2"
    env.step([1, 4])                                    # agent 1 moves left onto the lemon at (0, 5)
    assert env.render(mode="ansi").split("\n")[0].split() == ["R", "Y", "R", "Y", "R", "1", ".", "."]


def test_qmix_toplevel_train_matches_reference(cuda):
    """qmix/_train.py Train_dqn.train recorded from the reference (tests/golden/qmix_toplevel.npz): 80 chunk adds into
    64 slots (whole-tree-argsort eviction), two update iterations with Q_Net + recurrent Mix_Net on both sides."""
    from minimarl_b200.algorithms import Mix_Net, Prioritized_Experience_Replay, Q_Net, Train_dqn
    torch.backends.cuda.matmul.allow_tf32 = False
    g = np.load(os.path.join(GOLDEN, "qmix_toplevel.npz"))
    args = argparse.Namespace(use_recurrent=True, buffer_limit=64, batch_size=32, update_iter=2, chunk_size=10,
                              gamma=float(g["arg_gamma"]), eps=float(g["arg_eps"]), alpha=float(g["arg_alpha"]),
                              beta=float(g["arg_beta"]), use_step_weight=False, update_alpha_beta=bool(g["arg_update_alpha_beta"]),
                              max_episodes=int(g["arg_max_episodes"]), grad_clip_norm=float(g["arg_grad_clip_norm"]),
                              lr=float(g["arg_lr"]), seed=0)
    spaces = [Box(np.zeros(3), np.ones(3)) for _ in range(2)]
    acts = [Discrete(5) for _ in range(2)]
    load = lambda m, pre: m.load_state_dict({k[len(pre):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(pre)})
    q, tq = Q_Net(spaces, acts, args).to(cuda), Q_Net(spaces, acts, args).to(cuda)
    mix, tmix = Mix_Net(spaces, args).to(cuda), Mix_Net(spaces, args).to(cuda)
    load(q, "q."); load(tq, "tq."); load(mix, "mixnet."); load(tmix, "tmixnet.")
    per = Prioritized_Experience_Replay(args, n_agents=2, obs_dim=3, device=cuda)
    for i in range(g["chunk_td"].shape[0]):
        per.collect_sample([g["chunk_states"][i], g["chunk_actions"][i], g["chunk_rewards"][i], g["chunk_next_states"][i],
                            g["chunk_dones"][i]], float(g["chunk_td"][i]), warm_up=True)
    assert np.array_equal(per.sum_tree.priority_tree.cpu().numpy(), g["tree_after_add"])
    opt = torch.optim.Adam(list(q.parameters()) + list(mix.parameters()), lr=args.lr)
    loss = Train_dqn(args, cuda).train(per, q, mix, tq, tmix, opt, 0.5, uniforms=g["uniforms"])
    want = g["step_losses"].sum(1).mean()
    assert abs(loss.item() - want) < 1e-4 * want, (loss.item(), want)
    assert per.alpha == float(g["alpha_after"]) and per.beta == float(g["beta_after"])
    assert np.allclose(per.sum_tree.priority_tree.cpu().numpy(), g["tree_after_train"], rtol=2e-4, atol=1e-7)
    for name, p in q.qmix_state_dict().items():              # the key spelling of qmix/_network.py
        assert np.allclose(p.cpu().numpy(), g["q_after." + name], atol=2e-3, rtol=1e-3), name
    for name, p in mix.state_dict().items():
        assert np.allclose(p.cpu().numpy(), g["mixnet_after." + name], atol=2e-3, rtol=1e-3), name


@pytest.mark.parametrize("fixture,cap,D,splits", [("toplevel.npz", 50, 94, [70]), ("toplevel.npz", 50, 94, [30, 25, 15]),
                                                  ("qmix_toplevel.npz", 64, 3, [80]), ("qmix_toplevel.npz", 64, 3, [1, 63, 1, 15])])
def test_batched_chunk_insert_equals_sequential_reference_adds(cuda, fixture, cap, D, splits):
    """collect_samples_batch (mm_heaptree_add_batch: K SumTree.add calls in one launch, eviction inside the batch
    included) leaves the tree bit-identical to the reference's one-by-one adds, and every slot holds the payload of
    the chunk the reference put there (checked through the stratified sample recorded from the reference)."""
    from minimarl_b200.algorithms import Prioritized_Experience_Replay
    g = np.load(os.path.join(GOLDEN, fixture))
    args = argparse.Namespace(use_recurrent=True, buffer_limit=cap, batch_size=8, update_iter=2, chunk_size=10,
                              eps=float(g["arg_eps"]), alpha=float(g["arg_alpha"]), beta=float(g["arg_beta"]),
                              step_weight=1.0, use_step_weight=False, update_alpha_beta=False, max_episodes=1, seed=0)
    per = Prioritized_Experience_Replay(args, n_agents=2, obs_dim=D, device=cuda)
    seq = Prioritized_Experience_Replay(args, n_agents=2, obs_dim=D, device=cuda)
    lo = 0
    for k in splits:
        sl = slice(lo, lo + k)
        n = per.collect_samples_batch([g["chunk_states"][sl], g["chunk_actions"][sl], g["chunk_rewards"][sl],
                                       g["chunk_next_states"][sl], g["chunk_dones"][sl]], g["chunk_td"][sl], True, host_pow=True)
        lo += k
        assert n == min(cap, lo)
    assert np.array_equal(per.sum_tree.priority_tree.cpu().numpy(), g["tree_after_add"])
    for i in range(lo):
        seq.collect_sample([g["chunk_states"][i], g["chunk_actions"][i], g["chunk_rewards"][i], g["chunk_next_states"][i],
                            g["chunk_dones"][i]], float(g["chunk_td"][i]), True)
    for name in ["states", "actions", "rewards", "next_states", "dones"]:
        assert torch.equal(getattr(per, name)[:cap], getattr(seq, name)[:cap]), name
    # device pow path: same slots, priorities within 1 ulp
    dev = Prioritized_Experience_Replay(args, n_agents=2, obs_dim=D, device=cuda)
    dev.collect_samples_batch([g["chunk_states"][:lo], g["chunk_actions"][:lo], g["chunk_rewards"][:lo],
                               g["chunk_next_states"][:lo], g["chunk_dones"][:lo]], torch.from_numpy(g["chunk_td"][:lo]).to(cuda), True)
    assert np.allclose(dev.sum_tree.priority_tree.cpu().numpy(), g["tree_after_add"], rtol=1e-14)
    assert torch.equal(dev.states[:cap], seq.states[:cap])


def test_batched_tree_kernels_equal_one_by_one_at_scale(cuda):
    """mm_heaptree_add_batch and the batched mm_heaptree_update beyond one staging pass (k, n > 1 024), non-power-of-two
    capacity, duplicate update targets: trees and slots bit-identical to the same operations issued one at a time (the
    single-update / arg-min kernels, already pinned against the reference fixtures)."""
    from minimarl_b200.algorithms.toplevel import SumTree
    from minimarl_b200 import _lib
    cap = 3001
    g = torch.Generator(device=cuda).manual_seed(7)
    a, b = SumTree(cap, cuda), SumTree(cap, cuda)
    n_data = 0
    for k in (2500, 1700, 1300):                       # fill, then evict (also inside the batch)
        prio = (torch.rand(k, dtype=torch.float64, device=cuda, generator=g) * 4 + 1e-3)
        slots = torch.empty(k, dtype=torch.int64, device=cuda)
        _lib.check(_lib.lib().mm_heaptree_add_batch(_lib.ptr(a.priority_tree), cap, n_data, _lib.ptr(prio), k, _lib.ptr(slots),
                                                    _lib.stream()))
        want, nd = [], n_data
        for i in range(k):
            s = nd if nd < cap else int(b.argmin_leaf().item())
            nd = min(cap, nd + 1)
            b.update(torch.tensor([s + cap - 1], device=cuda), prio[i:i + 1])
            want.append(s)
        n_data = min(cap, n_data + k)
        assert slots.cpu().tolist() == want
        assert torch.equal(a.priority_tree, b.priority_tree)
    idx = torch.randint(0, cap, (4096,), device=cuda, generator=g) + cap - 1          # many duplicates
    pr = torch.rand(4096, dtype=torch.float64, device=cuda, generator=g) + 0.01
    a.update(idx, pr)                                   # one launch, four staging passes
    for i in range(4096):
        b.update(idx[i:i + 1], pr[i:i + 1])             # n = 1: the sequential kernel
    assert torch.equal(a.priority_tree, b.priority_tree)
