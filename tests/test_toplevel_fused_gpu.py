"""GPU parity of the fused top-level kernels (csrc/toplevel_qnet.cu, toplevel_mix.cu, heaptree_batch.cu).

* Q_Net / Mix_Net sequence kernels vs the PyTorch modules (whose parity with the reference is pinned on
  tests/golden/toplevel.npz / qmix_toplevel.npz in test_toplevel_gpu.py): outputs 2e-5 abs, every weight gradient 1e-4 of
  its tensor's max (fp32 both sides, different summation order).
* fused rollout: the recorded action sequence replayed through the C oracle env reproduces states / rewards / dones bit for
  bit; Q values vs the module on the recorded states 2e-5; exploration law (one draw per env).
* trainers: fused iteration vs the module iteration on the same draws (loss 1e-5 rel, weights after the Adam step 1e-5).
* batch PER kernels vs a numpy statement of their contract.
"""
import argparse

import numpy as np
import pytest
import torch

from minimarl_b200.envs.spaces import Box, Discrete

pytestmark = pytest.mark.gpu


def _args(**kw):
    d = dict(use_recurrent=True, buffer_limit=256, batch_size=70, update_iter=1, chunk_size=10, gamma=0.99, eps=1e-6, alpha=0.8,
             beta=0.2, step_weight=1.0, use_step_weight=False, update_alpha_beta=True, max_episodes=100, grad_clip_norm=5.0,
             lr=1e-3, seed=3, max_step=100)
    d.update(kw)
    return argparse.Namespace(**d)


def _qnet(cuda, D, args, seed=0):
    from minimarl_b200.algorithms import Q_Net
    torch.manual_seed(seed)
    net = Q_Net([Box(np.zeros(D), np.ones(D))] * 2, [Discrete(5)] * 2, args).to(cuda)
    with torch.no_grad():
        for p in net.parameters():                      # default init has zero-ish biases; make every term matter
            p.add_(0.05 * torch.randn_like(p))
    return net


def _module_chunk_forward(net, states, dones):
    """The reference's chunk loop (vdn/_train.py:66-86): hidden reset after done steps.  states [B,L,A,D], dones [B,L]."""
    B, L = dones.shape
    h = net.init_hidden(B)
    qs = []
    for l in range(L):
        q, h = net(states[:, l], h)
        qs.append(q)
        h = h * (1 - dones[:, l]).view(-1, 1, 1)
    return torch.stack(qs, 1)                           # [B,L,A,NA]


@pytest.mark.parametrize("D,B", [(94, 70), (3, 33), (47, 128)])
def test_qnet_seq_and_backward_match_module(cuda, D, B):
    from minimarl_b200.algorithms import toplevel_fused as tf
    torch.backends.cuda.matmul.allow_tf32 = False
    args = _args()
    L = 10
    net = _qnet(cuda, D, args)
    g = torch.Generator(device=cuda).manual_seed(1)
    rows = 200
    states = torch.rand((rows, L, 2, D), device=cuda, generator=g)
    dones = (torch.rand((rows, L, 1), device=cuda, generator=g) < 0.15).float()
    idx = torch.randperm(rows, device=cuda, generator=g)[:B].contiguous()
    want = _module_chunk_forward(net, states[idx], dones[idx, :, 0])
    q, save = tf.qnet_seq(net, states, B, L, row_index=idx, dones=dones, reset_on_done=True, save=True)
    assert (q[0] - want).abs().max().item() < 2e-5
    q2 = tf.qnet_seq(net, states, B, L, row_index=idx, dones=dones, reset_on_done=True)
    assert torch.equal(q2, q)
    # backward: d/dW of sum(q * G)
    G = torch.randn(want.shape, device=cuda, generator=g)
    net.zero_grad()
    (want * G).sum().backward()
    ref = [p.grad.clone() for p in tf.qnet_tensors(net)]
    for p in net.parameters():
        p.grad = None
    tf.qnet_backward(net, states, idx, dones, save, G.contiguous(), B, L)
    for got, w in zip([p.grad for p in tf.qnet_tensors(net)], ref):
        assert (got - w).abs().max().item() <= 1e-4 * max(1.0, w.abs().max().item()), (got - w).abs().max().item()
    # accumulate adds
    tf.qnet_backward(net, states, idx, dones, save, G.contiguous(), B, L, accumulate=True)
    for got, w in zip([p.grad for p in tf.qnet_tensors(net)], ref):
        assert (got - 2 * w).abs().max().item() <= 2e-4 * max(1.0, w.abs().max().item())


def test_qnet_seq_carries_state_across_chunks(cuda):
    """C > 1: the rollout layout [C,N,L,A,D]; the hidden state runs through all C*L steps (target pass of the rollout)."""
    from minimarl_b200.algorithms import toplevel_fused as tf
    torch.backends.cuda.matmul.allow_tf32 = False
    args = _args()
    net = _qnet(cuda, 94, args, seed=5)
    Cn, N, L = 3, 45, 10
    obs = torch.rand((Cn, N, L, 2, 94), device=cuda)
    q = tf.qnet_seq(net, obs, N, L, C_=Cn)
    h = net.init_hidden(N)
    for c in range(Cn):
        for l in range(L):
            with torch.no_grad():
                want, h = net(obs[c, :, l], h)
            assert (q[c, :, l] - want).abs().max().item() < 3e-5


@pytest.mark.parametrize("S_D", [3, 47])
def test_mixnet_seq_and_backward_match_module(cuda, S_D):
    from minimarl_b200.algorithms import Mix_Net
    from minimarl_b200.algorithms import toplevel_fused as tf
    torch.backends.cuda.matmul.allow_tf32 = False
    args = _args()
    torch.manual_seed(2)
    mix = Mix_Net([Box(np.zeros(S_D), np.ones(S_D))] * 2, args).to(cuda)
    with torch.no_grad():
        for p in mix.parameters():
            p.add_(0.05 * torch.randn_like(p))
    B, L, rows, NA = 70, 10, 150, 5
    g = torch.Generator(device=cuda).manual_seed(4)
    states = torch.rand((rows, L, 2, S_D), device=cuda, generator=g)
    dones = (torch.rand((rows, L, 1), device=cuda, generator=g) < 0.15).float()
    actions = torch.randint(0, NA, (rows, L, 2), device=cuda, generator=g).float()
    idx = torch.randperm(rows, device=cuda, generator=g)[:B].contiguous()
    q_all = torch.randn((B, L, 2, NA), device=cuda, generator=g).requires_grad_(True)

    def module_pass(take_max):
        mh = mix.init_hidden(B)
        outs = []
        for l in range(L):
            if take_max:
                qa = q_all[:, l].max(dim=2)[0]
            else:
                qa = q_all[:, l].gather(2, actions[idx, l].long().unsqueeze(-1)).squeeze(-1)
            y, mh = mix(qa, states[idx, l], mh)
            outs.append(y)
            mh = mh * (1 - dones[idx, l])
        return torch.cat(outs, 1)                       # [B,L]

    want_max = module_pass(True)
    got_max = tf.mix_seq(mix, states, idx, dones, q_all.detach().contiguous(), None, B, L)
    assert (got_max - want_max).abs().max().item() < 3e-5
    want = module_pass(False)
    got, msave = tf.mix_seq(mix, states, idx, dones, q_all.detach().contiguous(), actions, B, L, save=True)
    assert (got - want).abs().max().item() < 3e-5
    G = torch.randn((B, L), device=cuda, generator=g)
    mix.zero_grad()
    (want * G).sum().backward()
    ref = [p.grad.clone() for p in tf.mix_tensors(mix)]
    ref_dq = q_all.grad.clone()
    for p in mix.parameters():
        p.grad = None
    dq = tf.mix_backward(mix, states, idx, dones, q_all.detach().contiguous(), actions, msave, G.contiguous(), B, L)
    assert (dq - ref_dq).abs().max().item() <= 1e-4 * max(1.0, ref_dq.abs().max().item())
    for got_g, w in zip([p.grad for p in tf.mix_tensors(mix)], ref):
        assert (got_g - w).abs().max().item() <= 1e-4 * max(1.0, w.abs().max().item()), (got_g - w).abs().max().item()


@pytest.mark.parametrize("env_name", ["checkers_full", "checkers_partial", "switch2"])
def test_fused_rollout_replays_through_the_oracle(cuda, env_name):
    from minimarl_b200 import BatchedCheckers, BatchedSwitch
    from minimarl_b200.algorithms import toplevel_fused as tf
    from oracle.magym_c import CheckersBatchRef, SwitchBatchRef
    torch.backends.cuda.matmul.allow_tf32 = False
    N, T, L, eps = 77, 40, 10, 0.3
    if env_name == "switch2":
        env = BatchedSwitch(N, n_agents=2, max_steps=30, step_cost=-0.1, device=cuda)
        ref = SwitchBatchRef(N, 2, False, 30, -0.1, True)
    else:
        full = env_name == "checkers_full"
        env = BatchedCheckers(N, full_observable=full, max_steps=25, step_cost=-0.01, device=cuda)
        ref = CheckersBatchRef(N, full, 25, -0.01)
    args = _args()
    net = _qnet(cuda, env.obs_dim, args, seed=7)
    out = tf.rollout_episode(net, env, T, L, eps, seed=11, step_offset=5)
    torch.cuda.synchronize()
    Cn = T // L
    st = out["states"].cpu().numpy(); ns = out["next_states"].cpu().numpy()
    act = out["actions"].cpu().numpy(); rew = out["rewards"].cpu().numpy(); dn = out["dones"].cpu().numpy()
    alive = out["alive"].cpu().numpy(); behv = out["behv"]
    obs = ref.reset()
    score = np.zeros(N, dtype=np.float32)
    was_done = np.zeros(N, dtype=bool)
    h = net.init_hidden(N)
    n_explore = 0
    for t in range(T):
        c, l = divmod(t, L)
        assert np.array_equal(st[c, :, l], obs), f"state mismatch at step {t}"
        assert np.array_equal(alive[t].astype(bool), ~was_done)
        a = act[c, :, l].astype(np.int32)
        with torch.no_grad():
            q, h = net(out["states"][c, :, l], h)
        bv = q.gather(2, torch.from_numpy(a).to(cuda).long().unsqueeze(-1)).squeeze(-1).sum(1)
        assert (bv - behv[t]).abs().max().item() < 3e-5
        greedy = q.argmax(2).cpu().numpy()
        top2 = q.topk(2, dim=2)[0]
        clear = ((top2[..., 0] - top2[..., 1]) > 1e-4).all(1).cpu().numpy()       # rows whose argmax is not a near-tie
        explored = (a != greedy).any(1) & clear
        n_explore += int(explored.sum())
        obs, r, d = ref.step(a)
        assert np.array_equal(ns[c, :, l], obs), f"next_state mismatch at step {t}"
        assert np.array_equal(rew[c, :, l], r)
        assert np.array_equal(dn[c, :, l] != 0, d.all(1))
        score += np.where(was_done, 0.0, r.sum(1)).astype(np.float32)
        was_done |= d.all(1)
    assert np.allclose(out["score"].cpu().numpy(), score, atol=1e-4)
    # exploration: one draw per env -> a row that explores draws both agents at random; P(differs from greedy | explore) = 24/25
    frac = n_explore / (N * T)
    assert abs(frac - eps * 24 / 25) < 0.05, frac
    # eps = 0 with the same network: every clear row is greedy
    out0 = tf.rollout_episode(net, env, L, L, 0.0, seed=1)
    with torch.no_grad():
        q0, _ = net(out0["states"][0, :, 0], net.init_hidden(N))
    assert torch.equal(out0["actions"][0, :, 0], q0.argmax(2).float()) or \
        ((q0.topk(2, dim=2)[0][..., 0] - q0.topk(2, dim=2)[0][..., 1]).min().item() < 1e-4)
    # forced actions are taken verbatim
    fa = torch.randint(0, 5, (L, N, 2), device=cuda, dtype=torch.int32)
    outf = tf.rollout_episode(net, env, L, L, 0.7, seed=1, forced_actions=fa)
    assert torch.equal(outf["actions"][0].permute(1, 0, 2), fa.float())


def test_td_chunks_match_cal_td_error(cuda):
    from minimarl_b200.algorithms import cal_td_error
    from minimarl_b200.algorithms import toplevel_fused as tf
    Cn, N, L, NA = 3, 50, 10, 5
    g = torch.Generator(device=cuda).manual_seed(0)
    q_tgt = torch.randn((Cn, N, L, 2, NA), device=cuda, generator=g)
    rewards = torch.randn((Cn, N, L, 2), device=cuda, generator=g)
    dones = (torch.rand((Cn, N, L), device=cuda, generator=g) < 0.1).float()
    behv = torch.randn((Cn * L, N), device=cuda, generator=g)
    alive = (torch.rand((Cn * L, N), device=cuda, generator=g) < 0.8).to(torch.uint8)
    td, valid = tf.td_chunks(q_tgt, rewards, dones, behv, alive, 0.99)
    for c in range(Cn):
        want = torch.zeros(N, device=cuda)
        for l in range(L):
            t = c * L + l
            step = (rewards[c, :, l].sum(-1) + (1 - dones[c, :, l]) * 0.99 * q_tgt[c, :, l].max(2)[0].sum(1) - behv[t]).abs()
            want += step * alive[t]
        assert (td[c] - want).abs().max().item() < 1e-4
        assert torch.equal(valid[c], alive[c * L])


def _filled_per(cuda, args, D, n, seed=0):
    from minimarl_b200.algorithms import Prioritized_Experience_Replay
    per = Prioritized_Experience_Replay(args, n_agents=2, obs_dim=D, device=cuda)
    g = torch.Generator(device=cuda).manual_seed(seed)
    L = args.chunk_size
    s = torch.rand((n, L, 2, D), device=cuda, generator=g)
    a = torch.randint(0, 5, (n, L, 2), device=cuda, generator=g).float()
    r = torch.randn((n, L, 2), device=cuda, generator=g)
    s2 = torch.rand((n, L, 2, D), device=cuda, generator=g)
    d = (torch.rand((n, L), device=cuda, generator=g) < 0.1).float()
    td = torch.rand(n, device=cuda, generator=g) * 3
    per.collect_samples_batch([s, a, r, s2, d], td, warm_up=True)
    return per


@pytest.mark.parametrize("mixer", [False, True])
def test_fused_training_iteration_matches_module_iteration(cuda, mixer):
    """Same PER contents, same uniforms, same initial weights: the fused iteration and the autograd iteration give the same
    loss, priorities and updated weights."""
    from minimarl_b200.algorithms import Mix_Net, Q_Net, Target_Dqn, Train_dqn
    torch.backends.cuda.matmul.allow_tf32 = False
    D = 3 if mixer else 94
    results = []
    from minimarl_b200.optim import FlatAdam
    for fused, flat in ((False, False), (True, False), (True, True)):       # flat: FlatAdam (clip + Adam as two launches)
        adam = FlatAdam if flat else torch.optim.Adam
        args = _args(fused=fused, batch_size=64)
        per = _filled_per(cuda, args, D, 200)
        beh, tgt = _qnet(cuda, D, args, seed=1), _qnet(cuda, D, args, seed=2)
        u = torch.rand((1, 64), dtype=torch.float64, generator=torch.Generator().manual_seed(9)).numpy()
        if mixer:
            torch.manual_seed(3)
            mix = Mix_Net(beh.observation_space, args).to(cuda)
            tmix = Mix_Net(beh.observation_space, args).to(cuda)
            opt = adam(list(beh.parameters()) + list(mix.parameters()), lr=1e-3)
            loss = Train_dqn(args, cuda).train(per, beh, mix, tgt, tmix, opt, 0.5, uniforms=u)
            weights = [p.detach().clone() for p in list(beh.parameters()) + list(mix.parameters())]
        else:
            opt = adam(list(beh.parameters()), lr=1e-3)
            loss = Target_Dqn(per, beh, tgt, args, cuda).train(target_network=tgt, optimizer=opt, epsilon=0.5, uniforms=u)
            weights = [p.detach().clone() for p in beh.parameters()]
        results.append((float(loss), per.sum_tree.priority_tree.clone(), weights))
    l0, t0, w0 = results[0]
    for l1, t1, w1 in results[1:]:
        assert abs(l0 - l1) <= 1e-5 * abs(l0), (l0, l1)
        assert torch.allclose(t0, t1, rtol=1e-4, atol=1e-9)
        for a, b in zip(w0, w1):
            # one Adam step moves every weight by ~lr; a sign flip of a ~0 gradient would show as 2e-3
            assert (a - b).abs().max().item() < 2e-5, (a - b).abs().max().item()


def test_add_set_and_set_leaves_contract(cuda):
    from minimarl_b200.algorithms.toplevel import SumTree
    cap = 300
    tree = SumTree(cap, cuda)
    rng = np.random.default_rng(0)
    leaves = np.zeros(cap)
    n_data = 0
    for k, p_valid in ((120, 1.0), (250, 0.7), (90, 0.5), (400, 0.9)):
        prio = rng.random(k) * 4 + 1e-3
        valid = rng.random(k) < p_valid
        slot = tree.add_set(torch.from_numpy(prio).to(cuda), torch.from_numpy(valid.astype(np.uint8)).to(cuda)).cpu().numpy()
        items = np.flatnonzero(valid)
        fill = min(len(items), cap - n_data)
        want = np.full(k, -1, dtype=np.int64)
        want[items[:fill]] = n_data + np.arange(fill)
        rest = items[fill:]
        order = np.lexsort((np.arange(n_data), leaves[:n_data]))          # by priority, ties by slot
        n_ev = min(len(rest), n_data)
        want[rest[:n_ev]] = order[:n_ev]
        assert np.array_equal(slot, want)
        for i in np.flatnonzero(want >= 0):
            leaves[want[i]] = prio[i]
        n_data += fill
        assert tree.sync_n_data() == n_data
        got = tree.priority_tree.cpu().numpy()
        assert np.array_equal(got[cap - 1:], leaves)
        assert abs(got[0] - leaves.sum()) < 1e-9 * leaves.sum()
        for i in range(cap - 1):
            assert got[i] == got[2 * i + 1] + got[2 * i + 2]
    # write-back: last write wins, out-of-range nodes ignored
    idx = rng.integers(cap - 1, 2 * cap - 1, size=500)
    idx[7] = 3                                                       # an inner node: ignored
    pr = rng.random(500) + 0.01
    tree.set_leaves(torch.from_numpy(idx).to(cuda), torch.from_numpy(pr).to(cuda))
    for i in range(500):
        if idx[i] >= cap - 1:
            leaves[idx[i] - (cap - 1)] = pr[i]
    got = tree.priority_tree.cpu().numpy()
    assert np.array_equal(got[cap - 1:], leaves)
    for i in range(cap - 1):
        assert got[i] == got[2 * i + 1] + got[2 * i + 2]


@pytest.mark.parametrize("which", ["vdn", "qmix"])
def test_fused_runner_loops_run_and_store_only_live_chunks(cuda, which):
    from minimarl_b200.runners import run_qmix, run_vdn
    args = _args(buffer_limit=512, batch_size=64, update_iter=2, max_step=30, step_cost=-0.01, max_epsilon=0.9, min_epsilon=0.05,
                 epsilon_anneal_episode=100, update_target_interval=2, train_graph=(which == "qmix"))
    timing = {}
    run = run_vdn if which == "vdn" else run_qmix
    hist, net = run(args, 64, 4, device=cuda, timing=timing)
    assert timing["fused"] and len(hist) == 4
    assert all(np.isfinite(h["train_loss"]) for h in hist), hist
    assert timing["chunks_in_buffer"] == 512 and timing.get("warm_up_env_steps", 0) > 0
