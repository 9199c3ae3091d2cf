#!/usr/bin/env python
"""Generates tests/golden/*.npz by running the UNMODIFIED reference modules from /root/reference.

Run in the build container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference imports `gym` (not installed here) for space classes only; a ~20-line stub stands in.
Every fixture stores the seeded inputs, the network weights (state_dict) and the reference outputs, so
the tests can rebuild the same computation through the oracle and through the CUDA path.
"""
import argparse
import importlib
import os
import re
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


# ---------------------------------------------------------------------------------------------------
def install_gym_stub():
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")
    space_mod = types.ModuleType("gym.spaces.space")

    class Space:
        pass

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high = np.asarray(low, dtype=dtype), np.asarray(high, dtype=dtype)
            self.shape = self.low.shape if shape is None else shape
            self.dtype = dtype

    class Discrete(Space):
        def __init__(self, n):
            self.n, self.shape = n, ()

    class MultiDiscrete(Space):
        pass

    class Tuple(Space):
        pass

    space_mod.Space = Space
    spaces.Box, spaces.Discrete, spaces.Tuple, spaces.MultiDiscrete, spaces.Space = Box, Discrete, Tuple, MultiDiscrete, Space
    spaces.space = space_mod
    gym.spaces, gym.Space = spaces, Space
    sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.spaces.space": space_mod})
    return Box, Discrete


class ref_package:
    """Puts one reference package dir first on sys.path (its imports are CWD-relative) and drops the
    modules it loaded on exit so the next package's `utils`, `algorithms`... do not collide."""

    def __init__(self, name):
        self.dir = os.path.join(REF, name)

    def __enter__(self):
        self.before = set(sys.modules)
        sys.path.insert(0, self.dir)
        return self

    def __exit__(self, *exc):
        sys.path.remove(self.dir)
        for m in set(sys.modules) - self.before:
            f = getattr(sys.modules[m], "__file__", None) or ""
            if f.startswith(REF):
                del sys.modules[m]


def sd_numpy(module, prefix=""):
    # .copy(): a CPU tensor and its numpy view share memory, and the optimiser later updates in place
    return {prefix + k: v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


# ---------------------------------------------------------------------------------------------------
def golden_gae():
    """SharedReplayBuffer.compute_returns (mappo/runner/shared/shared_buffer.py:131-157)."""
    Box, Discrete = install_gym_stub()
    with ref_package("mappo"):
        from runner.shared.shared_buffer import SharedReplayBuffer
        from utils.valuenorm import ValueNorm
        out = {}
        for tag, (T, B, A, use_gae, use_vn) in {
            "vn": (16, 8, 2, True, True), "plain": (16, 8, 2, True, False), "nogae": (16, 8, 2, False, False),
            "long": (100, 32, 2, True, True),
        }.items():
            rng = np.random.default_rng(sum(map(ord, tag)))
            args = argparse.Namespace(use_gae=use_gae, use_popart=False, use_valuenorm=use_vn, hidden_size=32,
                                      recurrent_N=1, max_step=T, batch_size=B, gamma=0.99, gae_lambda=0.95)
            space = Box(np.zeros(47), np.ones(47))
            buf = SharedReplayBuffer(args, A, space, space, Discrete(5))
            buf.rewards[:] = rng.standard_normal(buf.rewards.shape).astype(np.float32)
            buf.value_preds[:] = rng.standard_normal(buf.value_preds.shape).astype(np.float32)
            buf.masks[:] = (rng.random(buf.masks.shape) < 0.9).astype(np.float32)
            next_value = rng.standard_normal((B, A, 1)).astype(np.float32)
            vn = None
            scale = shift = np.float32(0)
            if use_vn:
                vn = ValueNorm(1)
                for _ in range(3):
                    vn.update(torch.from_numpy(rng.standard_normal((64, 1)).astype(np.float32) * 3 + 0.7))
                mean, var = vn.running_mean_var()
                scale, shift = torch.sqrt(var).numpy()[0], mean.numpy()[0]
            out[f"{tag}_rewards"] = buf.rewards.copy()
            out[f"{tag}_masks"] = buf.masks.copy()
            out[f"{tag}_next_value"] = next_value
            out[f"{tag}_value_preds_in"] = buf.value_preds.copy()
            buf.compute_returns(next_value, vn)
            out[f"{tag}_returns"] = buf.returns.copy()
            out[f"{tag}_denorm"] = np.array([scale, shift], dtype=np.float32)
            out[f"{tag}_flags"] = np.array([use_gae, use_vn], dtype=np.int32)
            if use_gae:
                adv = buf.returns[:-1] - (vn.denormalize(buf.value_preds[:-1]) if vn else buf.value_preds[:-1])
                out[f"{tag}_advantages"] = adv.astype(np.float32)
        np.savez_compressed(os.path.join(OUT, "gae.npz"), **out)
        print("gae.npz", len(out), "arrays")


def offpolicy_args(**over):
    from config import get_config
    args = get_config().parse_known_args([])[0]
    for k, v in over.items():
        setattr(args, k, v)
    return args


def golden_segment_tree_and_per():
    """SumSegmentTree / MinSegmentTree / PrioritizedRecReplayBuffer (offpolicy/utils/segment_tree.py,
    rec_buffer.py:243-324)."""
    Box, Discrete = install_gym_stub()
    with ref_package("offpolicy"):
        from utils.segment_tree import SumSegmentTree, MinSegmentTree
        from utils.rec_buffer import PrioritizedRecReplayBuffer
        rng = np.random.default_rng(7)
        out = {}
        cap = 64
        st, mt = SumSegmentTree(cap), MinSegmentTree(cap)
        pri = rng.random(50) + 0.01
        st[np.arange(50)] = pri
        mt[np.arange(50)] = pri
        # scattered update with duplicates of parents
        idx = np.array([3, 4, 17, 18, 19, 49])
        newp = rng.random(6) * 2 + 0.01
        st[idx] = newp
        mt[idx] = newp
        out["tree_values"] = st._value.copy()
        out["min_values"] = mt._value.copy()
        mass = rng.random(200) * st.sum(0, 49)
        out["tree_mass"] = mass
        out["tree_idx"] = st.find_prefixsum_idx(mass)
        out["tree_sum_0_49"] = np.array(st.sum(0, 49))
        out["tree_min"] = np.array(mt.min())

        # full PER buffer: insert episodes, update priorities, sample with recorded uniforms
        T, A, D, nA, size = 6, 2, 47, 5, 24
        obs_space, act_space = Box(np.zeros(D), np.ones(D)), Discrete(nA)
        share_space = Box(np.zeros(A * D), np.ones(A * D))
        info = {"policy_0": {"obs_space": obs_space, "share_obs_space": share_space, "act_space": act_space}}
        buf = PrioritizedRecReplayBuffer(0.6, info, {"policy_0": [0, 1]}, size, T, True, False)
        eps = []
        for it in range(5):
            n = 7
            ep = dict(
                obs=rng.random((T + 1, n, A, D)).astype(np.float32),
                share_obs=rng.random((T + 1, n, A, A * D)).astype(np.float32),
                acts=np.eye(nA, dtype=np.float32)[rng.integers(0, nA, (T, n, A))],
                rewards=rng.standard_normal((T, n, A, 1)).astype(np.float32),
                dones=(rng.random((T, n, A, 1)) < 0.1).astype(np.float32),
                dones_env=(rng.random((T, n, 1)) < 0.1).astype(np.float32))
            eps.append(ep)
            idx_range = buf.insert(n, {"policy_0": ep["obs"]}, {"policy_0": ep["share_obs"]}, {"policy_0": ep["acts"]},
                                   {"policy_0": ep["rewards"]}, {"policy_0": ep["dones"]}, {"policy_0": ep["dones_env"]})
            out[f"per_insert{it}_idx"] = idx_range
            for k, v in ep.items():
                out[f"per_insert{it}_{k}"] = v
            if it >= 1:
                upd_idx = rng.choice(len(buf), 5, replace=False)
                upd_p = rng.random(5) * 3 + 0.05
                buf.update_priorities(upd_idx, upd_p, "policy_0")
                out[f"per_update{it}_idx"], out[f"per_update{it}_p"] = upd_idx, upd_p
        out["per_sum_tree"] = buf._it_sums["policy_0"]._value.copy()
        out["per_min_tree"] = buf._it_mins["policy_0"]._value.copy()
        out["per_max_priority"] = np.array(buf.max_priorities["policy_0"])
        out["per_len"] = np.array(len(buf))
        u = rng.random(8)
        state = np.random.get_state()
        np.random.seed(0)
        orig = np.random.random
        np.random.random = lambda size=None: u      # record the uniforms the sampler consumes
        try:
            s = buf.sample(8, beta=0.4, p_id="policy_0")
        finally:
            np.random.random = orig
            np.random.set_state(state)
        out["per_sample_u"] = u
        names = ["obs", "share_obs", "acts", "rewards", "dones", "dones_env"]
        for k, v in zip(names, s[:6]):
            out[f"per_sample_{k}"] = v["policy_0"]
        out["per_sample_weights"], out["per_sample_idx"] = s[7], s[8]
        np.savez_compressed(os.path.join(OUT, "per.npz"), **out)
        print("per.npz", len(out), "arrays")


def golden_offpolicy_nets():
    """AgentQFunction / QMixer / VDNMixer forward and one QMix.train_policy_on_batch
    (offpolicy/algorithms/qmix/*)."""
    Box, Discrete = install_gym_stub()
    with ref_package("offpolicy"):
        torch.manual_seed(3)
        np.random.seed(3)
        from algorithms.qmix.algorithm.QMixPolicy import QMixPolicy
        from algorithms.qmix.qmix import QMix
        from algorithms.vdn.algorithm.vdn_mixer import VDNMixer
        T, B, A, D, nA = 12, 6, 2, 47, 5
        args = offpolicy_args(episode_length=T, batch_size=B, use_per=True, algorithm_name="qmix")
        dev = torch.device("cpu")
        obs_space, act_space = Box(np.zeros(D), np.ones(D)), Discrete(nA)
        pconf = {"cent_obs_dim": A * D, "cent_act_dim": A * nA, "obs_space": obs_space,
                 "share_obs_space": Box(np.zeros(A * D), np.ones(A * D)), "act_space": act_space}
        config = {"args": args, "device": dev, "num_agents": A}
        policy = QMixPolicy(args, config, pconf)
        policies = {"policy_0": policy}
        trainer = QMix(args, A, B, policies, lambda a: "policy_0", device=dev, episode_length=T)
        # perturb LayerNorm affine / biases so that they are exercised (they initialise to 1 / 0)
        with torch.no_grad():
            for m in [policy.q_network, trainer.mixer]:
                for name, p in m.named_parameters():
                    if p.dim() == 1:
                        p.add_(0.1 * torch.randn_like(p))
            trainer.target_policies["policy_0"].q_network.load_state_dict(policy.q_network.state_dict())
            for p in trainer.target_policies["policy_0"].q_network.parameters():
                p.add_(0.01 * torch.randn_like(p))
            trainer.target_mixer.load_state_dict(trainer.mixer.state_dict())
            for p in trainer.target_mixer.parameters():
                p.add_(0.01 * torch.randn_like(p))
        rng = np.random.default_rng(11)
        obs = rng.random((A, T + 1, B, D)).astype(np.float32)
        share_obs = rng.random((T + 1, B, A * D)).astype(np.float32)
        act_idx = rng.integers(0, nA, (A, T, B))
        acts = np.eye(nA, dtype=np.float32)[act_idx]
        rewards = rng.standard_normal((A, T, B, 1)).astype(np.float32)
        dones = np.zeros((A, T, B, 1), dtype=np.float32)
        dones_env = np.zeros((T, B, 1), dtype=np.float32)
        for b in range(B):
            end = rng.integers(T // 2, T + 1)
            dones_env[end - 1:, b] = 1.0
        weights = (rng.random(B) * 0.8 + 0.2).astype(np.float64)
        idxes = np.arange(B)
        out = dict(obs=obs, share_obs=share_obs, acts=acts, rewards=rewards, dones=dones, dones_env=dones_env,
                   weights=weights, gamma=np.array(args.gamma), per_nu=np.array(args.per_nu),
                   per_eps=np.array(args.per_eps), max_grad_norm=np.array(args.max_grad_norm))
        out.update(sd_numpy(policy.q_network, "q."))
        out.update(sd_numpy(trainer.target_policies["policy_0"].q_network, "tq."))
        out.update(sd_numpy(trainer.mixer, "mix."))
        out.update(sd_numpy(trainer.target_mixer, "tmix."))

        # forward pieces
        stacked_obs = torch.cat(list(torch.from_numpy(obs)), dim=-2)            # [T+1, A*B, D]
        with torch.no_grad():
            q_all, h_final = policy.q_network(stacked_obs, policy.init_hidden(-1, A * B))
            out["q_all"], out["h_final"] = q_all.numpy(), h_final.numpy()
            q1, h1 = policy.q_network(stacked_obs[0], torch.zeros(A * B, args.hidden_size))
            out["q_step0"], out["h_step0"] = q1.numpy(), h1.numpy()
            agent_q = torch.from_numpy(rng.standard_normal((T, B, A)).astype(np.float32))
            out["mixer_in_q"] = agent_q.numpy()
            out["mixer_out"] = trainer.mixer(agent_q, torch.from_numpy(share_obs[:-1])).numpy()
            vdn = VDNMixer(args, A, A * D, dev)
            out["vdn_out"] = vdn(agent_q, torch.from_numpy(share_obs[:-1])).numpy()

        batch = ({"policy_0": obs}, {"policy_0": share_obs}, {"policy_0": acts}, {"policy_0": rewards},
                 {"policy_0": dones}, {"policy_0": dones_env}, {"policy_0": None}, weights, idxes)
        # capture the gradients before the optimiser step
        grads = {}
        orig_clip = torch.nn.utils.clip_grad_norm_

        def spy(params, max_norm, *a, **k):
            params = list(params)
            for i, p in enumerate(params):
                grads[i] = None if p.grad is None else p.grad.detach().clone().numpy()
            return orig_clip(params, max_norm, *a, **k)
        torch.nn.utils.clip_grad_norm_ = spy
        try:
            info, new_prio, _ = trainer.train_policy_on_batch(batch)
        finally:
            torch.nn.utils.clip_grad_norm_ = orig_clip
        out["train_loss"] = info["loss"].detach().numpy()
        out["train_grad_norm"] = np.array(float(info["grad_norm"]))
        out["train_Q_tot"] = info["Q_tot"].detach().numpy()
        out["train_new_priorities"] = new_prio
        names = [n for n, _ in policy.q_network.named_parameters()] + ["mix." + n for n, _ in trainer.mixer.named_parameters()]
        for i, n in enumerate(names):
            if grads[i] is not None:
                out["grad." + n] = grads[i]
        np.savez_compressed(os.path.join(OUT, "offpolicy_qmix.npz"), **out)
        print("offpolicy_qmix.npz", len(out), "arrays; loss", float(out["train_loss"]))


def golden_log_scores():
    """Episode scores logged by the author's VDN runs: the only numeric evidence about the env."""
    out = {}
    for name in ["vdn-1710766189.log", "vdn-1710733227.log", "double_dqn-1710514686.log"]:
        sc = []
        for line in open(os.path.join(REF, "vdn", "logs", name)):
            m = re.search(r"episodes \| train score: (-?[\d.]+)", line)
            if m:
                sc.append(float(m.group(1)))
        out[name.split(".")[0].replace("-", "_")] = np.round(np.array(sc) * 100).astype(np.int32)   # centi-points
    np.savez_compressed(os.path.join(OUT, "vdn_log_scores.npz"), **out)
    print("vdn_log_scores.npz", {k: len(v) for k, v in out.items()})


def golden_mappo_generators():
    """SharedReplayBuffer.insert / after_update / recurrent_generator / feed_forward_generator
    (mappo/runner/shared/shared_buffer.py:82-129,159-219,318-427) with a recorded permutation."""
    Box, Discrete = install_gym_stub()
    with ref_package("mappo"):
        from runner.shared.shared_buffer import SharedReplayBuffer
        T, B, A, D, H = 10, 6, 2, 47, 32
        rng = np.random.default_rng(21)
        args = argparse.Namespace(use_gae=True, use_popart=False, use_valuenorm=False, hidden_size=H,
                                  recurrent_N=1, max_step=T, batch_size=B, gamma=0.99, gae_lambda=0.95)
        space = Box(np.zeros(D), np.ones(D))
        buf = SharedReplayBuffer(args, A, space, space, Discrete(5))
        out = {}
        f = lambda *shape: rng.standard_normal(shape).astype(np.float32)
        buf.obs[0] = f(B, A, D); buf.share_obs[0] = buf.obs[0]
        out["obs0"] = buf.obs[0].copy()
        for t in range(T):
            step = dict(share_obs=f(B, A, D), obs=f(B, A, D), rnn_states_actor=f(B, A, 1, H),
                        rnn_states_critic=f(B, A, 1, H), actions=rng.integers(0, 5, (B, A, 1)).astype(np.float32),
                        action_log_probs=f(B, A, 1), value_preds=f(B, A, 1), rewards=f(B, A, 1),
                        masks=(rng.random((B, A, 1)) < 0.9).astype(np.float32))
            buf.insert(**step)
            for k, v in step.items():
                out[f"step{t}_{k}"] = v
        next_value = f(B, A, 1)
        out["next_value"] = next_value
        buf.compute_returns(next_value, None)
        out["returns"] = buf.returns.copy()
        adv = f(T, B, A, 1)
        out["advantages"] = adv
        names = ["share_obs", "obs", "rnn_states", "rnn_states_critic", "actions", "value_preds", "returns",
                 "masks", "active_masks", "old_action_log_probs", "adv_targ"]
        orig = torch.randperm
        for gen_name, kw, n_items in [("recurrent", dict(num_mini_batch=2, data_chunk_length=5), (T * B * A) // 5),
                                      ("feed_forward", dict(num_mini_batch=3), T * B * A),
                                      ("naive", dict(num_mini_batch=2), B * A)]:
            perm = torch.from_numpy(rng.permutation(n_items))
            out[f"{gen_name}_perm"] = perm.numpy()
            torch.randperm = lambda n, *a, **k: perm
            try:
                gen = {"recurrent": buf.recurrent_generator, "feed_forward": buf.feed_forward_generator,
                       "naive": buf.naive_recurrent_generator}[gen_name]
                for mb, sample in enumerate(gen(advantages=adv, **kw)):
                    for name, x in zip(names, sample[:11]):
                        out[f"{gen_name}_mb{mb}_{name}"] = np.asarray(x)
            finally:
                torch.randperm = orig
        buf.after_update()
        out["after_update_obs0"] = buf.obs[0].copy()
        out["after_update_masks0"] = buf.masks[0].copy()
        np.savez_compressed(os.path.join(OUT, "mappo_buffer.npz"), **out)
        print("mappo_buffer.npz", len(out), "arrays")


def golden_mappo_train():
    """R_MAPPOPolicy forward pieces and one R_MAPPO.train (mappo/algorithms/*), recorded perms."""
    Box, Discrete = install_gym_stub()
    with ref_package("mappo"):
        from _config import get_config
        from algorithms.rmappo_policy import R_MAPPOPolicy
        from algorithms.ramppo_network import R_MAPPO
        from runner.shared.shared_buffer import SharedReplayBuffer
        torch.manual_seed(5)
        np.random.seed(5)
        args = get_config().parse_known_args([])[0]
        T, B, A, D, H = 10, 6, 2, 47, 32
        args.max_step, args.batch_size, args.ppo_epoch, args.train_batch_size, args.data_chunk_length = T, B, 2, 2, 5
        args.use_cuda = False
        dev = torch.device("cpu")
        space = Box(np.zeros(D), np.ones(D))
        policy = R_MAPPOPolicy(args, space, space, Discrete(5), device=dev)
        trainer = R_MAPPO(args, policy, device=dev)
        with torch.no_grad():
            for m in (policy.actor, policy.critic):
                for p_ in m.parameters():
                    if p_.dim() == 1:
                        p_.add_(0.1 * torch.randn_like(p_))
        out = {k: np.array(getattr(args, k)) for k in ["hidden_size", "ppo_epoch", "train_batch_size", "data_chunk_length",
               "clip_param", "entropy_coef", "value_loss_coef", "max_grad_norm", "huber_delta", "actor_lr", "critic_lr",
               "opti_eps", "gamma", "gae_lambda"]}
        out.update(sd_numpy(policy.actor, "actor."))
        out.update(sd_numpy(policy.critic, "critic."))
        rng = np.random.default_rng(31)
        f = lambda *shape: rng.standard_normal(shape).astype(np.float32)
        buf = SharedReplayBuffer(args, A, space, space, Discrete(5))
        buf.obs[:] = rng.random(buf.obs.shape).astype(np.float32)
        buf.share_obs[:] = buf.obs
        buf.rnn_states[:] = 0.3 * f(*buf.rnn_states.shape)
        buf.rnn_states_critic[:] = 0.3 * f(*buf.rnn_states_critic.shape)
        buf.actions[:] = rng.integers(0, 5, buf.actions.shape).astype(np.float32)
        buf.masks[:] = (rng.random(buf.masks.shape) < 0.85).astype(np.float32)
        buf.rewards[:] = f(*buf.rewards.shape)
        # one-step forward pieces (rollout path): values, log-probs of given actions, new rnn states
        with torch.no_grad():
            obs0 = np.concatenate(buf.obs[0]); h_a = np.concatenate(buf.rnn_states[0]); h_c = np.concatenate(buf.rnn_states_critic[0])
            m0 = np.concatenate(buf.masks[0])
            values, h_c1 = policy.critic(obs0, h_c, m0)
            a_feat = policy.actor.base(torch.from_numpy(obs0))
            a_feat, h_a1 = policy.actor.rnn(a_feat, torch.from_numpy(h_a), torch.from_numpy(m0))
            logits = policy.actor.act.action_out.linear(a_feat)
            det_actions, h_a2 = policy.act(obs0, h_a, m0, deterministic=True)
            out["step_obs"], out["step_h_actor"], out["step_h_critic"], out["step_masks"] = obs0, h_a, h_c, m0
            out["step_values"], out["step_h_critic_out"] = values.numpy(), h_c1.numpy()
            out["step_logits"], out["step_h_actor_out"] = logits.numpy(), h_a1.numpy()
            out["step_det_actions"] = det_actions.numpy()
            # fill value_preds / log-probs the way a rollout would
            for t in range(T):
                o = np.concatenate(buf.obs[t]); m = np.concatenate(buf.masks[t])
                v, _ = policy.critic(o, np.concatenate(buf.rnn_states_critic[t]), m)
                lp, _ = policy.actor.evaluate_actions(o, np.concatenate(buf.rnn_states[t]), np.concatenate(buf.actions[t]), m)
                buf.value_preds[t] = v.numpy().reshape(B, A, 1)
                buf.action_log_probs[t] = lp.numpy().reshape(B, A, 1) + 0.05 * f(B, A, 1)
            next_value = policy.get_values(np.concatenate(buf.obs[-1]), np.concatenate(buf.rnn_states_critic[-1]),
                                           np.concatenate(buf.masks[-1])).numpy().reshape(B, A, 1)
        buf.compute_returns(next_value, trainer.value_normalizer)
        for k in ["obs", "share_obs", "rnn_states", "rnn_states_critic", "actions", "masks", "rewards", "value_preds",
                  "action_log_probs", "returns", "active_masks"]:
            out["buf_" + k] = getattr(buf, k).copy()
        out["next_value"] = next_value
        n_chunks = (T * B * A) // args.data_chunk_length
        perms = [torch.from_numpy(rng.permutation(n_chunks)) for _ in range(args.ppo_epoch)]
        out["perms"] = np.stack([p_.numpy() for p_ in perms])
        it = iter(perms)
        orig = torch.randperm
        torch.randperm = lambda n, *a, **k: next(it)
        try:
            info = trainer.train(buf)
        finally:
            torch.randperm = orig
        for k, v in info.items():
            out["info_" + k] = np.array(float(v))
        out.update(sd_numpy(policy.actor, "actor_after."))
        out.update(sd_numpy(policy.critic, "critic_after."))
        out.update(sd_numpy(trainer.value_normalizer, "vn_after."))
        np.savez_compressed(os.path.join(OUT, "mappo_train.npz"), **out)
        print("mappo_train.npz", len(out), "arrays;", {k: float(v) for k, v in info.items()})


def golden_vdn_toplevel():
    """vdn/ top level: Q_Net forward, chunk PER (SumTree add / evict / sample / decay / update), Target_Dqn.train,
    and qmix/ Mix_Net forward.  random.uniform is patched to consume recorded uniforms."""
    import random as pyrandom
    Box, Discrete = install_gym_stub()
    out = {}
    with ref_package("vdn"):
        from _config import get_config
        from _network import Q_Net
        from _train import Target_Dqn
        from _utils import cal_td_error
        from replay_buffer.buffer import Prioritized_Experience_Replay
        torch.manual_seed(9)
        args = get_config().parse_known_args([])[0]
        args.buffer_limit, args.batch_size, args.update_iter, args.chunk_size, args.use_cuda = 50, 8, 2, 10, False
        A, D, L = 2, 94, args.chunk_size
        obs_space = [Box(np.zeros(D), np.ones(D)) for _ in range(A)]
        act_space = [Discrete(5) for _ in range(A)]
        beh, tgt = Q_Net(obs_space, act_space, args), Q_Net(obs_space, act_space, args)
        out.update(sd_numpy(beh, "beh."))
        out.update(sd_numpy(tgt, "tgt."))
        for k in ["gamma", "eps", "alpha", "beta", "step_weight", "grad_clip_norm", "lr", "max_episodes", "update_iter"]:
            out["arg_" + k] = np.array(getattr(args, k))
        rng = np.random.default_rng(17)
        obs = rng.random((3, A, D)).astype(np.float32)
        hid = (0.3 * rng.standard_normal((3, A, 32))).astype(np.float32)
        with torch.no_grad():
            q, h = beh(torch.from_numpy(obs), torch.from_numpy(hid))
            tq, _ = tgt(torch.from_numpy(obs), torch.from_numpy(hid))
        out["fwd_obs"], out["fwd_hidden"], out["fwd_q"], out["fwd_h"] = obs, hid, q.numpy(), h.numpy()
        act = torch.tensor([1.0, 3.0])
        out["td_error"] = np.array(cal_td_error(action=act, reward=[-0.01, 9.99], done=0, behavior_q=q[:1], target_q=tq[:1], gamma=args.gamma))
        out["td_action"], out["td_target_q"] = act.numpy(), tq[:1].numpy()

        per = Prioritized_Experience_Replay(args)
        n_add = 70                                           # > capacity: exercises the arg-min eviction
        chunks = dict(states=rng.random((n_add, L, A, D)).astype(np.float32), actions=rng.integers(0, 5, (n_add, L, 1, A)).astype(np.float32),
                      rewards=rng.standard_normal((n_add, L, A)).astype(np.float32), next_states=rng.random((n_add, L, A, D)).astype(np.float32),
                      dones=(rng.random((n_add, L)) < 0.05).astype(np.float32), td=(rng.random(n_add) * 5).astype(np.float64))
        for k, v in chunks.items():
            out["chunk_" + k] = v
        for i in range(n_add):
            sample = [chunks["states"][i].tolist(), chunks["actions"][i].tolist(), chunks["rewards"][i].tolist(),
                      chunks["next_states"][i].tolist(), chunks["dones"][i].tolist()]
            per.collect_sample(sample=sample, td_error=float(chunks["td"][i]), warm_up=True)
        out["tree_after_add"] = per.sum_tree.priority_tree.copy()
        u = rng.random((args.update_iter + 1, args.batch_size))
        out["uniforms"] = u
        seq = iter(u.reshape(-1))
        orig = pyrandom.uniform
        pyrandom.uniform = lambda a, b: a + (b - a) * next(seq)
        try:
            s = per.sample(args.batch_size, L)
            out["sample_states"], out["sample_actions"], out["sample_rewards"] = s[0].numpy(), s[1].numpy(), s[2].numpy()
            out["sample_next_states"], out["sample_dones"] = s[3].numpy(), s[4].numpy()
            out["sample_idx"], out["sample_w"] = np.array(s[5]), s[6].numpy()
            out["tree_after_sample"] = per.sum_tree.priority_tree.copy()
            out["alpha_after_sample"], out["beta_after_sample"] = np.array(per.alpha), np.array(per.beta)
            trainer = Target_Dqn(per, beh, tgt, args, torch.device("cpu"))
            opt = torch.optim.Adam(beh.parameters(), lr=args.lr)
            loss = trainer.train(target_network=tgt, optimizer=opt, epsilon=0.5)
        finally:
            pyrandom.uniform = orig
        out["train_loss"] = np.array(float(loss))
        out["tree_after_train"] = per.sum_tree.priority_tree.copy()
        out.update(sd_numpy(beh, "beh_after."))
    with ref_package("qmix"):
        from _config import get_config as qcfg
        from _network import Mix_Net
        qargs = qcfg().parse_known_args([])[0]
        torch.manual_seed(10)
        mix = Mix_Net([Box(np.zeros(3), np.ones(3)) for _ in range(2)], qargs)
        out.update(sd_numpy(mix, "mix."))
        rng = np.random.default_rng(4)
        qv = rng.standard_normal((5, 2)).astype(np.float32)
        ob = rng.random((5, 2, 3)).astype(np.float32)
        hd = (0.2 * rng.standard_normal((5, 32))).astype(np.float32)
        with torch.no_grad():
            y, h = mix(torch.from_numpy(qv), torch.from_numpy(ob), torch.from_numpy(hd))
        out["mix_q"], out["mix_obs"], out["mix_hidden"], out["mix_out"], out["mix_h"] = qv, ob, hd, y.numpy(), h.numpy()
    np.savez_compressed(os.path.join(OUT, "toplevel.npz"), **out)
    print("toplevel.npz", len(out), "arrays; vdn train loss", float(out["train_loss"]))


def golden_qmix_toplevel():
    """qmix/ top level: chunk PER with the whole-tree-argsort eviction of qmix/replay_buffer/sumtree.py,
    then ONE Train_dqn.train call (update_iter 2: Q_Net + recurrent Mix_Net on both sides, done resets, q-net-only
    clipping, last-step priorities) on the qmix default env shape (Switch2: A=2, D=3).  batch_size must be 32: the
    reference's mix-hidden reset indexes a [32] row with the [B] done mask (qmix/_train.py:97-99).
    random.uniform is patched to consume recorded uniforms; F.mse_loss is wrapped to record the per-step losses."""
    import random as pyrandom
    Box, Discrete = install_gym_stub()
    out = {}
    with ref_package("qmix"):
        from _config import get_config
        from _network import Mix_Net, Q_Net
        from _train import Train_dqn
        import _train as train_mod
        from replay_buffer.per import Prioritized_Experience_Replay
        torch.manual_seed(21)
        args = get_config().parse_known_args([])[0]
        args.buffer_limit, args.batch_size, args.update_iter, args.chunk_size, args.use_recurrent = 64, 32, 2, 10, True
        A, D, L = 2, 3, args.chunk_size
        obs_space = [Box(np.zeros(D), np.ones(D)) for _ in range(A)]
        act_space = [Discrete(5) for _ in range(A)]
        q, tq = Q_Net(obs_space, act_space, args), Q_Net(obs_space, act_space, args)
        mix, tmix = Mix_Net(obs_space, args), Mix_Net(obs_space, args)
        for name, m in [("q.", q), ("tq.", tq), ("mixnet.", mix), ("tmixnet.", tmix)]:
            out.update(sd_numpy(m, name))
        for k in ["gamma", "eps", "alpha", "beta", "grad_clip_norm", "lr", "max_episodes", "update_iter", "update_alpha_beta"]:
            out["arg_" + k] = np.array(getattr(args, k))
        rng = np.random.default_rng(23)
        per = Prioritized_Experience_Replay(args)
        n_add = 80
        chunks = dict(states=rng.random((n_add, L, A, D)).astype(np.float32), actions=rng.integers(0, 5, (n_add, L, 1, A)).astype(np.float32),
                      rewards=rng.standard_normal((n_add, L, A)).astype(np.float32), next_states=rng.random((n_add, L, A, D)).astype(np.float32),
                      dones=(rng.random((n_add, L)) < 0.08).astype(np.float32), td=(rng.random(n_add) * 5 + 0.01).astype(np.float64))
        for k, v in chunks.items():
            out["chunk_" + k] = v
        for i in range(n_add):
            sample = [chunks["states"][i].tolist(), chunks["actions"][i].tolist(), chunks["rewards"][i].tolist(),
                      chunks["next_states"][i].tolist(), chunks["dones"][i].tolist()]
            per.collect_sample(sample=sample, td_error=float(chunks["td"][i]), warm_up=True)
        out["tree_after_add"] = per.sum_tree.priority_tree.copy()
        u = rng.random((args.update_iter, args.batch_size))
        out["uniforms"] = u
        seq = iter(u.reshape(-1))
        orig_uniform, orig_mse = pyrandom.uniform, train_mod.F.mse_loss
        step_losses = []

        def mse(a, b):
            v = orig_mse(a, b)
            step_losses.append(float(v))
            return v
        pyrandom.uniform = lambda a, b: a + (b - a) * next(seq)
        train_mod.F.mse_loss = mse
        try:
            opt = torch.optim.Adam(list(q.parameters()) + list(mix.parameters()), lr=args.lr)
            Train_dqn(args, torch.device("cpu")).train(per, q, mix, tq, tmix, opt, 0.5)
        finally:
            pyrandom.uniform, train_mod.F.mse_loss = orig_uniform, orig_mse
        out["step_losses"] = np.array(step_losses).reshape(args.update_iter, L)
        out["tree_after_train"] = np.asarray(per.sum_tree.priority_tree, dtype=np.float64).copy()
        out["alpha_after"], out["beta_after"] = np.array(per.alpha), np.array(per.beta)
        out.update(sd_numpy(q, "q_after."))
        out.update(sd_numpy(mix, "mixnet_after."))
    np.savez_compressed(os.path.join(OUT, "qmix_toplevel.npz"), **out)
    print("qmix_toplevel.npz", len(out), "arrays; iteration losses", out["step_losses"].sum(1))


def golden_vdn_test():
    """vdn/_test.py Test.execute (the eval adaptor) run by the reference itself on the oracle env:
    greedy episodes of reference Q_Nets with recorded weights -> (score, TD loss)."""
    Box, Discrete = install_gym_stub()
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from oracle.ma_gym_ref import CheckersRef
    out = {}
    with ref_package("vdn"):
        from _config import get_config
        from _network import Q_Net
        from _test import Test
        args = get_config().parse_known_args([])[0]
        args.use_cuda, args.test_episodes = False, 2
        for seed, max_step in ((21, 40), (22, 100)):
            torch.manual_seed(seed)
            env = CheckersRef(full_observable=True, max_steps=max_step, step_cost=args.step_cost)
            obs_space = [Box(np.zeros(94), np.ones(94)) for _ in range(2)]
            act_space = [Discrete(5) for _ in range(2)]
            beh, tgt = Q_Net(obs_space, act_space, args), Q_Net(obs_space, act_space, args)
            beh.device = tgt.device = torch.device("cpu")
            score, loss = Test(env, args, torch.device("cpu")).execute(behavior_network=beh, target_network=tgt)
            tag = f"s{seed}_"
            out.update(sd_numpy(beh, tag + "beh."))
            out.update(sd_numpy(tgt, tag + "tgt."))
            out[tag + "score"], out[tag + "loss"], out[tag + "max_step"] = np.array(float(score)), np.array(float(loss)), np.array(max_step)
        out["gamma"], out["step_cost"], out["test_episodes"] = np.array(args.gamma), np.array(args.step_cost), np.array(2)
        out["use_recurrent"] = np.array(bool(args.use_recurrent))
    np.savez_compressed(os.path.join(OUT, "vdn_test.npz"), **out)
    print("vdn_test.npz", {k: float(out[k]) for k in out if k.endswith(("score", "loss"))})


if __name__ == "__main__":
    which = sys.argv[1:] or ["gae", "per", "nets", "logs", "mappo_buffer", "mappo_train", "toplevel", "qmix_toplevel", "vdn_test"]
    if "gae" in which:
        golden_gae()
    if "per" in which:
        golden_segment_tree_and_per()
    if "nets" in which:
        golden_offpolicy_nets()
    if "logs" in which:
        golden_log_scores()
    if "mappo_buffer" in which:
        golden_mappo_generators()
    if "mappo_train" in which:
        golden_mappo_train()
    if "toplevel" in which:
        golden_vdn_toplevel()
    if "qmix_toplevel" in which:
        golden_qmix_toplevel()
    if "vdn_test" in which:
        golden_vdn_test()
