"""GPU parity: device SharedReplayBuffer vs fixtures recorded from the reference's own class
(tests/golden/mappo_buffer.npz, gae.npz): insert / compute_returns / generators / after_update."""
import argparse
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from minimarl_b200.envs.spaces import Box, Discrete

pytestmark = pytest.mark.gpu
NAMES = ["share_obs", "obs", "rnn_states", "rnn_states_critic", "actions", "value_preds", "returns", "masks",
         "active_masks", "old_action_log_probs", "adv_targ"]


def _buffer(cuda, T, B, use_gae=True, use_vn=False):
    from minimarl_b200.buffers import SharedReplayBuffer
    args = argparse.Namespace(use_gae=use_gae, use_popart=False, use_valuenorm=use_vn, hidden_size=32, recurrent_N=1,
                              max_step=T, batch_size=B, gamma=0.99, gae_lambda=0.95)
    space = Box(np.zeros(47), np.ones(47))
    return SharedReplayBuffer(args, 2, space, space, Discrete(5), device=cuda)


def test_insert_returns_generators_after_update(cuda):
    g = np.load(os.path.join(GOLDEN, "mappo_buffer.npz"))
    T, B = 10, 6
    buf = _buffer(cuda, T, B)
    buf.obs[0].copy_(torch.from_numpy(g["obs0"]))
    buf.share_obs[0].copy_(torch.from_numpy(g["obs0"]))
    keys = ["share_obs", "obs", "rnn_states_actor", "rnn_states_critic", "actions", "action_log_probs", "value_preds",
            "rewards", "masks"]
    for t in range(T):
        buf.insert(**{k: g[f"step{t}_{k}"] for k in keys})
    buf.compute_returns(g["next_value"], None)
    assert np.array_equal(buf.returns[:-1].cpu().numpy(), g["returns"][:-1])
    adv = torch.from_numpy(g["advantages"]).to(cuda)
    gens = {"recurrent": lambda p: buf.recurrent_generator(adv, 2, 5, perm=p),
            "feed_forward": lambda p: buf.feed_forward_generator(adv, 3, perm=p),
            "naive": lambda p: buf.naive_recurrent_generator(adv, 2, perm=p)}
    for name, make in gens.items():
        n_mb = 0
        for mb, sample in enumerate(make(g[f"{name}_perm"])):
            assert sample[11] is None
            for field, x in zip(NAMES, sample[:11]):
                want = g[f"{name}_mb{mb}_{field}"]
                assert x.shape == want.shape, (name, mb, field, x.shape, want.shape)
                assert np.array_equal(x.cpu().numpy(), want), (name, mb, field)
            n_mb += 1
        assert n_mb == (3 if name == "feed_forward" else 2)
    buf.after_update()
    assert np.array_equal(buf.obs[0].cpu().numpy(), g["after_update_obs0"])
    assert np.array_equal(buf.masks[0].cpu().numpy(), g["after_update_masks0"])


class _VN:
    """Only what compute_returns reads from the reference's ValueNorm (valuenorm.py:31-35)."""

    def __init__(self, scale, shift):
        self.var, self.mean = torch.tensor([float(scale) ** 2], dtype=torch.float64), torch.tensor([float(shift)])
        self.scale = scale

    def running_mean_var(self):
        return self.mean, torch.tensor([float(self.scale)]) ** 2


@pytest.mark.parametrize("tag", ["vn", "plain", "nogae", "long"])
def test_compute_returns_matches_reference_fixture(cuda, tag):
    g = np.load(os.path.join(GOLDEN, "gae.npz"))
    use_gae, use_vn = (bool(x) for x in g[f"{tag}_flags"])
    T, B = g[f"{tag}_rewards"].shape[:2]
    buf = _buffer(cuda, T, B, use_gae, use_vn)
    buf.rewards.copy_(torch.from_numpy(g[f"{tag}_rewards"]))
    buf.value_preds.copy_(torch.from_numpy(g[f"{tag}_value_preds_in"]))
    buf.masks.copy_(torch.from_numpy(g[f"{tag}_masks"]))
    vn = None
    if use_vn:
        scale, shift = g[f"{tag}_denorm"]

        class VN:
            def running_mean_var(self):
                # sqrt(var) must round-trip to the recorded float32 scale
                return torch.tensor([shift]), torch.tensor([scale], dtype=torch.float64).pow(2)
        vn = VN()
    buf.compute_returns(g[f"{tag}_next_value"], vn)
    if use_gae:
        assert np.array_equal(buf.returns[:-1].cpu().numpy(), g[f"{tag}_returns"][:-1])
        assert np.array_equal(buf.advantages.cpu().numpy(), g[f"{tag}_advantages"])
    else:
        assert np.array_equal(buf.returns.cpu().numpy(), g[f"{tag}_returns"])


def test_env_writes_rollout_slots_in_place(cuda):
    """The env kernel fills obs[t+1] / rewards[t] of the buffer directly; insert() then skips those copies."""
    from minimarl_b200 import BatchedCheckers
    from oracle.magym_c import CheckersBatchRef
    T, N = 8, 64
    buf = _buffer(cuda, T, N)
    env = BatchedCheckers(N, max_steps=100, device=cuda)
    ref = CheckersBatchRef(N, False, 100, -0.01)
    env.reset(out=buf.obs[0])
    assert np.array_equal(buf.obs[0].cpu().numpy(), ref.reset())
    dones = torch.zeros((N, 2), dtype=torch.uint8, device=cuda)
    rng = np.random.default_rng(0)
    for t in range(T):
        acts = rng.integers(0, 5, (N, 2), dtype=np.int32)
        env.step_into(torch.from_numpy(acts).to(cuda), buf.obs[t + 1], buf.rewards[t].view(N, 2), dones)
        masks = (1.0 - dones.float()).unsqueeze(-1)
        zeros_h = torch.zeros((N, 2, 1, 32), device=cuda)
        buf.insert(buf.obs[t + 1], buf.obs[t + 1], zeros_h, zeros_h, torch.from_numpy(acts).to(cuda).float().unsqueeze(-1),
                   torch.zeros((N, 2, 1), device=cuda), torch.zeros((N, 2, 1), device=cuda), buf.rewards[t], masks)
        o, r, d = ref.step(acts)
        assert np.array_equal(buf.obs[t + 1].cpu().numpy(), o)
        assert np.array_equal(buf.rewards[t].view(N, 2).cpu().numpy(), r)
        assert np.array_equal(buf.share_obs[t + 1].cpu().numpy(), o)
    assert buf.step == 0
