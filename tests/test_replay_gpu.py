"""GPU parity: device segment trees / PER sampling / episode gather vs the reference-recorded fixtures
(tests/golden/per.npz) and vs the numpy oracle at larger sizes.  Indices and gathered payload are
bit-exact; importance weights go through pow(): tolerance 1e-12 relative."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle.per_ref import EpisodeReplayRef, SegTreeRef
from minimarl_b200.envs.spaces import Box, Discrete

pytestmark = pytest.mark.gpu


def _info(D=47, A=2, nA=5):
    return {"policy_0": {"obs_space": Box(np.zeros(D), np.ones(D)), "share_obs_space": Box(np.zeros(A * D), np.ones(A * D)),
                         "act_space": Discrete(nA)}}


def test_segment_trees_match_reference_fixture(cuda):
    from minimarl_b200.buffers import MinSegmentTree, SumSegmentTree
    g = np.load(os.path.join(GOLDEN, "per.npz"))
    rng = np.random.default_rng(7)
    st, mt = SumSegmentTree(64, cuda), MinSegmentTree(64, cuda)
    pri = rng.random(50) + 0.01
    st[np.arange(50)] = pri
    mt[np.arange(50)] = pri
    idx = np.array([3, 4, 17, 18, 19, 49])
    newp = rng.random(6) * 2 + 0.01
    st[idx] = newp
    mt[idx] = newp
    assert np.array_equal(st._value.cpu().numpy(), g["tree_values"])
    assert np.array_equal(mt._value.cpu().numpy(), g["min_values"])
    assert np.array_equal(st.find_prefixsum_idx(g["tree_mass"]).cpu().numpy(), g["tree_idx"])
    assert st.sum(0, 49).item() == g["tree_sum_0_49"]
    assert mt.min().item() == g["tree_min"]
    st[5] = 0.125                       # scalar index / scalar value form
    assert st[5].item() == 0.125


def test_prioritized_buffer_matches_reference_fixture(cuda):
    from minimarl_b200.buffers import PrioritizedRecReplayBuffer
    g = np.load(os.path.join(GOLDEN, "per.npz"))
    buf = PrioritizedRecReplayBuffer(0.6, _info(), {"policy_0": [0, 1]}, 24, 6, True, False, device=cuda)
    for it in range(5):
        ep = {k: {"policy_0": g[f"per_insert{it}_{k}"]} for k in ["obs", "share_obs", "acts", "rewards", "dones", "dones_env"]}
        slots = buf.insert(7, ep["obs"], ep["share_obs"], ep["acts"], ep["rewards"], ep["dones"], ep["dones_env"])
        assert np.array_equal(slots, g[f"per_insert{it}_idx"])
        if it >= 1:
            buf.update_priorities(g[f"per_update{it}_idx"], g[f"per_update{it}_p"], "policy_0")
    assert np.array_equal(buf._it_sums["policy_0"]._value.cpu().numpy(), g["per_sum_tree"])
    assert np.array_equal(buf._it_mins["policy_0"]._value.cpu().numpy(), g["per_min_tree"])
    assert buf.max_priorities["policy_0"] == g["per_max_priority"] and len(buf) == g["per_len"]
    out = buf.sample(8, beta=0.4, p_id="policy_0", uniforms=g["per_sample_u"])
    for k, v in zip(["obs", "share_obs", "acts", "rewards", "dones", "dones_env"], out[:6]):
        assert np.array_equal(v["policy_0"].cpu().numpy(), g[f"per_sample_{k}"]), k
    assert out[6]["policy_0"] is None
    assert np.array_equal(out[8].cpu().numpy(), g["per_sample_idx"])
    assert np.allclose(out[7].cpu().numpy(), g["per_sample_weights"], rtol=1e-12, atol=0)
    with pytest.raises(AssertionError):
        buf.update_priorities(np.array([0]), np.array([0.0]), "policy_0")
    with pytest.raises(AssertionError):
        buf.sample(64, beta=0.4, p_id="policy_0")


def test_large_buffer_against_numpy_oracle(cuda):
    """1 000-episode buffer, batch 512: indices / payload exact vs the oracle, ragged ring wrap-around."""
    from minimarl_b200.buffers import PrioritizedRecReplayBuffer
    T, A, D, nA, S = 20, 2, 47, 5, 1000
    rng = np.random.default_rng(3)
    dev = PrioritizedRecReplayBuffer(0.6, _info(), {"policy_0": [0, 1]}, S, T, True, False, device=cuda)
    ref = EpisodeReplayRef(0.6, S, T, A, D, A * D, nA)
    for n in (300, 450, 400):            # the last insert wraps around the ring
        ep = dict(obs=rng.random((T + 1, n, A, D), dtype=np.float32),
                  share_obs=rng.random((T + 1, n, A, A * D), dtype=np.float32),
                  acts=np.eye(nA, dtype=np.float32)[rng.integers(0, nA, (T, n, A))],
                  rewards=rng.standard_normal((T, n, A, 1)).astype(np.float32),
                  dones=(rng.random((T, n, A, 1)) < 0.1).astype(np.float32),
                  dones_env=(rng.random((T, n, 1)) < 0.1).astype(np.float32))
        a = dev.insert(n, *[{"policy_0": ep[k]} for k in ["obs", "share_obs", "acts", "rewards", "dones", "dones_env"]])
        b = ref.insert(n, *[ep[k] for k in ["obs", "share_obs", "acts", "rewards", "dones", "dones_env"]])
        assert np.array_equal(a, b)
        idx = rng.choice(len(ref), 256, replace=False)
        pr = rng.random(256) * 5 + 1e-3
        dev.update_priorities(idx, pr, "policy_0")
        ref.update_priorities(idx, pr)
    assert np.array_equal(dev._it_sums["policy_0"]._value.cpu().numpy(), ref.sum_tree.value)   # host pow: exact
    # make the trees identical (pow on the device may differ from libm in the last ulp) and compare sampling exactly
    dev._it_sums["policy_0"]._value.copy_(torch.from_numpy(ref.sum_tree.value))
    dev._it_mins["policy_0"]._value.copy_(torch.from_numpy(ref.min_tree.value))
    u = rng.random(512)
    out = dev.sample(512, beta=0.7, p_id="policy_0", uniforms=u)
    want = ref.sample(u, 0.7)
    assert np.array_equal(out[8].cpu().numpy(), want[7])
    for k, (v, w) in enumerate(zip(out[:6], want[:6])):
        assert np.array_equal(v["policy_0"].cpu().numpy(), w), k
    assert np.allclose(out[7].cpu().numpy(), want[6], rtol=1e-12, atol=0)


def test_uniform_buffer_and_gather_roundtrip(cuda):
    """Round trip: gather(all indices in order) then un-transpose == storage (a size-independent property)."""
    from minimarl_b200.buffers import RecReplayBuffer
    T, A, D, nA, S = 10, 2, 47, 5, 64
    buf = RecReplayBuffer(_info(), {"policy_0": [0, 1]}, S, T, True, False, device=cuda)
    pb = buf.policy_buffers["policy_0"]
    g = torch.Generator(device=cuda).manual_seed(0)
    for x in (pb.obs, pb.share_obs, pb.acts, pb.rewards, pb.dones, pb.dones_env):
        x.copy_(torch.rand(x.shape, device=cuda, generator=g))
    pb.filled_i = S
    obs, share_obs, acts, rewards, dones, dones_env, _ = pb.sample_inds(torch.arange(S, device=cuda))
    assert torch.equal(obs.permute(1, 2, 0, 3), pb.obs)
    assert torch.equal(share_obs, pb.share_obs)
    assert torch.equal(acts.permute(1, 2, 0, 3), pb.acts)
    assert torch.equal(rewards.permute(1, 2, 0, 3), pb.rewards) and torch.equal(dones.permute(1, 2, 0, 3), pb.dones)
    assert torch.equal(dones_env, pb.dones_env)
    s = buf.sample(16)
    assert s[0]["policy_0"].shape == (A, T + 1, 16, D) and s[7] is None and s[8] is None


def test_gather_layouts_agree(cuda):
    """mm_replay_gather (agent-major [A, T, B, d], the reference's `_cast`) and mm_replay_gather_stacked ([T, A, B, d], what
    the buffer stores so that the learner's torch.cat over the agents is a view) hold the same values, for the fast (A = 2) and the
    generic (A = 3) kernel; the stacked rows are the reference's cat order r = a * B + b."""
    from minimarl_b200 import _lib
    for A, D, SD, NA in ((2, 47, 94, 5), (3, 10, 30, 4)):
        T, S, B = 6, 40, 17
        g = torch.Generator(device=cuda).manual_seed(A)
        r = lambda *s: torch.rand(*s, device=cuda, generator=g)
        obs, share, acts = r(T + 1, S, A, D), r(T + 1, S, SD), r(T, S, A, NA)
        rew, dn, de = r(T, S, A, 1), r(T, S, A, 1), r(T, S, 1)
        idx = torch.randint(0, S, (B,), device=cuda, generator=g)
        outs = []
        for name in ("mm_replay_gather", "mm_replay_gather_stacked"):
            st = name.endswith("stacked")
            o_obs = torch.empty((T + 1, A, B, D) if st else (A, T + 1, B, D), device=cuda)
            o_acts = torch.empty((T, A, B, NA) if st else (A, T, B, NA), device=cuda)
            o_sh, o_rw, o_dn, o_de = torch.empty(T + 1, B, SD, device=cuda), torch.empty(A, T, B, 1, device=cuda), \
                torch.empty(A, T, B, 1, device=cuda), torch.empty(T, B, 1, device=cuda)
            p = _lib.ptr
            _lib.check(getattr(_lib.lib(), name)(p(obs), p(share), p(acts), p(rew), p(dn), p(de), p(idx), p(o_obs), p(o_sh), p(o_acts),
                                                 p(o_rw), p(o_dn), p(o_de), T, S, B, A, D, SD, NA, _lib.stream()))
            outs.append((o_obs.permute(1, 0, 2, 3) if st else o_obs, o_acts.permute(1, 0, 2, 3) if st else o_acts, o_sh, o_rw, o_dn, o_de))
        for a, b in zip(*outs):
            assert torch.equal(a, b)
        assert torch.equal(outs[0][0], obs[:, idx].permute(2, 0, 1, 3))              # [A, T+1, B, D]
        stacked = outs[1][0].permute(1, 0, 2, 3).reshape(T + 1, A * B, D)
        assert stacked.data_ptr() == outs[1][0].data_ptr() and torch.equal(stacked, torch.cat(list(outs[0][0]), dim=-2))
