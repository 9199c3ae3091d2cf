"""CPU: the numpy replay/PER oracle against fixtures recorded from the reference's own SumSegmentTree /
MinSegmentTree / PrioritizedRecReplayBuffer (tests/golden/make_golden.py)."""
import os

import numpy as np

from conftest import GOLDEN
from oracle.per_ref import EpisodeReplayRef, SegTreeRef


def _g():
    return np.load(os.path.join(GOLDEN, "per.npz"))


def test_segment_trees_match_reference():
    g = _g()
    rng = np.random.default_rng(7)            # same stream as make_golden.golden_segment_tree_and_per
    st, mt = SegTreeRef(64, "sum"), SegTreeRef(64, "min")
    pri = rng.random(50) + 0.01
    st.set(np.arange(50), pri); mt.set(np.arange(50), pri)
    idx = np.array([3, 4, 17, 18, 19, 49])
    newp = rng.random(6) * 2 + 0.01
    st.set(idx, newp); mt.set(idx, newp)
    assert np.array_equal(st.value, g["tree_values"])
    assert np.array_equal(mt.value, g["min_values"])
    assert np.array_equal(st.find_prefixsum(g["tree_mass"]), g["tree_idx"])
    assert st.reduce(0, 49) == g["tree_sum_0_49"]
    assert mt.reduce() == g["tree_min"]


def replay_fixture(buf_factory):
    """Replays the recorded insert / update sequence into any buffer with the oracle's interface."""
    g = _g()
    buf = buf_factory()
    for it in range(5):
        ep = {k: g[f"per_insert{it}_{k}"] for k in ["obs", "share_obs", "acts", "rewards", "dones", "dones_env"]}
        slots = buf.insert(7, ep["obs"], ep["share_obs"], ep["acts"], ep["rewards"], ep["dones"], ep["dones_env"])
        assert np.array_equal(np.asarray(slots), g[f"per_insert{it}_idx"])
        if it >= 1:
            buf.update_priorities(g[f"per_update{it}_idx"], g[f"per_update{it}_p"])
    return g, buf


def test_episode_replay_matches_reference():
    g, buf = replay_fixture(lambda: EpisodeReplayRef(0.6, 24, 6, 2, 47, 94, 5))
    assert np.array_equal(buf.sum_tree.value, g["per_sum_tree"])
    assert np.array_equal(buf.min_tree.value, g["per_min_tree"])
    assert buf.max_priority == g["per_max_priority"] and len(buf) == g["per_len"]
    out = buf.sample(g["per_sample_u"], 0.4)
    for k, v in zip(["obs", "share_obs", "acts", "rewards", "dones", "dones_env"], out[:6]):
        assert np.array_equal(v, g[f"per_sample_{k}"]), k
    assert np.array_equal(out[7], g["per_sample_idx"])
    assert np.allclose(out[6], g["per_sample_weights"], rtol=1e-12, atol=0)
