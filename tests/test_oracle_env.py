"""CPU: the two statements of the env oracle agree, and the oracle reproduces what the reference's own
logs say about the env (the only evidence there is: ma-gym itself is not vendored -- parity unpinned)."""
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN
from oracle.ma_gym_ref import CheckersRef, SwitchRef, make
from oracle.magym_c import CheckersBatchRef, SwitchBatchRef


@pytest.mark.parametrize("full", [False, True])
def test_checkers_c_equals_python(full):
    rnd = random.Random(0)
    n = 48
    pys = [CheckersRef(full_observable=full, max_steps=60) for _ in range(n)]
    c = CheckersBatchRef(n, full, 60, -0.01)
    assert np.array_equal(c.reset(), np.array([e.reset() for e in pys], dtype=np.float32))
    for t in range(75):
        acts = np.array([[rnd.randrange(5), rnd.randrange(5)] for _ in range(n)], dtype=np.int32)
        oc, rc, dc = c.step(acts)
        res = [e.step(list(a)) for e, a in zip(pys, acts)]
        assert np.array_equal(oc, np.array([r[0] for r in res], dtype=np.float32)), t
        assert np.array_equal(rc, np.array([r[1] for r in res], dtype=np.float32)), t
        assert np.array_equal(dc, np.array([r[2] for r in res])), t


@pytest.mark.parametrize("n_agents", [2, 4])
@pytest.mark.parametrize("full", [False, True])
def test_switch_c_equals_python(n_agents, full):
    rnd = random.Random(1)
    n = 32
    pys = [SwitchRef(full_observable=full, n_agents=n_agents, max_steps=40, step_cost=-0.1) for _ in range(n)]
    c = SwitchBatchRef(n, n_agents, full, 40, -0.1)
    assert np.array_equal(c.reset(), np.array([e.reset() for e in pys], dtype=np.float32))
    for t in range(45):
        acts = np.array([[rnd.choice([1, 3, 3, 1, 0, 2, 4]) for _ in range(n_agents)] for _ in range(n)], dtype=np.int32)
        oc, rc, dc = c.step(acts)
        res = [e.step(list(a)) for e, a in zip(pys, acts)]
        assert np.array_equal(oc, np.array([r[0] for r in res], dtype=np.float32)), t
        assert np.array_equal(rc, np.array([r[1] for r in res], dtype=np.float32)), t
        assert np.array_equal(dc, np.array([r[2] for r in res])), t


def test_checkers_spec_sheet():
    """dims (정리:3-19), start cells, reward quanta, collision rule."""
    env = make("ma_gym:Checkers-v0", full_observable=False, max_steps=100, step_cost=-0.01)
    obs = env.reset()
    assert len(obs) == 2 and len(obs[0]) == 47 and env.observation_space[0].shape == (47,)
    assert obs[0][:2] == [0.0, 0.86] and obs[1][:2] == [0.67, 0.86]
    full = make("ma_gym:Checkers-v0", full_observable=True)
    assert len(full.reset()[0]) == 94 and full.reset()[0] == full.reset()[1]
    # agent 0 steps left onto the lemon at (0,5): -10 - 0.01; agent 1 up onto the apple (1,6)? (1,6) is empty
    _, r, d, _ = env.step([1, 2])
    assert r == [-0.01 - 10, -0.01] and d == [False, False]
    # agent 1 (now at (1,6)) moves left onto the apple at (1,5): +1
    _, r, _, _ = env.step([4, 1])
    assert r == [-0.01, -0.01 + 1]
    # agent 0 (0,5) tries to move down into agent 1 at (1,5): blocked
    before = [list(p) for p in env.agent_pos.values()]
    env.step([0, 4])
    assert [list(p) for p in env.agent_pos.values()] == before
    # agent order: agent 0 moves away first, so agent 1 may enter the vacated cell in the same step
    env.step([1, 2])
    assert env.agent_pos[0] == [0, 4] and env.agent_pos[1] == [0, 5]
    with pytest.raises(Exception, match="Action Not found!"):
        env.step([5, 0])


def test_checkers_horizon_and_beyond_done():
    env = CheckersRef(max_steps=5)
    env.reset()
    for t in range(5):
        _, r, d, _ = env.step([4, 4])
        assert d == [t == 4] * 2 and r == [-0.01, -0.01]
    _, r, d, _ = env.step([1, 1])          # stepping a finished episode: zero reward, agents still move
    assert r == [0, 0] and d == [True, True] and env.agent_pos[0] == [0, 5]


def _decompose(score_centi):
    """All (food, k) with score = food - 0.02 k, 1 <= k <= 100 (scores are logged with 2 decimals)."""
    out = []
    for k in range(1, 101):
        if (score_centi + 2 * k) % 100 == 0:
            out.append(((score_centi + 2 * k) // 100, k))
    return out


def test_logged_scores_support_apples_exhausted_rule():
    """52k episode scores from the author's runs (vdn/logs/*.log, score = sum of all rewards,
    vdn/main.py:173).  Episodes that stopped before step 100 must have eaten all 9 apples:
    food = 10 a0 + (9 - a0) - 10 l0 - l1.  Under "all 18 food items eaten" food would be 9 (a0 - l0),
    which most early-stopped episodes contradict."""
    path = os.path.join(GOLDEN, "vdn_log_scores.npz")
    data = np.load(path)
    feasible_apples = {9 * a0 + 9 - 10 * l0 - l1 for a0 in range(10) for l0 in range(10) for l1 in range(10 - l0)}
    n_early = n_bad_allfood = 0
    for name in data.files:
        for s in data[name]:
            s = int(s)
            if s % 100 == 0:
                continue                          # k in {50, 100}: cannot tell, skip
            cands = _decompose(s)
            n_early += 1
            assert any(f in feasible_apples and k >= 9 for f, k in cands), (name, s)
            if not any(f % 9 == 0 and k >= 18 for f, k in cands):
                n_bad_allfood += 1
    assert n_early > 20000
    assert n_bad_allfood > 0.4 * n_early        # the alternative rule cannot explain the logs


def test_random_policy_scores_live_on_the_logged_lattice():
    rnd = random.Random(3)
    env = CheckersRef(full_observable=True)
    early = 0
    for ep in range(300):
        env.reset()
        done, score, k = [False, False], 0.0, 0
        while not all(done):
            _, r, done, _ = env.step([rnd.randrange(5), rnd.randrange(5)])
            score += sum(r)
            k += 1
        food = round(score + 0.02 * k)
        assert abs(score - (food - 0.02 * k)) < 1e-9
        if k < 100:
            early += 1
            assert env._food_count["R"] == 0
    assert 0.15 < early / 300 < 0.6            # the logs show 32-66 % early stops under mostly random play
