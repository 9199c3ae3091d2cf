"""GPU parity: CUDA env kernels (through the C ABI) vs the CPU oracle, bit-exact.

Protocol: the envs are deterministic, so replaying one action sequence through
both sides is a complete parity check.  Compared every step: observations, rewards, dones, and the
decoded state (positions, step counts, food left).
"""
import numpy as np
import pytest
import torch

from oracle.magym_c import CheckersBatchRef, SwitchBatchRef
from oracle.ma_gym_ref import CheckersRef, SwitchRef

pytestmark = pytest.mark.gpu


def _food_counts(env):
    bm = env.food_bitmap().cpu().numpy()
    apple_cells = 0x15 | (0x2A << 8) | (0x15 << 16)
    apples = np.array([bin(int(b) & apple_cells).count("1") for b in bm])
    lemons = np.array([bin(int(b) & (0x3F3F3F ^ apple_cells)).count("1") for b in bm])
    return lemons, apples


@pytest.mark.parametrize("n_envs", [1, 16, 333, 1000, 4096])
@pytest.mark.parametrize("full", [False, True])
def test_checkers_step_matches_oracle(cuda, n_envs, full):
    from minimarl_b200 import BatchedCheckers
    max_steps = 40
    env = BatchedCheckers(n_envs, full_observable=full, max_steps=max_steps, step_cost=-0.01, device=cuda)
    ref = CheckersBatchRef(n_envs, full, max_steps, -0.01)
    assert np.array_equal(env.reset().cpu().numpy(), ref.reset())
    rng = np.random.default_rng(42 + n_envs)
    for t in range(max_steps + 8):           # runs past done: rewards must turn to 0, positions still move
        acts = rng.integers(0, 5, size=(n_envs, 2), dtype=np.int32)
        obs, rew, done = env.step(torch.from_numpy(acts).to(cuda))
        o_ref, r_ref, d_ref = ref.step(acts)
        assert np.array_equal(obs.cpu().numpy(), o_ref), f"obs differ at step {t}"
        assert np.array_equal(rew.cpu().numpy(), r_ref), f"rewards differ at step {t}"
        assert np.array_equal(done.cpu().numpy(), d_ref), f"dones differ at step {t}"
    pos, steps, food = ref.export()
    assert np.array_equal(env.positions().cpu().numpy(), pos)
    assert np.array_equal(env.step_count().cpu().numpy(), steps)
    lemons, apples = _food_counts(env)
    assert np.array_equal(lemons, food[:, 0]) and np.array_equal(apples, food[:, 1])
    env.check_actions()


def test_checkers_adversarial_sequences(cuda):
    """Collisions, swaps through a vacated cell, walls, eating every food, early termination."""
    from minimarl_b200 import BatchedCheckers
    seqs = []
    # both agents walk into each other in column 6/7 and try to swap
    seqs.append([(0, 2), (0, 2), (4, 2), (0, 4), (2, 0), (3, 3), (1, 1), (1, 1)] * 6)
    # agent 1 follows agent 0 into the cell it vacates
    seqs.append([(1, 2), (1, 2), (1, 3), (1, 1), (0, 1), (0, 1), (3, 3), (2, 2)] * 6)
    # snake over the whole board so all apples get eaten (early done)
    snake = [(1, 1)] * 6 + [(0, 2)] + [(3, 3)] * 5 + [(0, 2)] + [(1, 1)] * 6 + [(2, 0)] * 2 + [(3, 3)] * 7
    seqs.append((snake * 3)[:48])
    T = 48
    n = len(seqs)
    acts = np.array([[s[t] for s in seqs] for t in range(T)], dtype=np.int32)   # [T, n, 2]
    env = BatchedCheckers(n, full_observable=False, max_steps=100, device=cuda)
    ref = CheckersBatchRef(n, False, 100, -0.01)
    env.reset(); ref.reset()
    for t in range(T):
        obs, rew, done = env.step(torch.from_numpy(acts[t]).to(cuda))
        o, r, d = ref.step(acts[t])
        assert np.array_equal(obs.cpu().numpy(), o) and np.array_equal(rew.cpu().numpy(), r)
        assert np.array_equal(done.cpu().numpy(), d)


@pytest.mark.parametrize("full", [False, True])
def test_checkers_rollout_kernel_equals_stepwise(cuda, full):
    from minimarl_b200 import BatchedCheckers
    n, T = 777, 37
    a = BatchedCheckers(n, full_observable=full, max_steps=30, device=cuda)
    b = BatchedCheckers(n, full_observable=full, max_steps=30, device=cuda)
    a.reset(); b.reset()
    acts = torch.randint(0, 5, (T, n, 2), dtype=torch.int32, device=cuda,
                         generator=torch.Generator(device=cuda).manual_seed(7))
    obs, rew, done = a.rollout(T, actions=acts)
    for t in range(T):
        o, r, d = b.step(acts[t])
        assert torch.equal(obs[t], o) and torch.equal(rew[t], r) and torch.equal(done[t], d)
    assert torch.equal(a.state, b.state)


def test_checkers_device_random_rollout_replays_through_oracle(cuda):
    from minimarl_b200 import BatchedCheckers
    n, T = 500, 64
    env = BatchedCheckers(n, max_steps=50, device=cuda)
    env.reset()
    acts = torch.empty((T, n, 2), dtype=torch.int32, device=cuda)
    obs, rew, done = env.rollout(T, actions=None, seed=1234, actions_out=acts)
    a = acts.cpu().numpy()
    assert a.min() >= 0 and a.max() <= 4
    counts = np.bincount(a.reshape(-1), minlength=5) / a.size
    assert np.all(np.abs(counts - 0.2) < 0.02), counts        # uniform over the 5 actions
    ref = CheckersBatchRef(n, False, 50, -0.01)
    ref.reset()
    for t in range(T):
        o, r, d = ref.step(a[t])
        assert np.array_equal(obs[t].cpu().numpy(), o)
        assert np.array_equal(rew[t].cpu().numpy(), r) and np.array_equal(done[t].cpu().numpy(), d)
    # a different seed or step offset draws a different stream
    env.reset()
    acts2 = torch.empty_like(acts)
    env.rollout(T, actions=None, seed=1235, actions_out=acts2)
    assert not torch.equal(acts, acts2)


def test_checkers_zero_copy_into_buffer_slots(cuda):
    """step_into writes straight into [T+1,N,A,D] buffer slices, including 8-byte-aligned odd offsets."""
    from minimarl_b200 import BatchedCheckers
    n, T = 333, 5                     # odd n: slot t+1 starts 8-byte aligned only
    env = BatchedCheckers(n, max_steps=100, device=cuda)
    ref = CheckersBatchRef(n, False, 100, -0.01)
    obs_buf = torch.zeros((T + 1, n, 2, 47), device=cuda)
    rew_buf = torch.zeros((T, n, 2), device=cuda)
    done_buf = torch.zeros((T, n, 2), dtype=torch.uint8, device=cuda)
    env.reset(out=obs_buf[0])
    assert np.array_equal(obs_buf[0].cpu().numpy(), ref.reset())
    rng = np.random.default_rng(3)
    for t in range(T):
        acts = rng.integers(0, 5, size=(n, 2), dtype=np.int32)
        env.step_into(torch.from_numpy(acts).to(cuda), obs_buf[t + 1], rew_buf[t], done_buf[t])
        o, r, d = ref.step(acts)
        assert np.array_equal(obs_buf[t + 1].cpu().numpy(), o)
        assert np.array_equal(rew_buf[t].cpu().numpy(), r)
        assert np.array_equal(done_buf[t].cpu().numpy().astype(bool), d)


def test_checkers_auto_reset_and_masked_reset(cuda):
    from minimarl_b200 import BatchedCheckers
    n = 100
    env = BatchedCheckers(n, max_steps=7, auto_reset=True, device=cuda)
    ref = CheckersBatchRef(n, False, 7, -0.01, auto_reset=True)
    env.reset(); ref.reset()
    rng = np.random.default_rng(5)
    for t in range(30):
        acts = rng.integers(0, 5, size=(n, 2), dtype=np.int32)
        obs, rew, done = env.step(torch.from_numpy(acts).to(cuda))
        o, r, d = ref.step(acts)
        assert np.array_equal(obs.cpu().numpy(), o) and np.array_equal(rew.cpu().numpy(), r)
        assert np.array_equal(done.cpu().numpy(), d)
    # masked reset: only the selected envs go back to the start state
    env2 = BatchedCheckers(n, max_steps=50, device=cuda)
    first = env2.reset().clone()
    for t in range(5):
        env2.step(torch.from_numpy(rng.integers(0, 5, size=(n, 2), dtype=np.int32)).to(cuda))
    before = env2.state.clone()
    mask = torch.zeros(n, dtype=torch.bool, device=cuda)
    mask[::3] = True
    obs = env2.reset(mask=mask)
    assert torch.equal(obs[mask], first[mask])
    assert torch.equal(env2.state[:, ~mask], before[:, ~mask])
    assert torch.equal(env2.step_count()[mask], torch.zeros_like(env2.step_count()[mask]))


def test_checkers_bad_action_raises_like_reference(cuda):
    from minimarl_b200 import BatchedCheckers
    env = BatchedCheckers(8, device=cuda)
    env.reset()
    acts = torch.zeros((8, 2), dtype=torch.int32, device=cuda)
    acts[3, 1] = 7
    env.step(acts)
    with pytest.raises(Exception, match="Action Not found!"):
        env.check_actions()


@pytest.mark.parametrize("env_id,kw", [
    ("ma_gym:Checkers-v0", dict(full_observable=True, max_steps=100, step_cost=-0.01)),   # vdn/main.py:61-64
    ("ma_gym:Checkers-v0", dict(full_observable=False, max_steps=100, step_cost=-0.01)),  # mappo/main.py:41-46
    ("ma_gym:Switch2-v0", dict(max_steps=100, step_cost=-0.01)),                          # qmix/main.py:66-71
])
def test_single_env_adaptor_matches_python_oracle(cuda, env_id, kw):
    """The gym-style 1-env object the reference mains drive, against the pure-Python restatement."""
    import minimarl_b200
    from oracle import ma_gym_ref
    env = minimarl_b200.make(env_id, **kw)
    ref = ma_gym_ref.make(env_id, **kw)
    assert env.n_agents == ref.n_agents
    assert env.observation_space[0].shape == ref.observation_space[0].shape
    assert env.action_space[0].n == 5
    o, o_ref = env.reset(), ref.reset()
    assert np.array_equal(np.float32(o), np.float32(o_ref))
    rng = np.random.default_rng(0)
    done = [False] * env.n_agents
    steps = 0
    while not all(done):
        a = rng.integers(0, 5, size=env.n_agents)
        # the reference passes a float tensor row (vdn/main.py:93) -- so do we
        o, r, done, _ = env.step(torch.tensor(a, dtype=torch.float32))
        o_ref, r_ref, d_ref, _ = ref.step(list(a))
        assert np.array_equal(np.float32(o), np.float32(o_ref))
        assert r == r_ref                     # Python doubles, exactly the reference's values
        assert done == d_ref
        steps += 1
    assert steps <= kw["max_steps"]
    with pytest.raises(Exception, match="Action Not found!"):
        env.step([9] * env.n_agents)


@pytest.mark.parametrize("n_agents", [2, 4])
@pytest.mark.parametrize("full", [False, True])
def test_switch_step_matches_oracle(cuda, n_agents, full):
    from minimarl_b200 import BatchedSwitch
    n, max_steps = 1001, 60
    env = BatchedSwitch(n, n_agents=n_agents, full_observable=full, max_steps=max_steps, step_cost=-0.1,
                        device=cuda)
    ref = SwitchBatchRef(n, n_agents, full, max_steps, -0.1)
    assert np.array_equal(env.reset().cpu().numpy(), ref.reset())
    rng = np.random.default_rng(11)
    goal_rewards = partial_done = 0
    # biased walk so that agents do reach their goals through the corridor
    for t in range(max_steps + 5):
        acts = rng.integers(0, 5, size=(n, n_agents), dtype=np.int32)
        push = rng.random((n, n_agents)) < 0.5
        toward = np.where(np.arange(n_agents) % 2 == 0, 3, 1)[None, :].repeat(n, 0)
        acts = np.where(push, toward, acts).astype(np.int32)
        obs, rew, done = env.step(torch.from_numpy(acts).to(cuda))
        o, r, d = ref.step(acts)
        assert np.array_equal(obs.cpu().numpy(), o), t
        assert np.array_equal(rew.cpu().numpy(), r), t
        assert np.array_equal(done.cpu().numpy(), d), t
        if t < max_steps - 1:
            goal_rewards += int((rew.cpu().numpy() == 5).sum())
            partial_done += int((d.any(1) & ~d.all(1)).sum())
    assert np.array_equal(env.positions().cpu().numpy(), ref.export())
    # the +5 path and per-agent dones are really exercised: agents reach their goals before the horizon, and there are
    # steps at which only some agents of an env are done
    assert goal_rewards > n // 4, goal_rewards
    assert partial_done > 0


@pytest.mark.parametrize("n_agents,full", [(2, False), (4, True)])
def test_switch_rollout_kernel_matches_oracle(cuda, n_agents, full):
    """mm_switch_rollout (T steps in one launch): given actions vs the oracle step by step; device-drawn actions are
    written out and replay through the oracle to the same observations, rewards, dones and final positions."""
    from minimarl_b200 import BatchedSwitch
    n, T, max_steps = 777, 45, 40                                # runs past max_steps
    env = BatchedSwitch(n, n_agents=n_agents, full_observable=full, max_steps=max_steps, step_cost=-0.1, device=cuda)
    ref = SwitchBatchRef(n, n_agents, full, max_steps, -0.1)
    env.reset(); ref.reset()
    rng = np.random.default_rng(5)
    acts = rng.integers(0, 5, size=(T, n, n_agents), dtype=np.int32)
    toward = np.where(np.arange(n_agents) % 2 == 0, 3, 1)[None, None, :]
    acts = np.where(rng.random(acts.shape) < 0.5, toward, acts).astype(np.int32)
    obs, rew, done = env.rollout(T, actions=torch.from_numpy(acts).to(cuda))
    for t in range(T):
        o, r, d = ref.step(acts[t])
        assert np.array_equal(obs[t].cpu().numpy(), o), t
        assert np.array_equal(rew[t].cpu().numpy(), r) and np.array_equal(done[t].cpu().numpy(), d), t
    assert np.array_equal(env.positions().cpu().numpy(), ref.export())
    # device-drawn actions, two launches that continue each other (step_offset)
    env.reset(); ref.reset()
    drawn = torch.empty((T, n, n_agents), dtype=torch.int32, device=cuda)
    o1, r1, d1 = env.rollout(20, seed=9, step_offset=0, actions_out=drawn[:20])
    o2, r2, d2 = env.rollout(T - 20, seed=9, step_offset=20, actions_out=drawn[20:])
    a = drawn.cpu().numpy()
    assert a.min() >= 0 and a.max() <= 4 and len(np.unique(a)) == 5
    obs, rew, done = torch.cat([o1, o2]), torch.cat([r1, r2]), torch.cat([d1, d2])
    for t in range(T):
        o, r, d = ref.step(a[t])
        assert np.array_equal(obs[t].cpu().numpy(), o), t
        assert np.array_equal(rew[t].cpu().numpy(), r) and np.array_equal(done[t].cpu().numpy(), d), t
    assert np.array_equal(env.positions().cpu().numpy(), ref.export())


def test_full_size_rollout_properties(cuda):
    """BASELINE sizes (16 384 envs x T=128): size-independent properties instead of the slow oracle."""
    from minimarl_b200 import BatchedCheckers
    n, T = 16384, 128
    env = BatchedCheckers(n, max_steps=100, device=cuda)
    obs0 = env.reset()
    obs, rew, done = env.rollout(T, actions=None, seed=42)
    # every obs value is 0/1 or one of the position fractions, one-hot slots are one-hot
    cells = obs[..., 2:].reshape(T, n, 2, 9, 5)
    assert torch.all((cells == 0) | (cells == 1))
    assert torch.all(cells.sum(-1) <= 1)
    assert torch.all(cells[..., 4, 2 + 0].flatten(0, 1)[:, 0] == 1)   # agent 0 sees itself in the centre cell
    assert torch.all(cells[..., 4, 2 + 1].flatten(0, 1)[:, 1] == 1)
    # dones are monotone in t and everything is done from step 100 on; rewards are 0 after done
    d = done.to(torch.int8)
    assert torch.all(d[1:] >= d[:-1])
    assert torch.all(done[99:])
    after = done[:-1]
    assert torch.all(rew[1:][after] == 0)
    # reward lattice: step cost plus at most one food item per agent-step
    allowed = torch.tensor([-0.01, -10.01, 9.99, -1.01, 0.99, 0.0], device=cuda).to(torch.float32)
    assert torch.all((rew.unsqueeze(-1) == allowed).any(-1))
    # idempotence: same seed, same rollout
    env.reset()
    obs_b, rew_b, done_b = env.rollout(T, actions=None, seed=42)
    assert torch.equal(obs, obs_b) and torch.equal(rew, rew_b) and torch.equal(done, done_b)
    assert torch.equal(env.reset(), obs0)


def test_full_size_rollout_matches_oracle(cuda):
    """BASELINE configs[2] size, 16 384 envs x T = 128 (past the 100-step horizon), device-drawn actions: every observation,
    reward and done of all 2.1 M env-steps bit-equal to the C oracle replaying the recorded actions."""
    from minimarl_b200 import BatchedCheckers
    n, T = 16384, 128
    env = BatchedCheckers(n, max_steps=100, device=cuda)
    ref = CheckersBatchRef(n, False, 100, -0.01)
    assert np.array_equal(env.reset().cpu().numpy(), ref.reset())
    drawn = torch.empty((T, n, 2), dtype=torch.int32, device=cuda)
    obs, rew, done = env.rollout(T, actions=None, seed=42, actions_out=drawn)
    acts = drawn.cpu().numpy()
    obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
    for t in range(T):
        o, r, d = ref.step(acts[t])
        assert np.array_equal(obs[t], o), t
        assert np.array_equal(rew[t], r) and np.array_equal(done[t], d), t


@pytest.mark.parametrize("transport", ["direct", "compact"])
@pytest.mark.parametrize("T,chunk,full", [(1, 8, False), (5, 8, False), (8, 8, True), (37, 8, False), (48, 5, True)])
def test_checkers_rollout_host_matches_oracle(cuda, T, chunk, full, transport):
    """Host-buffer rollout (mm_checkers_rollout_host: chunked H2D / kernel / D2H pipeline) vs the oracle, incl.
    ragged last chunks, a single chunk, and a second call that continues from the state the first one left."""
    from minimarl_b200 import BatchedCheckers
    n = 333
    env = BatchedCheckers(n, full_observable=full, max_steps=30, device=cuda)
    ref = CheckersBatchRef(n, full, 30, -0.01)
    env.reset(); ref.reset()
    rng = np.random.default_rng(T * 100 + chunk)
    for call in range(2):
        acts = rng.integers(0, 5, size=(T, n, 2), dtype=np.int32)
        obs, rew, done = env.rollout_host(acts, chunk=chunk, transport=transport, n_threads=3)
        for t in range(T):
            o, r, d = ref.step(acts[t])
            assert np.array_equal(obs[t], o), f"call {call}: obs differ at step {t}"
            assert np.array_equal(rew[t], r) and np.array_equal(done[t], d), f"call {call}: step {t}"
    # actions written straight into the page-locked staging array, observations not requested
    h = env.host_buffers(T, chunk)
    h["act"].numpy()[...] = rng.integers(0, 5, size=(T, n, 2), dtype=np.int32)
    acts = h["act"].numpy().copy()
    obs, rew, done = env.rollout_host(h["act"].numpy(), chunk=chunk, want_obs=False, transport=transport)
    assert obs is None
    for t in range(T):
        _, r, d = ref.step(acts[t])
        assert np.array_equal(rew[t], r) and np.array_equal(done[t], d)
    pos, steps, _ = ref.export()
    assert np.array_equal(env.positions().cpu().numpy(), pos) and np.array_equal(env.step_count().cpu().numpy(), steps)


def test_rollout_host_auto_transport_picks_one_and_stays_exact(cuda):
    """transport="auto": one timed call of each path, then the faster one; every call returns the oracle's arrays."""
    from minimarl_b200 import BatchedCheckers
    n, T = 2048, 16
    env = BatchedCheckers(n, max_steps=30, device=cuda, auto_reset=True)
    ref = CheckersBatchRef(n, False, 30, -0.01, auto_reset=True)
    env.reset(); ref.reset()
    rng = np.random.default_rng(3)
    for call in range(4):
        acts = rng.integers(0, 5, size=(T, n, 2), dtype=np.int32)
        obs, rew, done = env.rollout_host(acts)
        for t in range(T):
            o, r, d = ref.step(acts[t])
            assert np.array_equal(obs[t], o) and np.array_equal(rew[t], r) and np.array_equal(done[t], d), (call, t)
    assert env.host_buffers(T)["transport"] in ("direct", "compact")
