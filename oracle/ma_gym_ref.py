"""CPU restatement of the two ma-gym grid worlds mini-marl steps (TEST INFRASTRUCTURE ONLY).

This file is the *oracle* for the env rows of SURVEY.md section 8 (a1, a2).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
may import it.  The product path (``minimarl_b200``) never does.

PARITY UNPINNED.  The arithmetic lives in the third-party package ``ma-gym==0.0.14``
(upstream koulanurag/ma-gym: ``ma_gym/envs/checkers/checkers.py`` and
``ma_gym/envs/switch/switch_one_corridor.py``) on ``gym==0.20.0``.  The pins are only recorded
in the reference's wandb artefacts (``vdn/wandb/run-20240312_201751-0trc02l2/files/
requirements.txt:19,23``).  Neither package is vendored under /root/reference, installed in
this image, or downloadable, and the reference holds no test or golden vector for the env.
This restatement follows the published algorithm of that version and is anchored on:

* the reference call sites: ``vdn/main.py:61-64,83,93`` (``gym.make(id, full_observable=True,
  max_steps, step_cost)``, ``reset()``, ``step(action[0])``), ``qmix/main.py:66-71``,
  ``mappo/runner/shared/magym_runner.py:52-57,98,108``,
  ``offpolicy/runner/shared/magym_runner.py:55,94-97``;
* obs dims 47 / 94 from the author's notes (``정리:3-19``) and the prose spec
  ``default/vdn/explain.txt`` (3x8 grid, 2 position values + 3x3x5 neighbourhood);
* the reward lattice and the termination rule recovered from 52k logged episode scores
  (``vdn/logs/vdn-1710766189.log``) -- see ``tests/test_oracle_env.py``.

Everything is plain Python on lists, like the original, so float results are Python doubles
(``round(x, 2)``, ``step_cost + 10``); callers cast to float32 exactly as the reference's
buffers do (``mappo/runner/shared/shared_buffer.py:73-74``, ``vdn/replay_buffer/buffer.py:83``).
"""
import copy

# cell codes of the Checkers grid (strings, like ma-gym's PRE_IDS)
EMPTY, LEMON, APPLE, WALL, AGENT = "0", "Y", "R", "W", "A"
# one-hot channel order of the 3x3 neighbourhood: lemon, apple, A1, A2, wall
_CHANNEL = {LEMON: 0, APPLE: 1, "A1": 2, "A2": 3, WALL: 4}


class _Box:
    """Just enough of gym.spaces.Box for the reference's constructors (``.shape``, ``.low``)."""

    def __init__(self, low, high):
        self.low, self.high = list(low), list(high)
        self.shape = (len(self.low),)
        self.dtype = "float32"


class _Discrete:
    def __init__(self, n):
        self.n = n
        self.shape = ()


class CheckersRef:
    """ma-gym ``Checkers-v0``: 3x8 grid, 2 agents, 9 apples + 9 lemons.

    ``done_rule`` selects the food-exhaustion clause of the termination test.  ``"apples"``
    (default) is the rule the reference's logged scores support; ``"all_food"`` is kept only so
    the log-anchor test can show that it is NOT what produced those logs.
    """

    grid_shape = (3, 8)
    n_agents = 2
    init_agent_pos = {0: [0, 6], 1: [2, 6]}
    agent_reward = {0: {LEMON: -10, APPLE: 10}, 1: {LEMON: -1, APPLE: 1}}

    def __init__(self, full_observable=False, step_cost=-0.01, max_steps=100, done_rule="apples"):
        self.full_observable = full_observable
        self._step_cost = step_cost
        self._max_steps = max_steps
        self._done_rule = done_rule
        d = 2 + 3 * 3 * 5
        if full_observable:
            d *= self.n_agents
        self._obs_low, self._obs_high = [0.0] * d, [1.0] * d
        self.observation_space = [_Box(self._obs_low, self._obs_high) for _ in range(self.n_agents)]
        self.action_space = [_Discrete(5) for _ in range(self.n_agents)]
        self.steps_beyond_done = None
        self._full_obs = None

    # -- grid ---------------------------------------------------------------------------------
    def _create_grid(self):
        rows, cols = self.grid_shape
        grid = []
        for r in range(rows):
            apple_parity = 0 if r % 2 == 0 else 1
            grid.append([APPLE if c % 2 == apple_parity else LEMON for c in range(cols - 2)]
                        + [EMPTY, EMPTY])
        return grid

    def _is_valid(self, pos):
        return 0 <= pos[0] < self.grid_shape[0] and 0 <= pos[1] < self.grid_shape[1]

    def _has_no_agent(self, pos):
        return self._is_valid(pos) and AGENT not in self._full_obs[pos[0]][pos[1]]

    def _stamp_agent(self, agent_i):
        pr, pc = self.agent_prev_pos[agent_i]
        self._full_obs[pr][pc] = EMPTY
        r, c = self.agent_pos[agent_i]
        self._full_obs[r][c] = AGENT + str(agent_i + 1)

    # -- api ----------------------------------------------------------------------------------
    def reset(self):
        self.agent_pos = copy.deepcopy(self.init_agent_pos)
        self.agent_prev_pos = copy.deepcopy(self.init_agent_pos)
        self._full_obs = self._create_grid()
        for i in range(self.n_agents):
            self._stamp_agent(i)
        self._step_count = 0
        self.steps_beyond_done = None
        per_kind = ((self.grid_shape[1] - 2) // 2) * self.grid_shape[0]
        self._food_count = {LEMON: per_kind, APPLE: per_kind}
        self._agent_dones = [False] * self.n_agents
        return self.get_agent_obs()

    def get_agent_obs(self):
        obs = []
        rows, cols = self.grid_shape
        for i in range(self.n_agents):
            r0, c0 = self.agent_pos[i]
            o = [round(r0 / rows, 2), round(c0 / (cols - 1), 2)]
            for r in range(r0 - 1, r0 + 2):
                for c in range(c0 - 1, c0 + 2):
                    item = [0.0] * 5
                    if self._is_valid((r, c)):
                        cell = self._full_obs[r][c]
                        if cell in _CHANNEL:
                            item[_CHANNEL[cell]] = 1.0
                    o += item
            obs.append(o)
        if self.full_observable:
            flat = [x for o in obs for x in o]
            obs = [list(flat) for _ in range(self.n_agents)]
        return obs

    def step(self, agents_action):
        assert len(agents_action) == self.n_agents
        self._step_count += 1
        rewards = [self._step_cost for _ in range(self.n_agents)]
        for i, action in enumerate(agents_action):
            action = int(action)
            r, c = self.agent_pos[i]
            if action == 0:
                nxt = [r + 1, c]          # down
            elif action == 1:
                nxt = [r, c - 1]          # left
            elif action == 2:
                nxt = [r - 1, c]          # up
            elif action == 3:
                nxt = [r, c + 1]          # right
            elif action == 4:
                nxt = None                # no-op
            else:
                raise Exception("Action Not found!")
            self.agent_prev_pos[i] = self.agent_pos[i]
            if nxt is not None and self._has_no_agent(nxt):
                self.agent_pos[i] = nxt
            if self.agent_pos[i] != self.agent_prev_pos[i]:
                cell = self._full_obs[self.agent_pos[i][0]][self.agent_pos[i][1]]
                if cell in (LEMON, APPLE):
                    rewards[i] += self.agent_reward[i][cell]
                    self._food_count[cell] -= 1
                self._stamp_agent(i)

        if self._done_rule == "apples":
            food_gone = self._food_count[APPLE] == 0
        else:  # "all_food"
            food_gone = self._food_count[APPLE] == 0 and self._food_count[LEMON] == 0
        if self._step_count >= self._max_steps or food_gone:
            self._agent_dones = [True] * self.n_agents

        # cartpole-style guard: stepping an already finished episode yields zero rewards
        if self.steps_beyond_done is None and all(self._agent_dones):
            self.steps_beyond_done = 0
        elif self.steps_beyond_done is not None:
            self.steps_beyond_done += 1
            rewards = [0 for _ in range(self.n_agents)]

        return self.get_agent_obs(), rewards, list(self._agent_dones), {"food_count": dict(self._food_count)}

    def render(self, *a, **k):
        return None

    def close(self):
        return None


class SwitchRef:
    """ma-gym ``Switch{2,4}-v0``: 3x7 grid, one-cell corridor, agents swap sides (Appendix A)."""

    grid_shape = (3, 7)

    def __init__(self, full_observable=False, step_cost=-0.1, n_agents=2, max_steps=50, clock=True):
        assert 2 <= n_agents <= 4
        self.n_agents = n_agents
        self.full_observable = full_observable
        self._step_cost = step_cost
        self._max_steps = max_steps
        self._add_clock = clock
        cols = self.grid_shape[1]
        init = {0: [0, 1], 1: [0, cols - 2], 2: [2, 1], 3: [2, cols - 2]}
        final = {0: [0, cols - 1], 1: [0, 0], 2: [2, cols - 1], 3: [2, 0]}
        self.init_agent_pos = {i: init[i] for i in range(n_agents)}
        self.final_agent_pos = {i: final[i] for i in range(n_agents)}
        d = 2 + (1 if clock else 0)
        if full_observable:
            d *= n_agents
        self._obs_low, self._obs_high = [0.0] * d, [1.0] * d
        self.observation_space = [_Box(self._obs_low, self._obs_high) for _ in range(n_agents)]
        self.action_space = [_Discrete(5) for _ in range(n_agents)]
        self._base_grid = self._create_grid()

    def _create_grid(self):
        rows, cols = self.grid_shape
        grid = [[-1] * cols for _ in range(rows)]       # walls everywhere ...
        for c in range(cols):
            grid[rows // 2][c] = 0                      # ... except the corridor row
        for r in range(rows):
            for c in (0, 1, cols - 2, cols - 1):
                grid[r][c] = 0                          # ... and the two end rooms
        return grid

    def reset(self):
        self.agent_pos = copy.deepcopy(self.init_agent_pos)
        self._full_obs = self._create_grid()
        for i in range(self.n_agents):
            r, c = self.agent_pos[i]
            self._full_obs[r][c] = i + 1
        self._step_count = 0
        self._agent_dones = [False] * self.n_agents
        return self.get_agent_obs()

    def get_agent_obs(self):
        rows, cols = self.grid_shape
        obs = []
        for i in range(self.n_agents):
            r, c = self.agent_pos[i]
            o = [round(r / (rows - 1), 2), round(c / (cols - 1), 2)]
            if self._add_clock:
                o += [self._step_count / self._max_steps]
            obs.append(o)
        if self.full_observable:
            flat = [x for o in obs for x in o]
            obs = [list(flat) for _ in range(self.n_agents)]
        return obs

    def _is_cell_vacant(self, pos):
        rows, cols = self.grid_shape
        return 0 <= pos[0] < rows and 0 <= pos[1] < cols and self._full_obs[pos[0]][pos[1]] == 0

    def step(self, agents_action):
        assert len(agents_action) == self.n_agents
        self._step_count += 1
        rewards = [self._step_cost for _ in range(self.n_agents)]
        for i, action in enumerate(agents_action):
            if self._agent_dones[i]:
                continue
            action = int(action)
            r, c = self.agent_pos[i]
            if action == 0:
                nxt = [r + 1, c]
            elif action == 1:
                nxt = [r, c - 1]
            elif action == 2:
                nxt = [r - 1, c]
            elif action == 3:
                nxt = [r, c + 1]
            elif action == 4:
                nxt = None
            else:
                raise Exception("Action Not found!")
            if nxt is not None and self._is_cell_vacant(nxt):
                self.agent_pos[i] = nxt
                self._full_obs[r][c] = 0
                self._full_obs[nxt[0]][nxt[1]] = i + 1
            self._agent_dones[i] = self.agent_pos[i] == self.final_agent_pos[i]
            if self._agent_dones[i]:
                rewards[i] = 5
        if self._step_count >= self._max_steps:
            self._agent_dones = [True] * self.n_agents
        return self.get_agent_obs(), rewards, list(self._agent_dones), {}

    def render(self, *a, **k):
        return None

    def close(self):
        return None


def make(id, **kwargs):
    """``gym.make`` look-alike for the ids the reference uses (``vdn/main.py:61``)."""
    name = id.split(":")[-1]
    if name == "Checkers-v0":
        kw = dict(max_steps=100, step_cost=-0.01)
        kw.update(kwargs)
        return CheckersRef(**kw)
    if name in ("Switch2-v0", "Switch4-v0"):
        kw = dict(n_agents=int(name[6]), full_observable=False, step_cost=-0.1, max_steps=50)
        kw.update(kwargs)
        return SwitchRef(**kw)
    raise Exception(f"unknown env id {id}")
