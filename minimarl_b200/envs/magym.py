"""Single-env adaptor with the list-based gym API the reference mains call.

``make("ma_gym:Checkers-v0", full_observable=True, max_steps=100, step_cost=-0.01)`` returns an
object the unmodified ``vdn/main.py`` / ``qmix/main.py`` / runners can drive: ``reset()`` gives a
list (per agent) of float lists, ``step(actions)`` takes whatever those callers pass (a float
tensor ``action[0]`` ``vdn/main.py:93``, an ``[A,1]`` int array
``offpolicy/runner/shared/magym_runner.py:94-97``, an ``[A,1]`` float array
``mappo/runner/shared/magym_runner.py:53-54``) and returns ``(obs, rewards, dones, info)``.
It is a 1-env ``BatchedCheckers``/``BatchedSwitch``: the arithmetic runs in the CUDA kernel.
"""
import numpy as np
import torch

from .batched import BatchedCheckers, BatchedSwitch


class MaGymEnv:
    def __init__(self, batched):
        self._env = batched
        self.n_agents = batched.n_agents
        self.observation_space = batched.observation_space
        self.action_space = batched.action_space
        self._obs_high, self._obs_low = batched._obs_high, batched._obs_low
        self._out = batched.new_outputs()

    def reset(self):
        obs = self._env.reset(out=self._out[0])
        return obs[0].cpu().tolist()

    def step(self, agents_action):
        if torch.is_tensor(agents_action):
            acts = agents_action.detach().reshape(-1).to("cpu").tolist()
        else:
            acts = np.asarray(agents_action).reshape(-1).tolist()
        assert len(acts) == self.n_agents
        acts = [int(a) for a in acts]
        if any(a < 0 or a > 4 for a in acts):
            raise Exception("Action Not found!")
        obs, rew, done = self._env.step(torch.tensor([acts], dtype=torch.int32), out=self._out)
        table = self._env.reward_table
        rewards = [table.get(r, r) for r in rew[0].cpu().tolist()]
        return obs[0].cpu().tolist(), rewards, done[0].cpu().tolist(), {}

    def render(self, mode="human", *args, **kwargs):
        """Text picture of the board decoded from the device state (the reference only calls ``env.render()`` on its
        eval path, offpolicy/runner/shared/magym_runner.py:72-73): ``R`` apple, ``Y`` lemon, ``1``/``2``.. agents.
        Printed for ``mode="human"``, returned as a string otherwise."""
        env = self._env
        rows, cols = env.grid_shape
        grid = [["." for _ in range(cols)] for _ in range(rows)]
        if hasattr(env, "food_bitmap"):
            food = int(env.food_bitmap()[0].item())
            for r in range(rows):
                for c in range(cols):
                    if food >> (r * 8 + c) & 1:
                        grid[r][c] = "R" if (r + c) % 2 == 0 else "Y"
        for i, (r, c) in enumerate(env.positions()[0].cpu().tolist()):
            grid[r][c] = str(i + 1)
        text = "\n".join(" ".join(row) for row in grid)
        if mode == "human":
            print(text)
            return None
        return text

    def close(self):
        return None

    def seed(self, n=None):
        return [n]


def make(id, device="cuda", **kwargs):
    """``gym.make`` for the two ids the reference uses."""
    name = id.split(":")[-1]
    if name == "Checkers-v0":
        kw = dict(max_steps=100, step_cost=-0.01, full_observable=False)
        kw.update(kwargs)
        return MaGymEnv(BatchedCheckers(1, device=device, **kw))
    if name in ("Switch2-v0", "Switch4-v0"):
        kw = dict(n_agents=int(name[6]), full_observable=False, step_cost=-0.1, max_steps=50)
        kw.update(kwargs)
        return MaGymEnv(BatchedSwitch(1, device=device, **kw))
    raise Exception(f"unknown env id {id}")
