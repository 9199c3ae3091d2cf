"""ctypes front end of oracle/magym_ref.c (TEST INFRASTRUCTURE; see that file's header).

Numpy-in / numpy-out lockstep batches of the C restatement, used by the GPU parity tests at sizes
the pure-Python restatement (oracle/ma_gym_ref.py) is too slow for, and by bench.py's CPU baseline.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmagym_ref.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "magym_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmagym_ref.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.ck_env_size.restype = C.c_int
        _lib.sw_env_size.restype = C.c_int
        _lib.ck_batch_step.restype = C.c_int
        _lib.sw_batch_step.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class CheckersBatchRef:
    n_agents = 2

    def __init__(self, n_envs, full_observable=False, max_steps=100, step_cost=-0.01, auto_reset=False):
        self.n, self.full, self.max_steps, self.step_cost = n_envs, int(full_observable), max_steps, step_cost
        self.auto_reset = int(auto_reset)
        self.obs_dim = 47 * (2 if full_observable else 1)
        self._envs = np.zeros(n_envs * lib().ck_env_size(), dtype=np.uint8)

    def reset(self):
        obs = np.empty((self.n, 2, self.obs_dim), dtype=np.float32)
        lib().ck_batch_reset(_p(self._envs), self.n, self.full, _p(obs))
        return obs

    def step(self, actions):
        actions = np.ascontiguousarray(actions, dtype=np.int32).reshape(self.n, 2)
        obs = np.empty((self.n, 2, self.obs_dim), dtype=np.float32)
        rew = np.empty((self.n, 2), dtype=np.float32)
        done = np.empty((self.n, 2), dtype=np.uint8)
        rc = lib().ck_batch_step(_p(self._envs), self.n, _p(actions), C.c_double(self.step_cost),
                                 self.max_steps, self.full, self.auto_reset, _p(obs), _p(rew), _p(done))
        if rc != 0:
            raise Exception("Action Not found!")
        return obs, rew, done.astype(bool)

    def step_no_obs(self, actions, rew, done):
        """Timing helper: transitions only, outputs into caller buffers."""
        return lib().ck_batch_step(_p(self._envs), self.n, _p(actions), C.c_double(self.step_cost),
                                   self.max_steps, self.full, self.auto_reset, None, _p(rew), _p(done))

    def export(self):
        pos = np.empty((self.n, 2, 2), dtype=np.int32)
        steps = np.empty(self.n, dtype=np.int32)
        food = np.empty((self.n, 2), dtype=np.int32)
        lib().ck_batch_export(_p(self._envs), self.n, _p(pos), _p(steps), _p(food))
        return pos, steps, food


class SwitchBatchRef:
    def __init__(self, n_envs, n_agents=2, full_observable=False, max_steps=50, step_cost=-0.1, clock=True):
        self.n, self.n_agents, self.full = n_envs, n_agents, int(full_observable)
        self.max_steps, self.step_cost, self.clock = max_steps, step_cost, int(clock)
        d = 2 + (1 if clock else 0)
        self.obs_dim = d * (n_agents if full_observable else 1)
        self._envs = np.zeros(n_envs * lib().sw_env_size(), dtype=np.uint8)

    def reset(self):
        obs = np.empty((self.n, self.n_agents, self.obs_dim), dtype=np.float32)
        lib().sw_batch_reset(_p(self._envs), self.n, self.n_agents, self.full, self.clock, self.max_steps,
                             _p(obs))
        return obs

    def step(self, actions):
        actions = np.ascontiguousarray(actions, dtype=np.int32).reshape(self.n, self.n_agents)
        obs = np.empty((self.n, self.n_agents, self.obs_dim), dtype=np.float32)
        rew = np.empty((self.n, self.n_agents), dtype=np.float32)
        done = np.empty((self.n, self.n_agents), dtype=np.uint8)
        rc = lib().sw_batch_step(_p(self._envs), self.n, _p(actions), C.c_double(self.step_cost),
                                 self.max_steps, self.full, self.clock, _p(obs), _p(rew), _p(done))
        if rc != 0:
            raise Exception("Action Not found!")
        return obs, rew, done.astype(bool)

    def export(self):
        pos = np.empty((self.n, self.n_agents, 2), dtype=np.int32)
        lib().sw_batch_export(_p(self._envs), self.n, _p(pos))
        return pos
