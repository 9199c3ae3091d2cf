"""N lockstep ma-gym grid worlds on one GPU (SURVEY.md 8b "new batched surface").

``BatchedCheckers`` / ``BatchedSwitch`` hold the env state of ``n_envs`` independent instances as
structure-of-arrays uint32 planes in HBM and step all of them with one kernel launch
(``mm_checkers_step`` / ``mm_switch_step``).  Semantics per env are those of
``gym.make("ma_gym:Checkers-v0" | "ma_gym:Switch2-v0", ...)`` as called by the reference
(``vdn/main.py:61-64``, ``qmix/main.py:66-71``, ``mappo/main.py:41-46``).
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from .spaces import Box, Discrete, MultiAgentSpace


def _as_actions(actions, n_envs, n_agents, device):
    """[N,A] int32 CUDA view of whatever the reference's callers pass (float tensors, [A,1] arrays...)."""
    if not torch.is_tensor(actions):
        actions = torch.as_tensor(np.asarray(actions))
    actions = actions.reshape(n_envs, n_agents)
    if actions.device != device or actions.dtype != torch.int32:
        actions = actions.to(device=device, dtype=torch.int32, non_blocking=True)
    return actions.contiguous()


class _BatchedBase:
    n_agents: int
    obs_dim: int

    def _alloc_common(self, n_envs, device, words):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("minimarl_b200 envs live on a CUDA device (no CPU fallback)")
        self.n_envs = int(n_envs)
        self.state = torch.zeros((words, self.n_envs), dtype=torch.int32, device=self.device)
        self.bad_action = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._host = None  # pinned staging for step_host

    def _spaces(self):
        d = self.obs_dim
        self._obs_low = np.zeros(d)
        self._obs_high = np.ones(d)
        self.observation_space = MultiAgentSpace(
            [Box(self._obs_low, self._obs_high) for _ in range(self.n_agents)])
        self.action_space = MultiAgentSpace([Discrete(5) for _ in range(self.n_agents)])

    def new_outputs(self, T=None):
        lead = (self.n_envs,) if T is None else (T, self.n_envs)
        obs = torch.empty(lead + (self.n_agents, self.obs_dim), dtype=torch.float32, device=self.device)
        rew = torch.empty(lead + (self.n_agents,), dtype=torch.float32, device=self.device)
        done = torch.empty(lead + (self.n_agents,), dtype=torch.uint8, device=self.device)
        return obs, rew, done

    def check_actions(self):
        """Raises like the reference env if any action so far was outside [0, 5).  Synchronises."""
        if int(self.bad_action.item()) != 0:
            self.bad_action.zero_()
            raise Exception("Action Not found!")

    def step(self, actions, out=None):
        """Lockstep ``env.step``: returns (obs [N,A,D] f32, rewards [N,A] f32, dones [N,A] bool)."""
        actions = _as_actions(actions, self.n_envs, self.n_agents, self.device)
        obs, rew, done = out if out is not None else self.new_outputs()
        self.step_into(actions, obs, rew, done)
        return obs, rew, done.view(torch.bool)

    # -- host-buffer path: what a caller holding numpy arrays pays (bench "e2e") ----------------
    def _host_staging(self):
        if self._host is None:
            n, a, d = self.n_envs, self.n_agents, self.obs_dim
            self._host = dict(
                act=torch.empty((n, a), dtype=torch.int32).pin_memory(),
                obs=torch.empty((n, a, d), dtype=torch.float32).pin_memory(),
                rew=torch.empty((n, a), dtype=torch.float32).pin_memory(),
                done=torch.empty((n, a), dtype=torch.uint8).pin_memory(),
                d_act=torch.empty((n, a), dtype=torch.int32, device=self.device),
                d_out=self.new_outputs(),
            )
        return self._host

    def step_host(self, actions_np):
        """``step`` with HOST numpy in/out: H2D of the actions, kernel, D2H of obs/rewards/dones."""
        h = self._host_staging()
        h["act"].numpy()[...] = np.asarray(actions_np).reshape(self.n_envs, self.n_agents)
        h["d_act"].copy_(h["act"], non_blocking=True)
        obs, rew, done = h["d_out"]
        self.step_into(h["d_act"], obs, rew, done)
        h["obs"].copy_(obs, non_blocking=True)
        h["rew"].copy_(rew, non_blocking=True)
        h["done"].copy_(done, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return h["obs"].numpy(), h["rew"].numpy(), h["done"].numpy().view(np.bool_)

    @property
    def h2d_bytes_per_step(self):
        return self.n_envs * self.n_agents * 4

    @property
    def d2h_bytes_per_step(self):
        return self.n_envs * self.n_agents * (self.obs_dim * 4 + 4 + 1)


class BatchedCheckers(_BatchedBase):
    """``n_envs`` independent ``Checkers-v0`` boards stepped in lockstep by one kernel."""

    n_agents = 2
    grid_shape = (3, 8)

    def __init__(self, n_envs, full_observable=False, max_steps=100, step_cost=-0.01, device="cuda",
                 auto_reset=False):
        self._alloc_common(n_envs, device, 3)
        self.full_observable = bool(full_observable)
        self.max_steps = int(max_steps)
        self.step_cost = step_cost
        self.obs_dim = 47 * (2 if self.full_observable else 1)
        self._spaces()
        cfg = _lib.CheckersCfg()
        cfg.n_envs, cfg.max_steps = self.n_envs, self.max_steps
        cfg.full_observable, cfg.auto_reset = int(self.full_observable), int(bool(auto_reset))
        # float32 of the Python-double expressions the reference env evaluates
        food = {0: (-10, 10), 1: (-1, 1)}
        self.reward_table = {}
        for a in range(2):
            vals = [step_cost, step_cost + food[a][0], step_cost + food[a][1]]
            for k, v in enumerate(vals):
                cfg.reward[a][k] = np.float32(v)
                self.reward_table[float(np.float32(v))] = v
        self.reward_table[0.0] = 0
        for r in range(3):
            cfg.pos_row[r] = np.float32(round(r / 3, 2))
        for c in range(8):
            cfg.pos_col[c] = np.float32(round(c / 7, 2))
        self.cfg = cfg
        self.steps_done = 0

    def reset(self, mask=None, out=None):
        """``env.reset()`` of every env (or of the envs where ``mask`` is True); returns obs [N,A,D]."""
        obs = out if out is not None else self.new_outputs()[0]
        m = None
        if mask is not None:
            m = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        _lib.check(_lib.lib().mm_checkers_reset(_lib.ptr(self.state), _lib.ptr(m), _lib.ptr(obs),
                                                C.byref(self.cfg), _lib.stream()))
        return obs

    def step_into(self, actions, obs, rewards, dones):
        """Zero-copy step: outputs go straight into (slices of) the caller's rollout/replay buffer."""
        _lib.check(_lib.lib().mm_checkers_step(
            _lib.ptr(self.state), _lib.ptr(actions), _lib.ptr(obs), _lib.ptr(rewards), _lib.ptr(dones),
            _lib.ptr(self.bad_action), C.byref(self.cfg), _lib.stream()))

    def rollout(self, T, actions=None, seed=0, step_offset=0, out=None, actions_out=None,
                want_obs=True):
        """T lockstep steps in ONE launch.  ``actions`` [T,N,A] int32, or None to draw uniform random
        actions on the device (written to ``actions_out`` if given).  Returns (obs [T,N,A,D] -- the
        obs after each step --, rewards [T,N,A], dones [T,N,A])."""
        if actions is not None:
            actions = actions.to(device=self.device, dtype=torch.int32).reshape(
                T, self.n_envs, self.n_agents).contiguous()
        obs, rew, done = out if out is not None else self.new_outputs(T)
        if not want_obs:
            obs = None
        _lib.check(_lib.lib().mm_checkers_rollout(
            _lib.ptr(self.state), _lib.ptr(actions), _lib.ptr(actions_out), _lib.ptr(obs), _lib.ptr(rew),
            _lib.ptr(done), _lib.ptr(self.bad_action), int(T), int(seed), int(step_offset),
            C.byref(self.cfg), _lib.stream()))
        return obs, rew, done.view(torch.bool)

    # -- host-buffer rollout: numpy in / numpy out, copies pipelined against the kernel --------------
    def host_buffers(self, T, chunk=8):
        """Page-locked numpy arrays for ``rollout_host``: actions [T,N,A] int32 (fill it), obs [T,N,A,D] f32,
        rewards [T,N,A] f32, dones [T,N,A] bool, plus the device staging for two chunks of ``chunk`` steps."""
        key = (int(T), int(chunk))
        if getattr(self, "_hostT", None) is None or self._hostT["key"] != key:
            n, a, d = self.n_envs, self.n_agents, self.obs_dim
            pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
            dev = lambda shape, dt: torch.empty(shape, dtype=dt, device=self.device)
            self._hostT = dict(
                key=key,
                act=pin((T, n, a), torch.int32), obs=pin((T, n, a, d), torch.float32),
                rew=pin((T, n, a), torch.float32), done=pin((T, n, a), torch.uint8),
                s_act=dev((2, chunk, n, a), torch.int32), s_obs=dev((2, chunk, n, a, d), torch.float32),
                s_rew=dev((2, chunk, n, a), torch.float32), s_done=dev((2, chunk, n, a), torch.uint8),
                codes=pin((T, n, a), torch.int64), s_codes=dev((2, chunk, n, a), torch.int64),
                copy_stream=torch.cuda.Stream(device=self.device), transport=None, trial={})
        return self._hostT

    def rollout_host(self, actions_np, chunk=8, want_obs=True, transport="auto", n_threads=0):
        """T lockstep steps for a caller holding HOST arrays (what the reference's runners hold):
        ``actions_np`` [T,N,A] ints -> numpy (obs [T,N,A,D], rewards [T,N,A], dones [T,N,A] bool), views of
        page-locked buffers that the next call overwrites.  H2D of the actions, the kernels and the D2H of the
        outputs are pipelined chunk by chunk; returns after the last output landed.
        Pass the ``host_buffers(T)["act"].numpy()`` array itself to skip the staging copy of the actions.

        ``transport``: "direct" moves the float32 observations over PCIe as they are (``mm_checkers_rollout_host``);
        "compact" moves one 64-bit word per agent-step and expands it on ``n_threads`` host threads into the same arrays,
        bit-identical (``mm_checkers_rollout_host_compact``); "auto" times one call of each on this env / host and keeps the
        faster (which one wins depends on how many ranks share the host's memory bandwidth)."""
        T = int(np.shape(actions_np)[0])
        h = self.host_buffers(T, chunk)
        act = h["act"].numpy()
        if not (isinstance(actions_np, np.ndarray) and np.shares_memory(actions_np, act)):
            act[...] = np.asarray(actions_np).reshape(T, self.n_envs, self.n_agents)
        cur = torch.cuda.current_stream(self.device)
        h["copy_stream"].wait_stream(cur)
        if transport == "auto":
            if not want_obs:
                transport = "direct"                         # without the observations there is nothing to compress
            elif h["transport"] is not None:
                transport = h["transport"]
            else:
                transport = "direct" if "direct" not in h["trial"] else "compact"
                import time
                cur.synchronize()
                t0 = time.perf_counter()
                out = self.rollout_host(actions_np, chunk, want_obs, transport, n_threads)
                h["trial"][transport] = time.perf_counter() - t0
                if len(h["trial"]) == 2:
                    h["transport"] = min(h["trial"], key=h["trial"].get)
                return out
        if transport == "compact":
            _lib.check(_lib.lib().mm_checkers_rollout_host_compact(
                _lib.ptr(self.state), h["act"].data_ptr(), h["obs"].data_ptr() if want_obs else None, h["rew"].data_ptr(),
                h["done"].data_ptr(), h["codes"].data_ptr(), _lib.ptr(h["s_act"]), _lib.ptr(h["s_codes"]),
                _lib.ptr(self.bad_action), T, int(chunk), int(n_threads), C.byref(self.cfg), cur.cuda_stream,
                h["copy_stream"].cuda_stream))
            cur.synchronize()
            return (h["obs"].numpy() if want_obs else None), h["rew"].numpy(), h["done"].numpy().view(np.bool_)
        if transport != "direct":
            raise Exception(f"unknown transport {transport!r}")
        _lib.check(_lib.lib().mm_checkers_rollout_host(
            _lib.ptr(self.state), h["act"].data_ptr(), h["obs"].data_ptr() if want_obs else None,
            h["rew"].data_ptr(), h["done"].data_ptr(), _lib.ptr(h["s_act"]),
            _lib.ptr(h["s_obs"]) if want_obs else None, _lib.ptr(h["s_rew"]), _lib.ptr(h["s_done"]),
            _lib.ptr(self.bad_action), T, int(chunk), C.byref(self.cfg), cur.cuda_stream,
            h["copy_stream"].cuda_stream))
        cur.synchronize()
        return (h["obs"].numpy() if want_obs else None), h["rew"].numpy(), h["done"].numpy().view(np.bool_)

    def rollout_host_bytes(self, T, transport="direct"):
        """(H2D, D2H) bytes one ``rollout_host`` call of T steps moves over PCIe."""
        if transport == "compact":
            return T * self.h2d_bytes_per_step, T * self.n_envs * self.n_agents * 8
        return T * self.h2d_bytes_per_step, T * self.d2h_bytes_per_step

    # -- state decoding (tests, eval, render shims) -------------------------------------------------
    def positions(self):
        """[N, A, 2] (row, col) of every agent."""
        cells = self.state[1].to(torch.int64)
        per = torch.stack([(cells >> (8 * a)) & 0xFF for a in range(self.n_agents)], dim=1)
        return torch.stack([per >> 3, per & 7], dim=-1)

    def food_bitmap(self):
        return self.state[0].to(torch.int64) & 0xFFFFFF

    def step_count(self):
        return self.state[2].clone()


class BatchedSwitch(_BatchedBase):
    """``n_envs`` independent ``Switch{2,4}-v0`` corridors stepped in lockstep by one kernel."""

    grid_shape = (3, 7)

    def __init__(self, n_envs, n_agents=2, full_observable=False, max_steps=50, step_cost=-0.1,
                 clock=True, device="cuda"):
        if n_agents not in (2, 4):
            raise Exception("Switch is built for 2 or 4 agents (Switch2-v0 / Switch4-v0)")
        self.n_agents = int(n_agents)
        self._alloc_common(n_envs, device, 3)
        self.full_observable = bool(full_observable)
        self.max_steps = int(max_steps)
        self.step_cost = step_cost
        d = 2 + (1 if clock else 0)
        self.obs_dim = d * (self.n_agents if self.full_observable else 1)
        self._spaces()
        cfg = _lib.SwitchCfg()
        cfg.n_envs, cfg.n_agents, cfg.max_steps = self.n_envs, self.n_agents, self.max_steps
        cfg.full_observable, cfg.clock = int(self.full_observable), int(bool(clock))
        cfg.step_cost, cfg.goal_reward = np.float32(step_cost), np.float32(5)
        for r in range(3):
            cfg.pos_row[r] = np.float32(round(r / 2, 2))
        for c in range(7):
            cfg.pos_col[c] = np.float32(round(c / 6, 2))
        self.cfg = cfg
        self.reward_table = {float(np.float32(step_cost)): step_cost, 5.0: 5}

    def reset(self, mask=None, out=None):
        obs = out if out is not None else self.new_outputs()[0]
        m = None
        if mask is not None:
            m = mask.to(device=self.device, dtype=torch.uint8).contiguous()
        _lib.check(_lib.lib().mm_switch_reset(_lib.ptr(self.state), _lib.ptr(m), _lib.ptr(obs),
                                              C.byref(self.cfg), _lib.stream()))
        return obs

    def step_into(self, actions, obs, rewards, dones):
        _lib.check(_lib.lib().mm_switch_step(
            _lib.ptr(self.state), _lib.ptr(actions), _lib.ptr(obs), _lib.ptr(rewards), _lib.ptr(dones),
            _lib.ptr(self.bad_action), C.byref(self.cfg), _lib.stream()))

    def rollout(self, T, actions=None, seed=0, step_offset=0, out=None, actions_out=None, want_obs=True):
        """T lockstep steps in ONE launch (``mm_switch_rollout``); arguments and results as ``BatchedCheckers.rollout``."""
        if actions is not None:
            actions = actions.to(device=self.device, dtype=torch.int32).reshape(T, self.n_envs, self.n_agents).contiguous()
        obs, rew, done = out if out is not None else self.new_outputs(T)
        if not want_obs:
            obs = None
        _lib.check(_lib.lib().mm_switch_rollout(
            _lib.ptr(self.state), _lib.ptr(actions), _lib.ptr(actions_out), _lib.ptr(obs), _lib.ptr(rew), _lib.ptr(done),
            _lib.ptr(self.bad_action), int(T), int(seed), int(step_offset), C.byref(self.cfg), _lib.stream()))
        return obs, rew, done.view(torch.bool)

    def positions(self):
        cells = self.state[0].to(torch.int64)
        per = torch.stack([(cells >> (8 * a)) & 0xFF for a in range(self.n_agents)], dim=1)
        return torch.stack([per >> 3, per & 7], dim=-1)
