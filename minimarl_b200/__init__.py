"""minimarl_b200: B200-native (sm_100a) hot path of singforai/mini-marl.

Lockstep multi-agent env step (ma-gym Checkers-v0 / Switch2-v0 semantics), MAPPO rollout buffer +
GAE, and the QMIX/VDN learner path, as hand-written CUDA kernels behind a C ABI
(``include/minimarl_b200.h``).  The classes here mirror the reference's Python surface
(SURVEY.md 8b).  There is no CPU fallback: every op raises if ``libminimarl_b200.so`` is missing.
"""
from . import _lib  # noqa: F401
from .envs import BatchedCheckers, BatchedSwitch, make  # noqa: F401

__all__ = ["BatchedCheckers", "BatchedSwitch", "make"]
