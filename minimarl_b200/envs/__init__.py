from .batched import BatchedCheckers, BatchedSwitch  # noqa: F401
from .magym import MaGymEnv, make  # noqa: F401
