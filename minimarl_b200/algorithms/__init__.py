from .qmix import QMix, QMixPolicy, VDN  # noqa: F401
from .mappo import R_MAPPO, R_Actor, R_Critic, R_MAPPOPolicy, ValueNorm  # noqa: F401
