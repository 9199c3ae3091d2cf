from .qmix import QMix, QMixPolicy, VDN  # noqa: F401
from .mappo import R_MAPPO, R_Actor, R_Critic, R_MAPPOPolicy, ValueNorm  # noqa: F401
from .toplevel import Mix_Net, Prioritized_Experience_Replay, Q_Net, SumTree, Target_Dqn, Test, Train_dqn, cal_td_error  # noqa: F401
