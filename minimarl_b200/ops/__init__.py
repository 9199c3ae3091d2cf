from .gae import gae_returns  # noqa: F401
from .td import q_select, qmix_td_loss  # noqa: F401
