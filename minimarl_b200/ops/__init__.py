from .gae import gae_returns  # noqa: F401
