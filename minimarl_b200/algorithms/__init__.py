from .qmix import QMix, QMixPolicy, VDN  # noqa: F401
