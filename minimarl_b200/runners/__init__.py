from .offpolicy import MAGYM_Runner as OffPolicyRunner  # noqa: F401
from .mappo import MAGYM_Runner as MappoRunner  # noqa: F401
from .vdn_main import run_qmix, run_vdn  # noqa: F401
