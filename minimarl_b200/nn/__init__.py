from .agent_q import AgentQFunction, MLPBase, RNNBase, ACTLayer, agentq_forward, pack_agentq  # noqa: F401
from .mixer import QMixer, VDNMixer, pack_qmixer  # noqa: F401
