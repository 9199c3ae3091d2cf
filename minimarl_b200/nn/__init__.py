from .agent_q import AgentQFunction, MLPBase, RNNBase, ACTLayer, agentq_forward, pack_agentq, pack_agentq_tc  # noqa: F401
from .mixer import QMixer, VDNMixer, pack_qmixer  # noqa: F401
from .layernorm import LayerNorm  # noqa: F401
