from .segment_tree import MinSegmentTree, SumSegmentTree  # noqa: F401
from .rec_buffer import PrioritizedRecReplayBuffer, RecPolicyBuffer, RecReplayBuffer  # noqa: F401
from .shared_buffer import SharedReplayBuffer  # noqa: F401
