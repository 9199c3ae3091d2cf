"""GAE / discounted-return reverse scan on the GPU (``mm_gae_returns``)."""
from .. import _lib


def gae_returns(rewards, value_preds, masks, returns, gamma, gae_lambda, use_gae=True, denorm=None,
                advantages=None):
    """In-place ``SharedReplayBuffer.compute_returns`` (mappo/runner/shared/shared_buffer.py:131-157).

    rewards [T,...], value_preds / masks / returns [T+1,...] float32 CUDA tensors sharing the
    trailing shape; ``value_preds[T]`` must already hold ``next_value``.  ``denorm`` is None or
    (sqrt(var), mean) of the value normaliser.  ``advantages`` (optional, [T,...]) receives
    ``returns[:-1] - denorm(value_preds[:-1])`` (mappo/algorithms/ramppo_network.py:225-230).
    """
    T = rewards.shape[0]
    rows = rewards[0].numel()
    assert masks.shape[0] == T + 1 and returns.shape[0] == T + 1
    scale, shift = (1.0, 0.0) if denorm is None else (float(denorm[0]), float(denorm[1]))
    _lib.check(_lib.lib().mm_gae_returns(
        _lib.ptr(rewards), _lib.ptr(value_preds), _lib.ptr(masks), _lib.ptr(returns),
        _lib.ptr(advantages), int(T), int(rows), float(gamma), float(gae_lambda), int(bool(use_gae)),
        int(denorm is not None), scale, shift, _lib.stream()))
    return returns
