"""numpy restatement of the MAPPO return/GAE computation -- TEST INFRASTRUCTURE.

Follows ``SharedReplayBuffer.compute_returns`` (mappo/runner/shared/shared_buffer.py:131-157) with
``ValueNorm.denormalize`` (mappo/utils/valuenorm.py:67-78) and the advantage line of
``R_MAPPO.train`` (mappo/algorithms/ramppo_network.py:225-230), keeping the reference's dtypes:
float32 arrays, Python-float scalars (weak in numpy promotion, so products with float32 arrays stay
float32) and a float64 ``gae`` accumulator created by ``np.zeros``.

Pinned: ``tests/golden/gae_*.npz`` were produced by the reference's own ``SharedReplayBuffer`` +
``ValueNorm`` (script ``tests/golden/make_golden.py``); ``tests/test_oracle_gae.py`` checks this file
against them bit for bit.
"""
import numpy as np


def denormalize(x, scale, shift):
    """torch float32: ``input * sqrt(var) + mean`` (two roundings)."""
    x = np.asarray(x, dtype=np.float32)
    return (x * np.float32(scale) + np.float32(shift)).astype(np.float32)


def compute_returns(rewards, value_preds, masks, gamma, gae_lambda, use_gae=True, denorm=None,
                    returns_last=None):
    """rewards [T,...], value_preds/masks [T+1,...] float32 -> returns [T+1,...] float32.

    ``denorm`` is None or (scale, shift) = (sqrt(var), mean) of the value normaliser.
    """
    rewards = np.asarray(rewards, dtype=np.float32)
    value_preds = np.asarray(value_preds, dtype=np.float32)
    masks = np.asarray(masks, dtype=np.float32)
    T = rewards.shape[0]
    returns = np.zeros_like(value_preds)
    if use_gae:
        gae = np.zeros(rewards.shape[1:])                    # float64, like the reference
        for step in reversed(range(T)):
            if denorm is not None:
                v_next = denormalize(value_preds[step + 1], *denorm)
                v = denormalize(value_preds[step], *denorm)
            else:
                v_next, v = value_preds[step + 1], value_preds[step]
            delta = rewards[step] + gamma * v_next * masks[step + 1] - v      # float32
            gae = delta + gamma * gae_lambda * masks[step + 1] * gae          # float64
            returns[step] = gae + v
    else:
        returns[-1] = returns_last
        for step in reversed(range(T)):
            returns[step] = rewards[step] + returns[step + 1] * gamma * masks[step + 1]
    return returns


def advantages(returns, value_preds, denorm=None):
    v = value_preds[:-1] if denorm is None else denormalize(value_preds[:-1], *denorm)
    return returns[:-1] - v
