"""PPO actor / critic losses as one forward and one backward kernel each (``mm_ppo_actor_loss_*`` / ``mm_ppo_value_loss_*``).

They compute exactly what ``R_MAPPO.ppo_update`` / ``cal_value_loss`` compute with ~40 elementwise PyTorch calls
(mappo/algorithms/ramppo_network.py:70-171): clipped surrogate, masked means, Categorical entropy, clipped Huber / MSE
value error against (ValueNorm-)normalised returns; autograd's tie rules for min / max / clamp included.
"""
import torch

from .. import _lib


def _rows(xs, R):
    """Per-row inputs: dense [R] vectors, or tables addressed like recurrent_generator's chunks (``ChunkRows`` objects, all
    sharing one row map).  -> (list of float32 tensors, first, step_rows, n_chunks)."""
    lazy = [x for x in xs if x is not None and hasattr(x, "first")]
    if not lazy:
        return [None if x is None else x.reshape(R).to(torch.float32).contiguous() for x in xs], None, 0, 0
    ref = lazy[0]
    assert len(lazy) == sum(x is not None for x in xs) and all(x.first is ref.first and x.step_rows == ref.step_rows for x in lazy)
    assert ref.L * ref.first.numel() == R
    return [None if x is None else x.table.reshape(-1) for x in xs], ref.first, ref.step_rows, ref.first.numel()


class _ActorLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, actions, old_logp, adv, active, clip):
        R, NA = logits.shape
        logits = logits.contiguous()
        (actions, old_logp, adv, active), first, step, n_chunks = _rows((actions, old_logp, adv, active), R)
        sums = torch.zeros(4, dtype=torch.float64, device=logits.device)
        _lib.check(_lib.lib().mm_ppo_actor_loss_fwd(_lib.ptr(logits), _lib.ptr(actions), _lib.ptr(old_logp), _lib.ptr(adv),
                                                    _lib.ptr(active), None, R, NA, float(clip), _lib.ptr(sums), _lib.ptr(first),
                                                    step, n_chunks, _lib.stream()))
        empty = logits.new_empty(0)
        ctx.save_for_backward(logits, actions, old_logp, adv, active if active is not None else empty, sums,
                              first if first is not None else empty)
        ctx.clip, ctx.has_active, ctx.map = float(clip), active is not None, (first is not None, step, n_chunks)
        policy_loss = (-sums[0] / sums[3]).to(torch.float32)
        entropy = (sums[1] / sums[3]).to(torch.float32)
        ratio = (sums[2] / R).to(torch.float32)
        ctx.mark_non_differentiable(ratio)
        return policy_loss, entropy, ratio

    @staticmethod
    def backward(ctx, g_pl, g_ent, _g_ratio):
        logits, actions, old_logp, adv, active, sums, first = ctx.saved_tensors
        R, NA = logits.shape
        inv = (1.0 / sums[3]).to(torch.float32)
        g = torch.stack([-g_pl.to(torch.float32) * inv, g_ent.to(torch.float32) * inv]).contiguous()
        dlogits = torch.empty_like(logits)
        _lib.check(_lib.lib().mm_ppo_actor_loss_bwd(_lib.ptr(logits), _lib.ptr(actions), _lib.ptr(old_logp), _lib.ptr(adv),
                                                    _lib.ptr(active if ctx.has_active else None), _lib.ptr(g), _lib.ptr(dlogits),
                                                    R, NA, ctx.clip, _lib.ptr(first if ctx.map[0] else None), ctx.map[1],
                                                    ctx.map[2], _lib.stream()))
        return dlogits, None, None, None, None, None


def ppo_actor_loss(logits, actions, old_action_log_probs, adv_targ, active_masks, clip_param):
    """-> (policy_loss, dist_entropy, mean importance weight).  ``active_masks=None``: plain means."""
    return _ActorLoss.apply(logits, actions, old_action_log_probs, adv_targ, active_masks, clip_param)


class _ValueLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, values, value_preds, returns, active, norm, clip, delta, use_huber, use_clipped):
        R = values.numel()
        v = values.reshape(R).to(torch.float32).contiguous()
        (vp, ret, active), first, step, n_chunks = _rows((value_preds, returns, active), R)
        sums = torch.zeros(2, dtype=torch.float64, device=values.device)
        cfg = (float(clip), float(delta), int(bool(use_huber)), int(bool(use_clipped)))
        _lib.check(_lib.lib().mm_ppo_value_loss_fwd(_lib.ptr(v), _lib.ptr(vp), _lib.ptr(ret), _lib.ptr(active), _lib.ptr(norm),
                                                    R, *cfg, _lib.ptr(sums), _lib.ptr(first), step, n_chunks, _lib.stream()))
        empty = values.new_empty(0)
        ctx.save_for_backward(v, vp, ret, active if active is not None else empty, norm if norm is not None else empty, sums,
                              first if first is not None else empty)
        ctx.cfg, ctx.has, ctx.map = cfg, (active is not None, norm is not None), (first is not None, step, n_chunks)
        ctx.shape = values.shape
        return (sums[0] / sums[1]).to(torch.float32)

    @staticmethod
    def backward(ctx, g_loss):
        v, vp, ret, active, norm, sums, first = ctx.saved_tensors
        g = (g_loss.to(torch.float32) / sums[1].to(torch.float32)).reshape(1).contiguous()
        dv = torch.empty_like(v)
        _lib.check(_lib.lib().mm_ppo_value_loss_bwd(_lib.ptr(v), _lib.ptr(vp), _lib.ptr(ret),
                                                    _lib.ptr(active if ctx.has[0] else None),
                                                    _lib.ptr(norm if ctx.has[1] else None), _lib.ptr(g), _lib.ptr(dv), v.numel(),
                                                    *ctx.cfg, _lib.ptr(first if ctx.map[0] else None), ctx.map[1], ctx.map[2],
                                                    _lib.stream()))
        return dv.view(ctx.shape), None, None, None, None, None, None, None, None


def ppo_value_loss(values, value_preds, returns, active_masks, norm, clip_param, huber_delta, use_huber, use_clipped):
    """``norm``: float32[2] device tensor {mean, 1/sqrt(var)} of the value normaliser, or None for raw returns."""
    return _ValueLoss.apply(values, value_preds, returns, active_masks, norm, clip_param, huber_delta, use_huber, use_clipped)
