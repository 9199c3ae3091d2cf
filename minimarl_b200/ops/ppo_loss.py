"""PPO actor / critic losses as one forward and one backward kernel each (``mm_ppo_actor_loss_*`` / ``mm_ppo_value_loss_*``).

They compute exactly what ``R_MAPPO.ppo_update`` / ``cal_value_loss`` compute with ~40 elementwise PyTorch calls
(mappo/algorithms/ramppo_network.py:70-171): clipped surrogate, masked means, Categorical entropy, clipped Huber / MSE
value error against (ValueNorm-)normalised returns; autograd's tie rules for min / max / clamp included.
"""
import torch

from .. import _lib


def _rows(xs, R, pack_key=None):
    """Per-row inputs: dense [R] vectors, or tables addressed like recurrent_generator's chunks (``ChunkRows`` objects, all
    sharing one row map).  -> (device pointers (int / None), tensors to keep alive, first, step_rows, n_chunks).

    Chunk-addressed tables are read by random 4-byte gathers -- one 32-byte DRAM sector per input and row, which is what bounds
    the loss kernels.  With a per-update cache (``ChunkRows.cache``) the inputs are interleaved ONCE per update into a [rows, 4]
    table (kept under ``pack_key``, with its sources under ``pack_key + ":src"`` so that a captured update can refresh it in
    place): a row's inputs then share one sector, and the kernels' row map ``first[c] + l * step`` runs on 4 x first / 4 x step."""
    lazy = [x for x in xs if x is not None and hasattr(x, "first")]
    if not lazy:
        ts = [None if x is None else x.reshape(R).to(torch.float32).contiguous() for x in xs]
        return [_lib.ptr(t) for t in ts], ts, None, 0, 0
    ref = lazy[0]
    assert len(lazy) == sum(x is not None for x in xs) and all(x.first is ref.first and x.step_rows == ref.step_rows for x in lazy)
    assert ref.L * ref.first.numel() == R
    cache = getattr(ref, "cache", None)
    if cache is None or pack_key is None or len(lazy) > 4:
        ts = [None if x is None else x.table.reshape(-1) for x in xs]
        return [_lib.ptr(t) for t in ts], ts, ref.first, ref.step_rows, ref.first.numel()
    pack = cache.get(pack_key)
    if pack is None:
        srcs = [x.table.reshape(-1) for x in lazy]
        n = min(t.numel() for t in srcs)                 # [T] and [T + 1]-step tensors share the (t, env, agent) row index
        pack = torch.zeros((n, 4), dtype=torch.float32, device=srcs[0].device)
        for j, t in enumerate(srcs):
            pack[:, j].copy_(t[:n])
        cache[pack_key], cache[pack_key + ":src"] = pack, srcs
    first4 = getattr(ref.first, "_mm_x4", None)
    if first4 is None:
        first4 = ref.first * 4
        ref.first._mm_x4 = first4
    base, ptrs, j = pack.data_ptr(), [], 0
    for x in xs:
        ptrs.append(None if x is None else base + 4 * j)
        j += x is not None
    return ptrs, [pack], first4, 4 * ref.step_rows, ref.first.numel()


def refresh_packs(cache):
    """Re-interleave the [rows, 4] loss-input tables of ``_rows`` in place from their (static) sources: a captured update."""
    for k, srcs in list(cache.items()):
        if k.endswith(":src"):
            pack = cache[k[:-4]]
            for j, t in enumerate(srcs):
                pack[:, j].copy_(t[:pack.shape[0]])


class _ActorLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, actions, old_logp, adv, active, clip):
        R, NA = logits.shape
        logits = logits.contiguous()
        ptrs, keep, first, step, n_chunks = _rows((actions, old_logp, adv, active), R, "actor_loss_pack")
        sums = torch.zeros(4, dtype=torch.float64, device=logits.device)
        _lib.check(_lib.lib().mm_ppo_actor_loss_fwd(_lib.ptr(logits), *ptrs, None, R, NA, float(clip), _lib.ptr(sums),
                                                    _lib.ptr(first), step, n_chunks, _lib.stream()))
        empty = logits.new_empty(0)
        ctx.save_for_backward(logits, sums, first if first is not None else empty, *[t for t in keep if t is not None])
        ctx.clip, ctx.ptrs, ctx.map = float(clip), ptrs, (first is not None, step, n_chunks)
        policy_loss = (-sums[0] / sums[3]).to(torch.float32)
        entropy = (sums[1] / sums[3]).to(torch.float32)
        ratio = (sums[2] / R).to(torch.float32)
        ctx.mark_non_differentiable(ratio)
        return policy_loss, entropy, ratio

    @staticmethod
    def backward(ctx, g_pl, g_ent, _g_ratio):
        logits, sums, first = ctx.saved_tensors[:3]          # (the per-row inputs stay alive in saved_tensors[3:]; ctx.ptrs address them)
        R, NA = logits.shape
        inv = (1.0 / sums[3]).to(torch.float32)
        g = torch.stack([-g_pl.to(torch.float32) * inv, g_ent.to(torch.float32) * inv]).contiguous()
        dlogits = torch.empty_like(logits)
        _lib.check(_lib.lib().mm_ppo_actor_loss_bwd(_lib.ptr(logits), *ctx.ptrs, _lib.ptr(g), _lib.ptr(dlogits),
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
        ptrs, keep, first, step, n_chunks = _rows((value_preds, returns, active), R, "value_loss_pack")
        sums = torch.zeros(2, dtype=torch.float64, device=values.device)
        cfg = (float(clip), float(delta), int(bool(use_huber)), int(bool(use_clipped)))
        _lib.check(_lib.lib().mm_ppo_value_loss_fwd(_lib.ptr(v), *ptrs, _lib.ptr(norm), R, *cfg, _lib.ptr(sums), _lib.ptr(first), step,
                                                    n_chunks, _lib.stream()))
        empty = values.new_empty(0)
        ctx.save_for_backward(v, norm if norm is not None else empty, sums, first if first is not None else empty,
                              *[t for t in keep if t is not None])
        ctx.cfg, ctx.ptrs, ctx.has_norm, ctx.map = cfg, ptrs, norm is not None, (first is not None, step, n_chunks)
        ctx.shape = values.shape
        return (sums[0] / sums[1]).to(torch.float32)

    @staticmethod
    def backward(ctx, g_loss):
        v, norm, sums, first = ctx.saved_tensors[:4]
        g = (g_loss.to(torch.float32) / sums[1].to(torch.float32)).reshape(1).contiguous()
        dv = torch.empty_like(v)
        _lib.check(_lib.lib().mm_ppo_value_loss_bwd(_lib.ptr(v), *ctx.ptrs, _lib.ptr(norm if ctx.has_norm else None), _lib.ptr(g),
                                                    _lib.ptr(dv), v.numel(), *ctx.cfg, _lib.ptr(first if ctx.map[0] else None),
                                                    ctx.map[1], ctx.map[2], _lib.stream()))
        return dv.view(ctx.shape), None, None, None, None, None, None, None, None


def ppo_value_loss(values, value_preds, returns, active_masks, norm, clip_param, huber_delta, use_huber, use_clipped):
    """``norm``: float32[2] device tensor {mean, 1/sqrt(var)} of the value normaliser, or None for raw returns."""
    return _ValueLoss.apply(values, value_preds, returns, active_masks, norm, clip_param, huber_delta, use_huber, use_clipped)
