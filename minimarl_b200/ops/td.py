"""Chosen-action gather / double-Q selection and the TD target + PER-weighted loss as CUDA kernels."""
import torch

from .. import _lib


def q_select(q_all, target_q_all, acts, T, B, A, double_q=True, want_greedy=False):
    """rows r = a*B + b.  q_all/target_q_all [T+1, A*B, NA], acts [T, A*B, NA] ->
    (q_taken [T,B,A], next_q [T,B,A] or None, act_idx [T, A*B] int64, greedy [T+1, A*B] int64 or None)."""
    NA = q_all.shape[-1]
    dev = q_all.device
    q_all = q_all.contiguous()
    acts = acts.contiguous()
    q_taken = torch.empty((T, B, A), dtype=torch.float32, device=dev)
    next_q = torch.empty((T, B, A), dtype=torch.float32, device=dev) if target_q_all is not None else None
    act_idx = torch.empty((T, A * B), dtype=torch.int64, device=dev)
    greedy = torch.empty((T + 1, A * B), dtype=torch.int64, device=dev) if want_greedy else None
    tq = target_q_all.contiguous() if target_q_all is not None else None
    _lib.check(_lib.lib().mm_q_select(_lib.ptr(q_all), _lib.ptr(tq), _lib.ptr(acts), _lib.ptr(q_taken), _lib.ptr(next_q),
                                      _lib.ptr(act_idx), _lib.ptr(greedy), T, B, A, NA, int(bool(double_q)), _lib.stream()))
    return q_taken, next_q, act_idx, greedy


class _QmixTdLoss(torch.autograd.Function):
    """loss = sum_b w_b sum_t err^2 / sum (1 - bad)  (qmix.py:176-186, mse_loss(e) = e^2)."""

    @staticmethod
    def forward(ctx, q_tot, next_q_tot, rewards, dones_env, weights, gamma, per_nu, per_eps):
        T, B = q_tot.shape[0], q_tot.shape[1]
        dev = q_tot.device
        z = dict(dtype=torch.float32, device=dev)
        err = torch.empty((T, B), **z)
        ep_loss, ep_valid, ep_qsum, prio = (torch.empty(B, **z) for _ in range(4))
        w = weights.to(**z).contiguous() if weights is not None else None
        _lib.check(_lib.lib().mm_qmix_td(
            _lib.ptr(q_tot.detach().reshape(T, B).contiguous()), _lib.ptr(next_q_tot.reshape(T, B).contiguous()),
            _lib.ptr(rewards.reshape(T, B).contiguous()), _lib.ptr(dones_env.reshape(T, B).contiguous()), _lib.ptr(w),
            float(gamma), float(per_nu), float(per_eps), T, B, _lib.ptr(err), _lib.ptr(ep_loss), _lib.ptr(ep_valid),
            _lib.ptr(ep_qsum), _lib.ptr(prio), _lib.stream()))
        denom = ep_valid.sum()
        loss = ep_loss.sum() / denom
        ctx.save_for_backward(err, w if w is not None else torch.ones(B, **z), denom)
        ctx.shape = q_tot.shape
        ctx.mark_non_differentiable(prio, ep_qsum, denom)
        return loss, prio, ep_qsum, denom

    @staticmethod
    def backward(ctx, g_loss, *_):
        err, w, denom = ctx.saved_tensors
        grad = (2.0 * g_loss / denom) * err * w.unsqueeze(0)
        return grad.reshape(ctx.shape), None, None, None, None, None, None, None


def qmix_td_loss(q_tot, next_q_tot, rewards, dones_env, weights, gamma, per_nu, per_eps):
    """Returns (loss, new_priorities [B] f32, masked-Q_tot mean, denominator)."""
    loss, prio, ep_qsum, denom = _QmixTdLoss.apply(q_tot, next_q_tot, rewards, dones_env, weights, gamma, per_nu, per_eps)
    return loss, prio, ep_qsum.sum() / ep_qsum.numel() / q_tot.shape[0], denom
