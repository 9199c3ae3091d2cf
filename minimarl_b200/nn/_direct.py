"""Opt-in: fused ops add their weight gradients straight into the parameters' existing ``.grad`` tensors."""
import torch


class direct_param_grads:
    """``with direct_param_grads(): loss.backward()`` -- the fused ops (PPO trunk, agent trunk, live mixer, LayerNorm + head,
    recurrent weight gradients) add their parameter gradients to the parameters' existing ``.grad`` tensors (the views of a
    ``GradBucket``) from the kernel that produces them and hand autograd ``None``: no accumulation launch per parameter tensor
    (18 per network pass and 540 per MAPPO update; ~30 per QMIX update).  Only valid for a plain ``backward()`` into ``.grad``
    (not ``torch.autograd.grad``), hence opt-in by the trainers."""
    active = False

    def __enter__(self):
        self._prev, direct_param_grads.active = direct_param_grads.active, True

    def __exit__(self, *exc):
        direct_param_grads.active = self._prev


def grad_targets(param_objs, needs):
    """The ``.grad`` tensors of ``param_objs`` when direct writes are on, every one of them needs a gradient and can take it in
    place (contiguous float32 CUDA storage); else None."""
    if not direct_param_grads.active or not all(needs):
        return None
    gs = [p.grad for p in param_objs]
    if all(g is not None and g.is_cuda and g.dtype == torch.float32 and g.is_contiguous() for g in gs):
        return gs
    return None
