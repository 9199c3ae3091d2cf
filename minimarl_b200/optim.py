"""``FlatAdam``: gradient-norm clipping + Adam of one learner as two kernel launches on flat buffers.

The reference builds ``torch.optim.Adam(params, lr, eps)`` and runs ``clip_grad_norm_`` + ``step()`` after every backward pass
(offpolicy/algorithms/qmix/qmix.py:57,200-208; mappo/algorithms/r_mappo_policy.py:21-28 with ramppo_network.py:181-209;
vdn/main.py:97 with vdn/_train.py:96-100).  The networks on this path have 10^4..10^5 parameters spread over ~40 tensors, so
PyTorch's foreach kernels are launch-bound (~25 launches, 0.4 ms per step).  Here every parameter is a VIEW into one flat fp32
buffer, every ``p.grad`` a view into the matching ``GradBucket`` (the same buffer the multi-GPU all-reduce runs on),
and ``clip_and_step`` is ``mm_clip_adam_step`` (csrc/optim.cu): partial sums of squares, then clip coefficient + update.

It is a ``torch.optim.Optimizer``: ``param_groups`` (lr is read at every step, so ``lr_decay`` works), ``state`` (per parameter
``step`` / ``exp_avg`` / ``exp_avg_sq`` views, the layout ``torch.optim.Adam.state_dict()`` has), ``zero_grad``, ``step``.
"""
import torch

from ._lib import check, lib, ptr, stream
from .parallel import GradBucket

_ALIGN = 128          # floats: every parameter starts on a 512-byte boundary, like a tensor of its own would


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        params = [p for p in params if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) != 1:
            raise NotImplementedError("FlatAdam: one parameter group (the reference's optimisers have one)")
        ps = self.param_groups[0]["params"]
        if not ps or not all(p.is_cuda and p.dtype == torch.float32 for p in ps):
            raise RuntimeError("FlatAdam runs on CUDA float32 parameters (no CPU fallback): use torch.optim.Adam on the host")
        self.bucket = GradBucket(ps, align=_ALIGN)
        n, dev = self.bucket.flat.numel(), ps[0].device
        self.flat_params = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self._step = torch.zeros((), dtype=torch.float32, device=dev)
        self._scratch = torch.zeros(lib().mm_clip_adam_scratch_doubles(), dtype=torch.float64, device=dev)
        self._norm = torch.zeros((), dtype=torch.float32, device=dev)
        self._lr, self._lr_host = torch.full((), float(lr), dtype=torch.float32, device=dev), float(lr)
        with torch.no_grad():
            for p, off in zip(ps, self.bucket.offsets):
                view = self.flat_params[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                self.state[p] = {"step": self._step, "exp_avg": self.exp_avg[off:off + p.numel()].view_as(p),
                                 "exp_avg_sq": self.exp_avg_sq[off:off + p.numel()].view_as(p)}

    def _params_attached(self):
        base = self.flat_params.data_ptr()
        return all(p.data_ptr() == base + 4 * off for p, off in zip(self.bucket.params, self.bucket.offsets))

    def zero_grad(self, set_to_none=True):
        """One memset of the bucket; the ``p.grad`` views stay (``set_to_none`` would drop them)."""
        if not self.bucket.attached():
            self.bucket.reattach(zero=True)
        else:
            self.bucket.zero()

    def sync_lr(self):
        """The kernel reads the learning rate from a device scalar (so that a captured graph follows ``lr_decay``): refreshed
        here when ``param_groups[0]["lr"]`` changed.  Call it before replaying a graph that contains ``clip_and_step``."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_host:
            self._lr.fill_(lr)
            self._lr_host = lr

    @torch.no_grad()
    def clip_and_step(self, max_norm=0.0, clip_params=None):
        """``clip_grad_norm_(clip_params or all, max_norm)`` (``max_norm`` <= 0 / None: no clipping) + ``Adam.step()``.
        ``clip_params`` must be the leading parameters of this optimiser (qmix/_train.py clips the q-network only).  Returns
        the gradient norm before clipping as a 0-dim device tensor (overwritten by the next call)."""
        clip_n = self.bucket.n_grad
        if clip_params is not None:
            clip_params = list(clip_params)
            k = len(clip_params)
            if k > len(self.bucket.params) or any(a is not b for a, b in zip(clip_params, self.bucket.params)):
                raise ValueError("FlatAdam: clip_params must be the first parameters handed to the optimiser")
            clip_n = self.bucket.offsets[k] if k < len(self.bucket.params) else self.bucket.n_grad
        if not self.bucket.attached():
            self.bucket.reattach(zero=False)            # someone replaced a p.grad: pack once, views again from here on
        if not self._params_attached():
            raise RuntimeError("FlatAdam: a parameter's storage was replaced (p.data = ...); copy into it in place instead")
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        self.sync_lr()
        check(lib().mm_clip_adam_step(ptr(self.flat_params), ptr(self.bucket.flat), ptr(self.exp_avg), ptr(self.exp_avg_sq),
                                      ptr(self._step), ptr(self._scratch), ptr(self._norm), self.bucket.n_grad, clip_n, float(g["lr"]),
                                      ptr(self._lr), float(b1), float(b2), float(g["eps"]), float(g["weight_decay"]),
                                      float(max_norm) if max_norm else 0.0, stream()))
        self.bump_versions()
        return self._norm

    def bump_versions(self):
        """The kernel writes the parameters through raw pointers: tell autograd's version counters, which the packed-parameter
        caches of the rollout / target kernels key on (``p._version``).  A host-side operation -- callers that REPLAY a captured
        step (no Python runs then) call it after the replay."""
        torch.autograd.graph.increment_version(self.bucket.params)

    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self.clip_and_step(0.0)
        return loss

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        with torch.no_grad():
            for p, off in zip(self.bucket.params, self.bucket.offsets):
                st = self.state.get(p)
                if not st:
                    continue
                m, v = self.exp_avg[off:off + p.numel()].view_as(p), self.exp_avg_sq[off:off + p.numel()].view_as(p)
                m.copy_(st["exp_avg"])
                v.copy_(st["exp_avg_sq"])
                self._step.copy_(torch.as_tensor(st["step"], dtype=torch.float32))
                self.state[p] = {"step": self._step, "exp_avg": m, "exp_avg_sq": v}


def make_adam(params, lr, eps, weight_decay=0.0, capturable=False):
    """``FlatAdam`` for CUDA parameters, ``torch.optim.Adam`` for host parameters (the gloo / CPU tests of the host logic)."""
    params = list(params)
    if params and all(p.is_cuda for p in params):
        return FlatAdam(params, lr=lr, eps=eps, weight_decay=weight_decay)
    return torch.optim.Adam(params, lr=lr, eps=eps, weight_decay=weight_decay)


def clip_and_step(optimizer, params, max_norm, prefix=False):
    """The reference's ``clip_grad_norm_(params)`` + ``optimizer.step()`` pair; one fused call when the optimiser is a
    ``FlatAdam``.  ``prefix``: ``params`` are only the leading parameters of the optimiser (the rest is stepped unclipped)."""
    params = list(params)
    if isinstance(optimizer, FlatAdam):
        return optimizer.clip_and_step(max_norm, clip_params=params if prefix else None)
    if max_norm:
        norm = torch.nn.utils.clip_grad_norm_(params, max_norm)
    else:
        norm = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(p.grad) for p in params if p.grad is not None]))
    optimizer.step()
    return norm
