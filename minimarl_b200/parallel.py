"""The two multi-GPU pieces of the hot path: env sharding and the gradient exchange.

Env instances are independent, so they are partitioned by contiguous blocks, one process per GPU, with
no data-path collective.  The only exchange step is the learner's gradient all-reduce: one flattened
fp32 bucket (83 KB MAPPO / 260 KB QMIX), summed with NCCL over NVLink and divided by the world size --
the semantics of the reference's (dead) ``average_gradients`` (offpolicy/utils/util.py:148-153).
The collective is latency-bound at this size; it is issued once per optimiser step, before clipping.
"""
import torch
import torch.distributed as dist


def shard_range(n_total, rank, world_size):
    """Contiguous block [lo, hi) of ``n_total`` env instances owned by ``rank`` (remainder to low ranks)."""
    base, rem = divmod(n_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class GradBucket:
    """One contiguous fp32 gradient buffer per optimiser with every ``p.grad`` a VIEW into it ("packing grads into
    one contiguous buffer in the backward epilogue").  Autograd accumulates into existing ``.grad`` tensors in place and the
    fused backward kernels write through the ``.grad`` pointers, so after a backward pass the bucket already holds the packed
    gradients: the exchange is ONE collective on ``flat`` -- no ``torch.cat``, no copy-back.  ``n_extra`` trailing slots carry
    scalars that must be summed over ranks in the same collective (loss / count statistics).

    ``zero()`` replaces ``optimizer.zero_grad()`` (which, with ``set_to_none=True``, would drop the views).  Parameters that
    never receive a gradient (the reference's unused ``fc_h`` duplicates, quirk Q6) still own a zero slice, on every rank alike.
    """

    def __init__(self, params, n_extra=0, align=1):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradBucket needs at least one parameter")
        dev = self.params[0].device
        self.offsets, n = [], 0                   # ``align`` (floats): start of every slice; the gaps stay zero
        for p in self.params:
            self.offsets.append(n)
            n += -(-p.numel() // align) * align
        self.n_grad, self.n_extra = n, int(n_extra)
        self.flat = torch.zeros(n + self.n_extra, dtype=torch.float32, device=dev)
        self.reattach(zero=False)

    def reattach(self, zero=True):
        """(Re)install the views; existing gradients are copied in unless ``zero``."""
        if zero:
            self.flat.zero_()
        for p, off in zip(self.params, self.offsets):
            view = self.flat[off:off + p.numel()].view_as(p)
            if p.grad is not None and not zero and p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
            p.grad = view

    @property
    def grads(self):
        return self.flat[:self.n_grad]

    @property
    def extra(self):
        return self.flat[self.n_grad:]

    def attached(self):
        """False once something replaced a ``p.grad`` (e.g. ``zero_grad(set_to_none=True)``)."""
        base = self.flat.data_ptr()
        return all(p.grad is not None and p.grad.data_ptr() == base + 4 * off for p, off in zip(self.params, self.offsets))

    def zero(self):
        self.flat.zero_()

    def allreduce_mean(self, group=None):
        """Sum over ranks, gradients divided by the world size, extras left as sums.  Returns the bytes exchanged."""
        if not (dist.is_available() and dist.is_initialized()):
            return 0
        world = dist.get_world_size(group)
        if world == 1:
            return 0
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        self.grads.div_(world)
        return self.flat.numel() * 4


def allreduce_mean_grads(params, group=None, bucket=None):
    """In-place average of ``p.grad`` over all ranks, one bucket, one collective.  With a ``GradBucket`` whose views are still
    attached this is a single all-reduce on its flat buffer; otherwise the gradients are packed and unpacked here.
    Parameters without a gradient (e.g. the reference's unused ``fc_h`` duplicates, Q6) stay out of the packed bucket on every
    rank."""
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    world = dist.get_world_size(group)
    if world == 1:
        return 0
    if bucket is not None and bucket.attached():
        return bucket.allreduce_mean(group)
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= world
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    return flat.numel() * flat.element_size()


def per_global_normaliser(p_min_local, p_sum_local, n_local, beta, group=None):
    """Sharded PER: every rank samples batch / G episodes from ITS tree, but the importance weights must follow
    the global formula of ``PrioritizedRecReplayBuffer.sample`` (offpolicy/utils/rec_buffer.py:278-304):

        p_min = min_i p_i / sum_i p_i,  max_weight = (p_min * N) ** -beta,  w_i = (p_i / sum * N) ** -beta / max_weight

    with min, sum and N taken over ALL shards.  One all-reduce of (-min, sum, len) with MAX / SUM semantics packed into a
    single SUM collective is not possible, so two tiny collectives are issued (MIN on one scalar, SUM on two); returns
    (p_min_global, p_sum_global, n_global, max_weight) as Python floats / tensors of the input kind."""
    t_min = torch.as_tensor(p_min_local, dtype=torch.float64).reshape(1).clone()
    t_sum = torch.stack([torch.as_tensor(p_sum_local, dtype=torch.float64).reshape(()),
                         torch.as_tensor(float(n_local), dtype=torch.float64).reshape(())])
    if torch.is_tensor(p_sum_local) and p_sum_local.is_cuda:
        t_min, t_sum = t_min.to(p_sum_local.device), t_sum.to(p_sum_local.device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t_min, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(t_sum, op=dist.ReduceOp.SUM, group=group)
    p_min, p_sum, n = t_min[0], t_sum[0], t_sum[1]
    max_weight = (p_min / p_sum * n) ** (-beta)
    return p_min, p_sum, n, max_weight


def per_global_weights(p_sampled, p_min_local, p_sum_local, n_local, beta, group=None):
    """Importance weights of locally sampled priorities ``p_sampled`` under the global normaliser."""
    _, p_sum, n, max_weight = per_global_normaliser(p_min_local, p_sum_local, n_local, beta, group)
    p = torch.as_tensor(p_sampled, dtype=torch.float64)
    return (p.to(p_sum.device) / p_sum * n) ** (-beta) / max_weight


def allreduce_stats(values, group=None):
    """Sum a small vector of scalars (advantage-normalisation sums, ValueNorm batch moments, PER min / sum /
    len) over ranks so that sharded runs reproduce the single-process formulas."""
    t = values if torch.is_tensor(values) else torch.tensor(values, dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t
