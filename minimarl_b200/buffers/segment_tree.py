"""Device-resident segment trees with the reference's interface (offpolicy/utils/segment_tree.py).

``SumSegmentTree`` / ``MinSegmentTree`` keep the reference's layout -- a float64 array of
``2 * capacity`` nodes, root at 1, leaf ``i`` at ``capacity + i`` -- in HBM and run ``__setitem__``,
``reduce`` and ``find_prefixsum_idx`` as CUDA kernels (``mm_segtree_*``).  Index results are bit-exact
with the numpy trees: they only involve float64 add / subtract / compare.
"""
import numpy as np
import torch

from .. import _lib


def _dev_f64(x, device):
    return torch.as_tensor(np.asarray(x, dtype=np.float64) if not torch.is_tensor(x) else x,
                           dtype=torch.float64, device=device).reshape(-1).contiguous()


def _dev_i64(x, device):
    return torch.as_tensor(np.asarray(x, dtype=np.int64) if not torch.is_tensor(x) else x,
                           dtype=torch.int64, device=device).reshape(-1).contiguous()


class SegmentTree:
    def __init__(self, capacity, is_min, device="cuda"):
        assert capacity > 0 and capacity & (capacity - 1) == 0, "capacity must be positive and a power of 2."
        self._capacity = capacity
        self._is_min = bool(is_min)
        self.neutral_element = float("inf") if is_min else 0.0
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("minimarl_b200 segment trees live on a CUDA device (no CPU fallback)")
        self._value = torch.full((2 * capacity,), self.neutral_element, dtype=torch.float64, device=self.device)
        self._scalar = torch.empty(1, dtype=torch.float64, device=self.device)

    def reduce(self, start=0, end=None):
        """op over leaves [start, end) as a device scalar tensor (call ``.item()`` for a Python float)."""
        if end is None:
            end = self._capacity
        if end < 0:
            end += self._capacity
        out = torch.empty(1, dtype=torch.float64, device=self.device)
        _lib.check(_lib.lib().mm_segtree_reduce(_lib.ptr(self._value), self._capacity, int(start), int(end),
                                                int(self._is_min), _lib.ptr(out), _lib.stream()))
        return out[0]

    def __setitem__(self, idx, val):
        if isinstance(idx, (int, np.integer)):
            idx = [int(idx)]
        idx = _dev_i64(idx, self.device)
        n = idx.numel()
        if torch.is_tensor(val) or isinstance(val, (np.ndarray, list, tuple)):
            val = _dev_f64(val, self.device)
            if val.numel() == 1 and n > 1:
                val = val.expand(n).contiguous()
            vptr, scalar = _lib.ptr(val), 0.0
        else:
            vptr, scalar = None, float(val)
        s = _lib.ptr(self._value) if not self._is_min else None
        m = _lib.ptr(self._value) if self._is_min else None
        _lib.check(_lib.lib().mm_segtree_set(s, m, self._capacity, _lib.ptr(idx), vptr, scalar, 0, n, _lib.stream()))

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            assert 0 <= idx < self._capacity
            return self._value[self._capacity + int(idx)]
        return self._value[self._capacity + _dev_i64(idx, self.device)]


class SumSegmentTree(SegmentTree):
    def __init__(self, capacity, device="cuda"):
        super().__init__(capacity, False, device)

    def sum(self, start=0, end=None):
        return self.reduce(start, end)

    def find_prefixsum_idx(self, prefixsum):
        """Highest i with sum(arr[:i]) <= prefixsum, per entry (segment_tree.py:115-146).  int64 CUDA tensor."""
        mass = _dev_f64(prefixsum, self.device)
        out = torch.empty(mass.numel(), dtype=torch.int64, device=self.device)
        _lib.check(_lib.lib().mm_segtree_find_prefixsum(_lib.ptr(self._value), self._capacity, _lib.ptr(mass),
                                                        _lib.ptr(out), mass.numel(), _lib.stream()))
        return out


class MinSegmentTree(SegmentTree):
    def __init__(self, capacity, device="cuda"):
        super().__init__(capacity, True, device)

    def min(self, start=0, end=None):
        return self.reduce(start, end)
