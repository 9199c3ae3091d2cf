"""Kernel-only timing of mm_ppo_trunk2_fwd / mm_ppo_trunk2_bwd at the MAPPO bench size (CUDA events around the launches, no
autograd, no host-side ops): 8-step chunks, 524 288 chunks = 4.19 M rows per pass, chunks addressed in place."""
import sys

import torch

sys.path.insert(0, ".")
from minimarl_b200 import _lib  # noqa: E402
from minimarl_b200.nn import ppo_trunk as T  # noqa: E402

cuda = torch.device("cuda:0")
L, R, D, NO, N = 8, 524288, 47, int(sys.argv[1]) if len(sys.argv) > 1 else 5, 32768
g = torch.Generator(device=cuda).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=cuda, generator=g)
p = dict(ln0_w=1 + 0.2 * rn(D), ln0_b=0.2 * rn(D), w1=rn(32, D) / D ** 0.5, b1=0.2 * rn(32), ln1_w=1 + 0.2 * rn(32),
         ln1_b=0.2 * rn(32), w2=rn(32, 32) / 32 ** 0.5, b2=0.2 * rn(32), ln2_w=1 + 0.2 * rn(32), ln2_b=0.2 * rn(32),
         wih=rn(96, 32) / 32 ** 0.5, whh=rn(96, 32) / 32 ** 0.5, bih=0.2 * rn(96), bhh=0.2 * rn(96),
         ln3_w=1 + 0.2 * rn(32), ln3_b=0.2 * rn(32), wo=rn(NO, 32) / 32 ** 0.5, bo=0.2 * rn(NO))
params = [p[k] for k in T.PARAM_ORDER]
rows = 129 * N
table = torch.rand((rows, D), device=cuda, generator=g)
hs = rn(rows, 32) * 0.3
mk = (torch.rand((rows,), device=cuda, generator=g) > 0.05).float()
perm = torch.randperm(R, device=cuda, generator=g)
first = ((perm % 16) * 8 * N + perm // 16).contiguous()
xhat = T.obs_table(table, mk)
packed = T._pack2(params, D, NO)
d = T.dims2()
out = torch.empty((L, R, NO), device=cuda)
n_tiles = (R + 127) // 128
saved = torch.empty(n_tiles * L * d["saved_bytes"], dtype=torch.uint8, device=cuda)
dout = rn(L, R, NO) * 1e-6
scale = torch.empty(2, device=cuda)
lib = _lib.lib()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=cuda)


def timeit(fn, n=8):
    ts = []
    for i in range(n + 2):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts), min(ts)


for fs, bs in ((4, 2), (3, 2), (2, 1)):
    _lib.check(lib.mm_ppo_trunk2_set_slots(fs, bs))
    n_ctas = lib.mm_ppo_trunk2_bwd_ctas(R)
    acc = torch.empty((n_ctas, 128, d["acc_cols"]), device=cuda)
    fwd = lambda: _lib.check(lib.mm_ppo_trunk2_fwd(_lib.ptr(packed), _lib.ptr(xhat), _lib.ptr(first), N, _lib.ptr(hs), None, 1,
                                                   _lib.ptr(out), None, _lib.ptr(saved), L, R, NO, _lib.stream()))
    inf = lambda: _lib.check(lib.mm_ppo_trunk2_fwd(_lib.ptr(packed), _lib.ptr(xhat), _lib.ptr(first), N, _lib.ptr(hs), None, 1,
                                                   _lib.ptr(out), None, None, L, R, NO, _lib.stream()))
    _lib.check(lib.mm_ppo_trunk2_scale(_lib.ptr(dout), dout.numel(), _lib.ptr(scale), _lib.stream()))
    bwd = lambda: _lib.check(lib.mm_ppo_trunk2_bwd(_lib.ptr(packed), _lib.ptr(xhat), _lib.ptr(first), N, None, 1, _lib.ptr(saved),
                                                   _lib.ptr(dout), _lib.ptr(scale), _lib.ptr(acc), n_ctas, L, R, NO, _lib.stream()))
    tf, ti, tb = timeit(fwd), timeit(inf), timeit(bwd)
    print(f"slots fwd {fs} bwd {bs}: fwd (training) {tf[0]:.3f} ms (min {tf[1]:.3f}), fwd (inference) {ti[0]:.3f}, "
          f"bwd {tb[0]:.3f} ms (min {tb[1]:.3f})", flush=True)
tt = timeit(lambda: T.obs_table(table, mk))
print(f"obs table {tt[0]:.3f} ms; pack + scale + unfold:",
      timeit(lambda: T._pack2(params, D, NO))[0], timeit(lambda: lib.mm_ppo_trunk2_scale(_lib.ptr(dout), dout.numel(), _lib.ptr(scale), _lib.stream()))[0])
