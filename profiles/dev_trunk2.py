"""Development check of the second-generation trunk op on the GPU box: parity against a float64 statement of the trunk,
then timings of both generations at the MAPPO bench size (4.19 M rows per network pass)."""
import os
import sys
import time

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from minimarl_b200 import _lib  # noqa: E402
from minimarl_b200.nn.ppo_trunk import PARAM_ORDER, ChunkRows, ppo_trunk, ppo_trunk_v1  # noqa: E402
from test_ppo_trunk_gpu import _trunk_ref  # noqa: E402

cuda = torch.device("cuda:0")
torch.zeros(1, device=cuda)
DBG = torch.zeros(4, dtype=torch.int32).pin_memory()
_lib.check(_lib.lib().mm_ppo_trunk2_set_debug(DBG.data_ptr()))
import atexit
atexit.register(lambda: print("debug record {code, block, thread, set}:", DBG.tolist(), flush=True))
torch.backends.cuda.matmul.allow_tf32 = True


def make(L, R, D, NO, seed=0):
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    p = dict(ln0_w=1 + 0.2 * rn(D), ln0_b=0.2 * rn(D), w1=rn(32, D) / D ** 0.5, b1=0.2 * rn(32), ln1_w=1 + 0.2 * rn(32),
             ln1_b=0.2 * rn(32), w2=rn(32, 32) / 32 ** 0.5, b2=0.2 * rn(32), ln2_w=1 + 0.2 * rn(32), ln2_b=0.2 * rn(32),
             wih=rn(96, 32) / 32 ** 0.5, whh=rn(96, 32) / 32 ** 0.5, bih=0.2 * rn(96), bhh=0.2 * rn(96),
             ln3_w=1 + 0.2 * rn(32), ln3_b=0.2 * rn(32), wo=rn(NO, 32) / 32 ** 0.5, bo=0.2 * rn(NO))
    obs, h0 = torch.rand((L, R, D), generator=g, dtype=torch.float64), 0.5 * rn(R, 32)
    masks = (torch.rand((L, R), generator=g) > 0.25).double()
    w = rn(L, R, NO)
    return p, obs, h0, masks, w


def parity(L, R, D, NO, fwd_only=False, scale_w=1.0):
    p, obs, h0, masks, w = make(L, R, D, NO, seed=L * 1000 + R + D + NO)
    w = w * scale_w
    pr = {k: v.clone().requires_grad_() for k, v in p.items()}
    ref = _trunk_ref(obs, h0, masks, pr)
    (ref * w).sum().backward()
    pc = {k: v.float().to(cuda).requires_grad_(not fwd_only) for k, v in p.items()}
    out = ppo_trunk(obs.float().to(cuda), h0.float().to(cuda), masks.float().to(cuda), pc)
    torch.cuda.synchronize()
    err = float((out.detach().cpu().double() - ref.detach()).abs().max())
    print(f"L={L} R={R} D={D} NO={NO}: out max abs err {err:.3e}", flush=True)
    if fwd_only:
        return
    (out * w.float().to(cuda)).sum().backward()
    torch.cuda.synchronize()
    for k in PARAM_ORDER:
        a, b = pc[k].grad.cpu().double(), pr[k].grad
        print(f"   {k:6s} rel {float((a - b).norm() / (b.norm() + 1e-12)):.3e}", flush=True)


def timing():
    L, R, D, NO = 8, 524288, 47, 5
    p, _, _, _, _ = make(1, 1, D, NO)
    pc = {k: v.float().to(cuda).requires_grad_() for k, v in p.items()}
    rows = (128 + 1) * 16384 * 2
    g = torch.Generator(device=cuda).manual_seed(0)
    table = torch.rand((rows, D), device=cuda, generator=g)
    hs = torch.randn((rows, 32), device=cuda, generator=g) * 0.3
    mk = (torch.rand((rows, 1), device=cuda, generator=g) > 0.05).float()
    N = 16384 * 2
    perm = torch.randperm(R, device=cuda, generator=g)
    t0 = (perm % 16) * 8
    first = t0 * N + perm // 16
    cache = {}
    ob = ChunkRows(table, first, N, L, cache, "obs")
    h0 = ChunkRows(hs, first, N, 1)
    ms = ChunkRows(mk, first, N, L)
    dout = torch.randn((L, R, NO), device=cuda, generator=g) * 1e-6
    for name, fn in (("v2", ppo_trunk), ("v1", ppo_trunk_v1)):
        for slots in ((4, 2), (3, 2), (2, 1), (4, 1)) if name == "v2" else ((0, 0),):
            if name == "v2":
                _lib.check(_lib.lib().mm_ppo_trunk2_set_slots(*slots))
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            tf = tb = 0.0
            for it in range(4):
                for q in pc.values():
                    q.grad = None
                ev[0].record()
                out = fn(ob, h0, ms, pc)
                ev[1].record()
                out.backward(dout)
                ev[2].record()
                torch.cuda.synchronize()
                if it:
                    tf += ev[0].elapsed_time(ev[1]) / 3
                    tb += ev[1].elapsed_time(ev[2]) / 3
            print(f"{name} slots={slots}: fwd {tf:.3f} ms, bwd (incl. host-side ops) {tb:.3f} ms", flush=True)
            del out


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("fwd", "all"):
        parity(2, 1, 9, 3, fwd_only=True)
        parity(8, 300, 47, 5, fwd_only=True)
        parity(5, 128, 48, 6, fwd_only=True)
    if what in ("bwd", "all"):
        parity(2, 1, 9, 3)
        parity(8, 300, 47, 5)
        parity(3, 200, 30, 8)
        parity(8, 1000, 47, 1, scale_w=1e-6)
    if what in ("time", "all"):
        timing()
