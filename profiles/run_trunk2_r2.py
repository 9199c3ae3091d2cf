"""ncu target: the second-generation PPO trunk at the MAPPO bench size (8-step chunks, 524 288 chunks = 4.19 M rows per pass,
chunks addressed in place in a [T + 1, N, A, 47] table): three forward + backward passes."""
import sys

import torch

sys.path.insert(0, ".")
from minimarl_b200.nn.ppo_trunk import PARAM_ORDER, ChunkRows, ppo_trunk  # noqa: E402

cuda = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = True
L, R, D, NO, N = 8, 524288, 47, 5, 32768
g = torch.Generator(device=cuda).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=cuda, generator=g)
p = dict(ln0_w=1 + 0.2 * rn(D), ln0_b=0.2 * rn(D), w1=rn(32, D) / D ** 0.5, b1=0.2 * rn(32), ln1_w=1 + 0.2 * rn(32),
         ln1_b=0.2 * rn(32), w2=rn(32, 32) / 32 ** 0.5, b2=0.2 * rn(32), ln2_w=1 + 0.2 * rn(32), ln2_b=0.2 * rn(32),
         wih=rn(96, 32) / 32 ** 0.5, whh=rn(96, 32) / 32 ** 0.5, bih=0.2 * rn(96), bhh=0.2 * rn(96),
         ln3_w=1 + 0.2 * rn(32), ln3_b=0.2 * rn(32), wo=rn(NO, 32) / 32 ** 0.5, bo=0.2 * rn(NO))
pc = {k: v.requires_grad_() for k, v in p.items()}
rows = 129 * N
table = torch.rand((rows, D), device=cuda, generator=g)
hs = rn(rows, 32) * 0.3
mk = (torch.rand((rows, 1), device=cuda, generator=g) > 0.05).float()
perm = torch.randperm(R, device=cuda, generator=g)
first = (perm % 16) * 8 * N + perm // 16
cache = {}
ob, h0, ms = ChunkRows(table, first, N, L, cache, "obs"), ChunkRows(hs, first, N, 1), ChunkRows(mk, first, N, L)
dout = rn(L, R, NO) * 1e-6
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(3):
    ev[0].record()
    out = ppo_trunk(ob, h0, ms, pc)
    ev[1].record()
    out.backward(dout)
    ev[2].record()
    torch.cuda.synchronize()
    print(f"pass {it}: fwd {ev[0].elapsed_time(ev[1]):.3f} ms, bwd {ev[1].elapsed_time(ev[2]):.3f} ms (host-side ops included)")
