"""In-kernel clock64() phase profile of ppo_trunk (mm_ppo_trunk_set_prof): CTA 0 / thread 0, cycles per tile-step."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from minimarl_b200 import _lib
from minimarl_b200.nn.ppo_trunk import PARAM_ORDER, ppo_trunk

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = True
L, R, D, NO = 8, 148 * 8 * 128, 47, 5
g = torch.Generator(device=dev).manual_seed(0)
H = 32
shapes = dict(ln0_w=(D,), ln0_b=(D,), w1=(H, D), b1=(H,), ln1_w=(H,), ln1_b=(H,), w2=(H, H), b2=(H,), ln2_w=(H,), ln2_b=(H,),
              wih=(3 * H, H), whh=(3 * H, H), bih=(3 * H,), bhh=(3 * H,), ln3_w=(H,), ln3_b=(H,), wo=(NO, H), bo=(NO,))
params = {k: (torch.randn(s, device=dev, generator=g) * 0.2 + (1.0 if k.endswith("_w") and k.startswith("ln") else 0.0)).requires_grad_(True)
          for k, s in shapes.items()}
obs = torch.rand((L, R, D), device=dev, generator=g)
h0 = torch.randn((R, H), device=dev, generator=g) * 0.1
masks = (torch.rand((L, R), device=dev, generator=g) < 0.98).float()
names_f = ["F1 ln0+tiles", "publish", "mma fc1 wait", "epi1+publish+issue", "mma fc2 wait", "epi2+publish+issue", "mma gates wait",
           "gates+ln3+head"]
names_b = {5: "F1: wait tiles, reload xh/hp, x2 tile, publish, issue", 6: "mma gates wait", 7: "gates+ln3", 13: "b1: dout, dy, S sums (TMEM)",
           14: "b1: ln3 bwd", 15: "b1: gru bwd + dG tile", 8: "publish + issue + G store", 9: "mma dx2/dhp wait",
           10: "ln2 bwd + publish + issue + dA2 store", 11: "mma dx1 wait", 12: "ln1 bwd + dA1 store"}
for _ in range(2):
    out = ppo_trunk(obs, h0, masks, params)
    out.sum().backward()
torch.cuda.synchronize()
prof = torch.zeros(16, dtype=torch.int64, device=dev)
lib = _lib.lib()
steps = (R // 128 // 148) * L
for mode in ("fwd", "bwd"):
    prof.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if mode == "fwd":
        lib.mm_ppo_trunk_set_prof(prof.data_ptr())
        e0.record(); out = ppo_trunk(obs, h0, masks, params); e1.record()
        lib.mm_ppo_trunk_set_prof(None)
        torch.cuda.synchronize()
    else:
        out = ppo_trunk(obs, h0, masks, params)
        go = torch.ones_like(out)
        torch.cuda.synchronize()
        lib.mm_ppo_trunk_set_prof(prof.data_ptr())
        e0.record(); out.backward(go); e1.record()
        lib.mm_ppo_trunk_set_prof(None)
        torch.cuda.synchronize()
    p = prof.cpu().tolist()
    print(f"{mode} total ms (incl. host-side ops) {e0.elapsed_time(e1):.3f}  cycles/step {sum(p) / steps:.0f}")
    items = list(enumerate(names_f)) if mode == "fwd" else sorted(names_b.items(), key=lambda kv: [5, 6, 7, 13, 14, 15, 8, 9, 10, 11, 12].index(kv[0]))
    for i, n in items:
        print(f"   {n:55s} {p[i] / steps:8.0f} cyc/step")
