"""ncu target for the round-2 learner kernels at the QMIX bench size (batch 4 096 episodes x T = 100, 2 agents): the row-parallel
half of the live agent network (agent_mlp_fwd / agent_mlp_bwd, 827 392 rows), the live mixer (qmixer_tc_kernel forward,
qmixer_bwd_kernel + unfold backward, 409 600 rows) and FlatAdam (gradnorm_partial_kernel, clip_adam_kernel).  Prints CUDA-event
times of each (L2 flushed between runs) when run plain."""
import argparse as ap
import sys

import torch

sys.path.insert(0, ".")
from minimarl_b200 import _lib  # noqa: E402
from minimarl_b200.nn import QMixer  # noqa: E402
from minimarl_b200.nn import agent_mlp as M  # noqa: E402
from minimarl_b200.optim import FlatAdam  # noqa: E402

cuda = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = True
g = torch.Generator(device=cuda).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=cuda, generator=g)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=cuda)


def timeit(fn, n=5):
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
    return sum(ts) / len(ts)


# ---- agent_mlp --------------------------------------------------------------------------------------------------------------
rows, D = 101 * 8192, 47
p = dict(ln0_w=1 + 0.2 * rn(D), ln0_b=0.2 * rn(D), w1=rn(64, D) / D ** 0.5, b1=0.2 * rn(64), ln1_w=1 + 0.2 * rn(64),
         ln1_b=0.2 * rn(64), w2=rn(64, 64) / 8, b2=0.2 * rn(64), ln2_w=1 + 0.2 * rn(64), ln2_b=0.2 * rn(64),
         wih=rn(192, 64) / 8, bih=0.2 * rn(192))
pc = {k: v.clone().requires_grad_() for k, v in p.items()}
obs = torch.rand((rows, D), device=cuda, generator=g)
dgi = rn(rows, 192) * 1e-5


def agent_op():
    out = M.agent_mlp_gi(obs, (pc["ln0_w"], pc["ln0_b"]), (pc["w1"], pc["b1"]), (pc["ln1_w"], pc["ln1_b"]), (pc["w2"], pc["b2"]),
                         (pc["ln2_w"], pc["ln2_b"]), pc["wih"], pc["bih"])
    out.backward(dgi)


# ---- live mixer -------------------------------------------------------------------------------------------------------------
Mrows, SD = 100 * 4096, 94
margs = ap.Namespace(mixer_hidden_dim=32, hypernet_hidden_dim=64, hypernet_layers=2, use_orthogonal=True, fused_precision="tf32")
mixer = QMixer(margs, 2, SD, cuda)
q = (rn(Mrows, 1, 2) * 3).requires_grad_()
s = torch.rand((Mrows, 1, SD), device=cuda, generator=g)
dq_tot = rn(Mrows, 1, 1, 1) * 1e-5


def mixer_op():
    mixer(q, s).backward(dq_tot)


# ---- FlatAdam ---------------------------------------------------------------------------------------------------------------
params = [torch.nn.Parameter(rn(*sh)) for sh in [(64, 47), (64,), (64, 64), (64,), (192, 64), (192, 64), (192,), (192,), (5, 64), (5,),
                                                  (64, 94), (64, 64), (64, 94), (32, 64), (32, 94), (64, 94), (1, 64)]]
opt = FlatAdam(params, lr=5e-4, eps=1e-5)
for pp in params:
    pp.grad.copy_(torch.randn_like(pp))

for name, fn in (("agent_mlp fwd + bwd (op)", agent_op), ("live mixer fwd + bwd (op)", mixer_op), ("FlatAdam clip + step", lambda: opt.clip_and_step(10.0))):
    print(f"{name}: {timeit(fn):.3f} ms")
