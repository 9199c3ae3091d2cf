"""Kernel-only and op-level timing of the fused row-parallel half of the agent network (csrc/agent_mlp.cu) at the QMIX bench
size: 101 steps x 8 192 (episode, agent) rows = 827 392 rows, D = 47, hidden 64."""
import sys

import torch

sys.path.insert(0, ".")
from minimarl_b200 import _lib  # noqa: E402
from minimarl_b200.nn import agent_mlp as M  # noqa: E402

cuda = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = True
rows, D = 101 * 8192, 47
g = torch.Generator(device=cuda).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=cuda, generator=g)
p = dict(ln0_w=1 + 0.2 * rn(D), ln0_b=0.2 * rn(D), w1=rn(64, D) / D ** 0.5, b1=0.2 * rn(64), ln1_w=1 + 0.2 * rn(64),
         ln1_b=0.2 * rn(64), w2=rn(64, 64) / 8, b2=0.2 * rn(64), ln2_w=1 + 0.2 * rn(64), ln2_b=0.2 * rn(64),
         wih=rn(192, 64) / 8, bih=0.2 * rn(192))
obs = torch.rand((rows, D), device=cuda, generator=g)
dgi = rn(rows, 192) * 1e-5
d = M.dims()
lib = _lib.lib()
order = "ln0_w ln0_b w1 b1 ln1_w ln1_b w2 b2 ln2_w ln2_b wih bih".split()
packed = torch.empty(d["n_params"], device=cuda)
_lib.check(lib.mm_agent_mlp_pack(M._ptr_array([p[k] for k in order]), _lib.ptr(packed), D, _lib.stream()))
gi = torch.empty((rows, 192), device=cuda)
n_tiles = (rows + 127) // 128
saved_rows = torch.empty((rows, d["saved_row_floats"]), device=cuda)
saved_x0 = torch.empty((n_tiles, d["x0_halves_per_tile"]), dtype=torch.float16, device=cuda)
scale = torch.empty(2, device=cuda)
n_ctas = lib.mm_agent_mlp_bwd_ctas(rows)
acc = torch.empty((n_ctas, 128, d["acc_cols"]), device=cuda)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=cuda)


def timeit(fn, n=6):
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


fwd = lambda: _lib.check(lib.mm_agent_mlp_fwd(_lib.ptr(packed), _lib.ptr(obs), _lib.ptr(gi), _lib.ptr(saved_rows), _lib.ptr(saved_x0),
                                              rows, D, _lib.stream()))
inf = lambda: _lib.check(lib.mm_agent_mlp_fwd(_lib.ptr(packed), _lib.ptr(obs), _lib.ptr(gi), None, None, rows, D, _lib.stream()))
sc = lambda: _lib.check(lib.mm_ppo_trunk2_scale(_lib.ptr(dgi), dgi.numel(), _lib.ptr(scale), _lib.stream()))
bwd = lambda: _lib.check(lib.mm_agent_mlp_bwd(_lib.ptr(packed), _lib.ptr(saved_rows), _lib.ptr(saved_x0), _lib.ptr(dgi), _lib.ptr(scale),
                                              _lib.ptr(acc), n_ctas, rows, _lib.stream()))
print(f"kernels: fwd (training) {timeit(fwd):.3f} ms, fwd (inference) {timeit(inf):.3f} ms, scale {timeit(sc):.3f} ms, bwd {timeit(bwd):.3f} ms")

pc = {k: v.clone().requires_grad_() for k, v in p.items()}
import torch.nn.functional as F
ln = lambda x, w, b: F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def op_fused():
    out = M.agent_mlp_gi(obs, (pc["ln0_w"], pc["ln0_b"]), (pc["w1"], pc["b1"]), (pc["ln1_w"], pc["ln1_b"]), (pc["w2"], pc["b2"]),
                         (pc["ln2_w"], pc["ln2_b"]), pc["wih"], pc["bih"])
    out.backward(dgi)


def op_eager():
    x = ln(obs, pc["ln0_w"], pc["ln0_b"])
    x = ln(torch.relu(x @ pc["w1"].t() + pc["b1"]), pc["ln1_w"], pc["ln1_b"])
    x = ln(torch.relu(x @ pc["w2"].t() + pc["b2"]), pc["ln2_w"], pc["ln2_b"])
    (x @ pc["wih"].t() + pc["bih"]).backward(dgi)


print(f"op forward + backward: fused {timeit(op_fused):.3f} ms, eager PyTorch (TF32 cuBLAS) {timeit(op_eager):.3f} ms")
