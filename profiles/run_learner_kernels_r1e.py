"""Driver of the r1e learner-kernel captures: GRU recurrence fwd + bwd (101 steps x 8 192 rows x 64, the QMIX update's
shape) and the tensor-core mixer forward (409 600 rows), each launched twice (the first pair is warm-up)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from minimarl_b200.nn import QMixer
from minimarl_b200.nn.gru import GRU

dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = True
torch.backends.cuda.matmul.allow_tf32 = True
R, T, H = 8192, 101, 64
g = GRU(H, H).to(dev)
x = torch.randn(T, R, H, device=dev, requires_grad=True)
h0 = torch.zeros(1, R, H, device=dev)
args = argparse.Namespace(mixer_hidden_dim=32, hypernet_hidden_dim=64, hypernet_layers=2, use_orthogonal=True,
                          fused_precision="tf32")
mixer = QMixer(args, 2, 94, dev)
q = torch.randn(100, 4096, 2, device=dev)
s = torch.rand(100, 4096, 94, device=dev)
for _ in range(2):
    y, _ = g(x, h0)
    y.sum().backward()
    with torch.no_grad():
        mixer(q, s)
torch.cuda.synchronize()
print("ok")
