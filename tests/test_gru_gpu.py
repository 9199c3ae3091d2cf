"""GPU parity: tcgen05 GRU recurrence (forward + BPTT) as an autograd op vs torch.nn.GRU in fp32 on the CPU.
TF32 operands for the h W_hh^T products: outputs 3e-3 abs, gradients 1% of their scale (relative Frobenius 3e-3)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a - b).norm() / b.norm().clamp(min=1e-12)).item()


@pytest.mark.parametrize("T,R,I,H,with_h0", [(1, 5, 64, 64, True), (12, 300, 64, 64, True), (7, 129, 32, 32, False), (40, 128, 64, 64, False),
                                               (5, 9600, 64, 64, True), (6, 9500, 32, 32, True), (9, 8192, 64, 64, False)])   # > 148 x 64 rows: 128-row tiles
def test_fused_gru_matches_nn_gru(cuda, T, R, I, H, with_h0):
    from minimarl_b200.nn.gru import GRU
    torch.backends.cudnn.allow_tf32 = True
    torch.manual_seed(T * 1000 + R)
    ref = torch.nn.GRU(I, H)
    with torch.no_grad():
        for p in ref.parameters():
            p.mul_(1.5)
    mine = GRU(I, H).to(cuda)
    mine.load_state_dict(ref.state_dict())
    x = torch.randn(T, R, I)
    h0 = torch.randn(1, R, H) * 0.5 if with_h0 else None
    w = torch.randn(T, R, H)
    xr = x.clone().requires_grad_(True)
    h0r = h0.clone().requires_grad_(True) if with_h0 else None
    out_r, hn_r = ref(xr, h0r)
    ((out_r * w).sum() + hn_r.sum()).backward()
    xm = x.to(cuda).requires_grad_(True)
    h0m = h0.to(cuda).requires_grad_(True) if with_h0 else None
    out_m, hn_m = mine(xm, h0m)
    ((out_m * w.to(cuda)).sum() + hn_m.sum()).backward()
    assert (out_m.cpu() - out_r).abs().max().item() < 3e-3
    assert (hn_m.cpu() - hn_r).abs().max().item() < 3e-3
    assert _rel(xm.grad.cpu(), xr.grad) < 3e-3
    if with_h0:
        assert _rel(h0m.grad.cpu(), h0r.grad) < 3e-3
    for (name, pm), (_, pr) in zip(mine.named_parameters(), ref.named_parameters()):
        assert _rel(pm.grad.cpu(), pr.grad) < 3e-3, name


def test_inference_path_saves_nothing(cuda):
    from minimarl_b200.nn.gru import GRU
    torch.backends.cudnn.allow_tf32 = True
    g = GRU(64, 64).to(cuda)
    with torch.no_grad():
        out, hn = g(torch.randn(5, 64, 64, device=cuda))
    assert out.shape == (5, 64, 64) and hn.shape == (1, 64, 64) and not out.requires_grad
