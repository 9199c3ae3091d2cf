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


@pytest.mark.parametrize("T,R,with_h0,accumulate", [(1, 5, True, False), (3, 77, True, True), (4, 200, False, False), (7, 128 * 160 + 9, True, False)])
def test_gru_wgrad_matches_float64(cuda, T, R, with_h0, accumulate):
    """mm_gru_wgrad (dW_hh / db_hh from dgi, dgh_n and the shifted h_seq on tcgen05, fp16 tiles with a power-of-two gradient
    scale) against the float64 contraction: 2e-3 relative (Frobenius) -- 11-bit operands, fp32 accumulation over T * R rows;
    ragged tiles, tiles straddling the first step, zero / given initial state, write and accumulate modes."""
    from minimarl_b200 import _lib
    H = 64
    g = torch.Generator().manual_seed(T * 1000 + R)
    dgi = torch.randn(T * R, 3 * H, generator=g) * 3e-4
    dghn = torch.randn(T * R, H, generator=g) * 3e-4
    h_seq = torch.tanh(torch.randn(T * R, H, generator=g))
    h0 = torch.tanh(torch.randn(R, H, generator=g)) if with_h0 else None
    hp = torch.cat([h0 if with_h0 else torch.zeros(R, H), h_seq[:-R]]).double()
    G = torch.cat([dgi[:, :2 * H], dghn], dim=1).double()
    dw_ref, db_ref = G.t() @ hp, G.sum(0)
    lib = _lib.lib()
    d = lambda x: x.to(cuda).contiguous()
    dgi_d, dghn_d, h_d, h0_d = d(dgi), d(dghn), d(h_seq), (d(h0) if with_h0 else None)
    scale = torch.empty(2, device=cuda)
    _lib.check(lib.mm_ppo_trunk2_scale(_lib.ptr(dgi_d), dgi_d.numel(), _lib.ptr(scale), _lib.stream()))
    n_ctas = lib.mm_gru_wgrad_ctas(T * R)
    acc = torch.empty((n_ctas, 128, 3 * H), device=cuda)
    base_w, base_b = torch.randn(3 * H, H, generator=g) * 1e-3, torch.randn(3 * H, generator=g) * 1e-3
    dw, db = (d(base_w), d(base_b)) if accumulate else (torch.empty(3 * H, H, device=cuda), torch.empty(3 * H, device=cuda))
    _lib.check(lib.mm_gru_wgrad(_lib.ptr(dgi_d), _lib.ptr(dghn_d), _lib.ptr(h_d), _lib.ptr(h0_d), _lib.ptr(scale), _lib.ptr(acc), n_ctas,
                                _lib.ptr(dw), _lib.ptr(db), int(accumulate), T, R, H, _lib.stream()))
    if accumulate:
        dw_ref, db_ref = dw_ref + base_w.double(), db_ref + base_b.double()
    assert _rel(dw.cpu().double(), dw_ref) < 2e-3 and _rel(db.cpu().double(), db_ref) < 2e-3
