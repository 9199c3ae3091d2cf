"""GPU parity: custom short-row LayerNorm forward / backward vs torch.nn.functional.layer_norm (fp32).
Tolerance 2e-6 abs on y and dx (same formula, different reduction order); dgamma / dbeta are sums over up to
10^5 rows accumulated per warp then atomically: 1e-4 relative."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,N", [(1, 47), (1000, 47), (4097, 64), (333, 32), (50000, 94), (77, 128), (5, 3), (100003, 64), (1, 64),
                                 (20001, 32), (9, 128)])
def test_layernorm_forward_backward(cuda, M, N):
    from minimarl_b200.nn import LayerNorm
    torch.manual_seed(M + N)
    ln = LayerNorm(N).to(cuda)
    with torch.no_grad():
        ln.weight.add_(0.2 * torch.randn(N, device=cuda))
        ln.bias.add_(0.2 * torch.randn(N, device=cuda))
    x = (torch.randn((M, N), device=cuda) * 2 + 0.5).requires_grad_(True)
    dy = torch.randn((M, N), device=cuda)
    y = ln(x)
    y.backward(dy)
    x2 = x.detach().clone().requires_grad_(True)
    w2, b2 = ln.weight.detach().clone().requires_grad_(True), ln.bias.detach().clone().requires_grad_(True)
    y2 = F.layer_norm(x2, (N,), w2, b2, 1e-5)
    y2.backward(dy)
    assert (y - y2).abs().max().item() < 5e-6
    assert (x.grad - x2.grad).abs().max().item() < 2e-5 * max(1.0, x2.grad.abs().max().item())
    assert torch.allclose(ln.weight.grad, w2.grad, rtol=1e-4, atol=1e-4 * max(1.0, w2.grad.abs().max().item()))
    assert torch.allclose(ln.bias.grad, b2.grad, rtol=1e-4, atol=1e-4 * max(1.0, b2.grad.abs().max().item()))


def test_layernorm_3d_input_and_state_dict_keys(cuda):
    from minimarl_b200.nn import LayerNorm
    ln = LayerNorm(64).to(cuda)
    assert set(ln.state_dict().keys()) == {"weight", "bias"}
    x = torch.randn((7, 33, 64), device=cuda)
    assert torch.allclose(ln(x), F.layer_norm(x, (64,), ln.weight, ln.bias), atol=5e-6)
    cpu = LayerNorm(64)                                      # parameters on the CPU: plain nn.LayerNorm path (fixture tooling)
    assert cpu(torch.randn(4, 64)).shape == (4, 64)


@pytest.mark.parametrize("rows,C", [(5000, 64), (100001, 5), (8192, 192), (4096, 1), (70000, 47)])
def test_colsum_and_gram_match_torch(cuda, rows, C):
    """Bias / weight gradient helpers of the learner: column-sum kernel and slab-split G^T X (fp32: 1e-5 relative)."""
    from minimarl_b200.ops.gram import colsum, gram
    g = torch.Generator(device=cuda).manual_seed(rows + C)
    x = torch.randn((rows, C), device=cuda, generator=g)
    ref = x.double().sum(0)
    assert float((colsum(x).double() - ref).abs().max()) < 1e-5 * float(x.abs().sum(0).max())
    y = torch.randn((rows, 32), device=cuda, generator=g)
    got, want = gram(x, y), x.double().t() @ y.double()
    assert float((got.double() - want).abs().max()) < 2e-5 * float(want.abs().max() + rows ** 0.5)
