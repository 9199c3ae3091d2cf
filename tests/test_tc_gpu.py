"""GPU: tcgen05 (TF32, TMEM accumulator) building blocks against an fp32 matmul.
TF32 keeps 10 mantissa bits: tolerance 2e-3 * sqrt(K) * |a||b| scale, checked as relative Frobenius error < 1e-3."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K", [(64, 48), (64, 64), (192, 64), (16, 8), (256, 128)])
def test_tcgen05_tf32_gemm_selftest(cuda, N, K):
    from minimarl_b200 import _lib
    g = torch.Generator(device=cuda).manual_seed(N * 1000 + K)
    A = torch.randn((128, K), device=cuda, generator=g)
    B = torch.randn((N, K), device=cuda, generator=g)
    D = torch.zeros((128, N), device=cuda)
    _lib.check(_lib.lib().mm_tc_gemm_selftest(_lib.ptr(A), _lib.ptr(B), _lib.ptr(D), N, K, _lib.stream()))
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    rel = ((D.double() - ref).norm() / ref.norm()).item()
    assert rel < 1e-3, rel
    # exactly representable inputs (small integers) must give the exact result: proves the layout, not just the scale
    Ai = torch.randint(-4, 5, (128, K), device=cuda, generator=g).float()
    Bi = torch.randint(-4, 5, (N, K), device=cuda, generator=g).float()
    _lib.check(_lib.lib().mm_tc_gemm_selftest(_lib.ptr(Ai), _lib.ptr(Bi), _lib.ptr(D), N, K, _lib.stream()))
    assert torch.equal(D, Ai @ Bi.t())


@pytest.mark.parametrize("ta,tb,N,K,FA", [(1, 1, 96, 128, 32), (1, 1, 32, 128, 128), (0, 1, 32, 96, 0), (0, 1, 48, 32, 0),
                                          (1, 0, 64, 64, 48), (1, 1, 64, 128, 64)])
def test_tcgen05_transposed_operand_descriptors(cuda, ta, tb, N, K, FA):
    """MN-major descriptors on the K-major tiles: contraction over tile rows (exact on small integers)."""
    from minimarl_b200 import _lib
    g = torch.Generator(device=cuda).manual_seed(7 + N + K)
    A = torch.randint(-4, 5, (K, FA) if ta else (128, K), device=cuda, generator=g).float()
    B = torch.randint(-4, 5, (K, N) if tb else (N, K), device=cuda, generator=g).float()
    D = torch.zeros((128, N), device=cuda)
    _lib.check(_lib.lib().mm_tc_gemm_t_selftest(_lib.ptr(A), _lib.ptr(B), _lib.ptr(D), N, K, FA, ta, tb, _lib.stream()))
    torch.cuda.synchronize()
    ref = (A.t() if ta else A) @ (B if tb else B.t())
    rows = FA if ta else 128
    assert torch.equal(D[:rows], ref)
