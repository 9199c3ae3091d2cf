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



def test_tcgen05_descriptor_strides_probe(cuda):
    """Raw probe of the K-major no-swizzle descriptor: a single 1.0 at byte offset (m & 7) * 16 + (m >> 3) * SBO +
    (k >> 2) * LBO + (k & 3) * 4 of the A image lands in row m of D (B = ones), for non-default LBO / SBO as well."""
    from minimarl_b200 import _lib
    N = 16
    ones = torch.ones(16384, device=cuda)
    for lbo, sbo in ((128, 256), (256, 128), (1024, 128)):
        for m, k in ((0, 0), (5, 3), (9, 6), (77, 1), (127, 7)):
            off = (m & 7) * 16 + (m >> 3) * sbo + (k >> 2) * lbo + (k & 3) * 4
            if off >= 65536:
                continue
            img = torch.zeros(16384, device=cuda)
            img[off // 4] = 1.0
            D = torch.full((128, N), -1.0, device=cuda)
            _lib.check(_lib.lib().mm_tc_raw_selftest(_lib.ptr(img), _lib.ptr(ones), _lib.ptr(D), N, 1, 0, lbo, sbo, 0, 128, 256,
                                                     0, 0, _lib.stream()))
            torch.cuda.synchronize()
            rows = (D[:, 0] != 0).nonzero().flatten().tolist()
            # with SBO = 128 and LBO = 256 row m and row m + 16 (second k group of m) share a byte: both must light up
            assert m in rows and all((r - m) % 8 == 0 for r in rows), (lbo, sbo, m, k, rows)


def test_tcgen05_f16_one_tile_both_majors(cuda):
    """kind::f16: ONE [rows][features] fp16 tile in the no-swizzle core-matrix layout (8 x 16-byte core matrices, feature chunks
    outermost) is a K-major operand when the features are contracted (LBO = chunk stride, SBO = 128) and an MN-major operand when
    the rows are contracted (LBO = 128 between groups of 8 rows, SBO = chunk stride).  Exact on small integers."""
    import numpy as np
    from profiles.probe_f16_layouts import run, tile_image
    rng = np.random.default_rng(1)
    G = rng.integers(-4, 5, (128, 128)).astype(np.float32)
    X = rng.integers(-4, 5, (128, 48)).astype(np.float32)
    W = rng.integers(-4, 5, (96, 32)).astype(np.float32)
    ig, chg = tile_image(G)
    ix, chx = tile_image(X)
    iw, chw = tile_image(W)
    D = run(ig, ix, 128, 48, 8, (256, 128, chg), (256, 128, chx), 1, 1)              # G^T X over the 128 rows
    assert np.array_equal(D, G.T @ X)
    D = run(ig, iw, 128, 32, 6, (2 * chg, chg, 128), (256, 128, chw), 0, 1)          # G[:, :96] W with W stored [96][32]
    assert np.array_equal(D, G[:, :96] @ W)
    D = run(ix, iw, 128, 96, 2, (2 * chx, chx, 128), (2 * chw, chw, 128), 0, 0)      # X[:, :32] W^T
    assert np.array_equal(D, X[:, :32] @ W.T)
