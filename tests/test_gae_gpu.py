"""GPU parity: GAE / returns reverse scan vs the numpy oracle -- bit-exact (the kernel reproduces
the reference's float32/float64 operation order), so the stated tolerance is 0 ulp."""
import numpy as np
import pytest
import torch

from oracle import gae_ref

pytestmark = pytest.mark.gpu


def _inputs(T, rows, seed):
    rng = np.random.default_rng(seed)
    rewards = rng.standard_normal((T, rows)).astype(np.float32)
    values = rng.standard_normal((T + 1, rows)).astype(np.float32)
    masks = (rng.random((T + 1, rows)) < 0.97).astype(np.float32)
    return rewards, values, masks


@pytest.mark.parametrize("T,rows", [(1, 1), (5, 3), (100, 64), (128, 4096), (37, 10001)])
@pytest.mark.parametrize("denorm", [None, (np.sqrt(np.float32(1.3)), np.float32(0.1))])
def test_gae_bit_exact(cuda, T, rows, denorm):
    from minimarl_b200.ops import gae_returns
    rewards, values, masks = _inputs(T, rows, T * 7 + rows)
    ref = gae_ref.compute_returns(rewards, values, masks, 0.99, 0.95, True, denorm)
    ref_adv = gae_ref.advantages(ref, values, denorm)
    returns = torch.zeros((T + 1, rows), device=cuda)
    adv = torch.empty((T, rows), device=cuda)
    gae_returns(torch.from_numpy(rewards).to(cuda), torch.from_numpy(values).to(cuda),
                torch.from_numpy(masks).to(cuda), returns, 0.99, 0.95, use_gae=True, denorm=denorm,
                advantages=adv)
    assert np.array_equal(returns[:-1].cpu().numpy(), ref[:-1])
    assert np.array_equal(adv.cpu().numpy(), ref_adv)


def test_discounted_returns_bit_exact(cuda):
    from minimarl_b200.ops import gae_returns
    T, rows = 100, 777
    rewards, values, masks = _inputs(T, rows, 5)
    last = values[-1].copy()
    ref = gae_ref.compute_returns(rewards, values, masks, 0.99, 0.95, False, None, returns_last=last)
    returns = torch.zeros((T + 1, rows), device=cuda)
    returns[-1] = torch.from_numpy(last).to(cuda)
    gae_returns(torch.from_numpy(rewards).to(cuda), None, torch.from_numpy(masks).to(cuda), returns,
                0.99, 0.95, use_gae=False)
    assert np.array_equal(returns.cpu().numpy(), ref)


def test_gae_full_size_linearity(cuda):
    """BASELINE size (T=128, 16 384 envs x 2 agents): with masks == 1 and no denorm the map
    (rewards, values) -> returns is linear; check on the GPU result itself (tolerance 2e-4 abs)."""
    from minimarl_b200.ops import gae_returns
    T, rows = 128, 32768
    g = torch.Generator(device=cuda).manual_seed(0)
    r1, r2 = (torch.randn((T, rows), device=cuda, generator=g) for _ in range(2))
    v1, v2 = (torch.randn((T + 1, rows), device=cuda, generator=g) for _ in range(2))
    ones = torch.ones((T + 1, rows), device=cuda)

    def run(r, v):
        out = torch.zeros((T + 1, rows), device=cuda)
        gae_returns(r, v, ones, out, 0.99, 0.95)
        return out[:-1]
    lhs = run(r1 + r2, v1 + v2)
    rhs = run(r1, v1) + run(r2, v2)
    assert torch.allclose(lhs, rhs, atol=2e-4, rtol=1e-5)


def test_gae_full_size_matches_oracle(cuda):
    """BASELINE size (T = 128, 16 384 envs x 2 agents = 32 768 rows), ValueNorm denormalisation and advantages in the same
    pass: bit-equal to the numpy oracle (pinned on tests/golden/gae.npz)."""
    from minimarl_b200.ops import gae_returns
    T, rows = 128, 32768
    rewards, values, masks = _inputs(T, rows, 11)
    denorm = (np.sqrt(np.float32(1.3)), np.float32(0.1))
    ref = gae_ref.compute_returns(rewards, values, masks, 0.99, 0.95, True, denorm)
    ref_adv = gae_ref.advantages(ref, values, denorm)
    returns = torch.zeros((T + 1, rows), device=cuda)
    adv = torch.zeros((T, rows), device=cuda)
    gae_returns(torch.from_numpy(rewards).to(cuda), torch.from_numpy(values).to(cuda), torch.from_numpy(masks).to(cuda),
                returns, 0.99, 0.95, use_gae=True, denorm=denorm, advantages=adv)
    assert np.array_equal(returns[:-1].cpu().numpy(), ref[:-1])
    assert np.array_equal(adv.cpu().numpy(), ref_adv)
