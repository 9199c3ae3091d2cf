"""CPU: the numpy GAE oracle against fixtures produced by the reference's own SharedReplayBuffer +
ValueNorm (tests/golden/make_golden.py) -- bit for bit."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import gae_ref


@pytest.mark.parametrize("tag", ["vn", "plain", "nogae", "long"])
def test_gae_oracle_matches_reference_fixture(tag):
    g = np.load(os.path.join(GOLDEN, "gae.npz"))
    use_gae, use_vn = (bool(x) for x in g[f"{tag}_flags"])
    values = g[f"{tag}_value_preds_in"].copy()
    denorm = tuple(g[f"{tag}_denorm"]) if use_vn else None
    if use_gae:
        values[-1] = g[f"{tag}_next_value"]
        got = gae_ref.compute_returns(g[f"{tag}_rewards"], values, g[f"{tag}_masks"], 0.99, 0.95, True, denorm)
        assert np.array_equal(got[:-1], g[f"{tag}_returns"][:-1])
        assert np.array_equal(gae_ref.advantages(got, values, denorm), g[f"{tag}_advantages"])
    else:
        got = gae_ref.compute_returns(g[f"{tag}_rewards"], values, g[f"{tag}_masks"], 0.99, 0.95, False, None,
                                      returns_last=g[f"{tag}_next_value"])
        assert np.array_equal(got, g[f"{tag}_returns"])
