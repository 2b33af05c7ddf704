"""CPU: the GAE oracle against the reference's own loop (ppo_multi_agent_new.py:205-218) -- golden vectors written by
executing the reference's statements, and, where the reference tree is present, the statements themselves."""
import os

import numpy as np
import pytest

import gae_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gae_golden.npz")


def gae_cases():
    g = np.load(GOLD)
    for i in range(int(g["n_cases"])):
        x = {k: g["c%d_%s" % (i, k)] for k in ("rewards", "values", "dones", "last_values", "last_dones")}
        gamma, lam = (float(v) for v in g["c%d_gamma_lam" % i])
        yield x, gamma, lam, g["c%d_advs" % i], g["c%d_returns" % i]


def test_oracle_reproduces_reference_golden_bit_for_bit():
    n = 0
    for x, gamma, lam, advs, rets in gae_cases():
        a, r = gae_oracle.gae(gamma=gamma, lam=lam, **x)
        assert a.dtype == np.float32 and np.array_equal(a.view(np.uint32), advs.view(np.uint32))
        assert np.array_equal(r.view(np.uint32), rets.view(np.uint32))
        n += 1
    assert n >= 5


@pytest.mark.reference
def test_oracle_equals_lifted_reference_statements():
    import ref_gae
    if not ref_gae.available():
        pytest.skip("reference learner source not present")
    lo, hi = ref_gae.cited_lines()
    assert 200 <= lo <= 210 and 215 <= hi <= 222, (lo, hi)   # ppo_multi_agent_new.py:205-218
    rng = np.random.RandomState(5)
    for T, N, gamma, lam in ((64, 50, 0.99, 0.95), (3, 4, 0.5, 0.5), (200, 3, 0.997, 1.0)):
        x = dict(rewards=rng.randn(T, N).astype(np.float32), values=rng.randn(T, N).astype(np.float32) * 10,
                 dones=rng.rand(T, N) < 0.2, last_values=rng.randn(N).astype(np.float32), last_dones=rng.rand(N) < 0.5)
        a, r = gae_oracle.gae(gamma=gamma, lam=lam, **x)
        ra, rr = ref_gae.reference_gae(gamma=gamma, lam=lam, **x)
        assert np.array_equal(a.view(np.uint32), ra.view(np.uint32)) and np.array_equal(r.view(np.uint32), rr.view(np.uint32))
