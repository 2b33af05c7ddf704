"""Golden vectors of the reference's GAE loop (ppo_multi_agent_new.py:205-218), produced by EXECUTING the reference's
statements (oracle/ref_gae.py) in the build container.  -> tests/golden/gae_golden.npz.  TEST INFRASTRUCTURE."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_gae

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "gae_golden.npz")
CASES = [(64, 257, 0.99, 0.95), (5, 33, 0.9, 1.0), (1, 7, 0.99, 0.95), (128, 64, 0.999, 0.9), (16, 100, 1.0, 0.0)]


def inputs(T, N, seed):
    rng = np.random.RandomState(seed)
    return dict(rewards=rng.choice([-1.0, 0.0, 0.0, 0.0, 1.0, 2.0], size=(T, N)).astype(np.float32),
                values=(rng.randn(T, N) * 3).astype(np.float32), dones=rng.rand(T, N) < 0.1,
                last_values=rng.randn(N).astype(np.float32), last_dones=rng.rand(N) < 0.1)


if __name__ == "__main__":
    assert ref_gae.available(), "needs /root/reference"
    out = {"n_cases": len(CASES), "reference_lines": np.array(ref_gae.cited_lines())}
    for i, (T, N, gamma, lam) in enumerate(CASES):
        x = inputs(T, N, 1000 + i)
        advs, rets = ref_gae.reference_gae(gamma=gamma, lam=lam, **x)
        assert advs.dtype == np.float32 and rets.dtype == np.float32
        for k, v in x.items():
            out["c%d_%s" % (i, k)] = v
        out["c%d_gamma_lam" % i] = np.array([gamma, lam], dtype=np.float64)
        out["c%d_advs" % i], out["c%d_returns" % i] = advs, rets
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, "from reference lines", ref_gae.cited_lines())
