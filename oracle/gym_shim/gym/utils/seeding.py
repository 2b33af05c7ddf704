"""`np_random(seed)` -> (numpy legacy RandomState, seed).

Real gym (<= 0.21) hashes the seed with SHA-512 before seeding the RandomState; that only
changes which stream a given integer selects.  Parity never depends on it: every test
records the draws the reference makes and replays those values.
"""
import numpy as np


def np_random(seed=None):
    if seed is not None and not (isinstance(seed, (int, np.integer)) and seed >= 0):
        raise ValueError("Seed must be a non-negative integer or omitted, not %r" % (seed,))
    if seed is None:
        seed = int(np.random.SeedSequence().generate_state(1)[0])
    rng = np.random.RandomState(int(seed) % (2 ** 32))
    return rng, seed
