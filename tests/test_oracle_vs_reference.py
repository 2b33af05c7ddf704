"""CPU: oracle vs the reference.  The sha256 goldens of SURVEY.md section 8c are reproduced by the
oracle anywhere (numpy's legacy RandomState is deterministic); the lock-step tests additionally
run the unmodified reference and only exist in the build container."""
import hashlib

import numpy as np
import pytest

import snake_oracle as so

SURVEY_HASHES = [
    # rules, S, D, sum reward, #done, sum num_snakes, sha256
    ("classic", 2, 19, -198.0, 221, 7048, "ba3abeea69b0d37d1e6485b4b1d75ca95eb89b5b487bc60f65d613e8bbc18491"),
    ("classic", 3, 10, -497.0, 639, 8693, "04be6b907c90f459fb3e422d8decaa93ab11f2ca127d091376535252c4baccf4"),
    ("classic", 1, 10, -512.0, 553, 4447, "095576c9ed478625ee43d8c8ff830d220c07ac955938a8f50ae0b9b608200dec"),
    ("adversarial", 3, 10, -444.0, 647, 9016, "9962e07ff6a1aded02233d51cbd249780256f4539e4450644768cb06e2bfc609"),
    ("adversarial", 2, 10, -505.0, 620, 6885, "4a1e0332d29019e31c9c0e54c713bb098b1504bf122f6ceb09b1d57cd858f369"),
]


def _survey_run(env, S):
    h = hashlib.sha256()
    h.update(env.reset().tobytes())
    arng = np.random.RandomState(123)
    sr = nd = ns = 0
    for _ in range(5000):
        a = arng.randint(0, 5, size=S)
        ob, r, d, info = env.step(a)
        h.update(ob.tobytes())
        h.update(np.float32(r).tobytes())
        h.update(bytes([int(d), info["num_snakes"]]))
        sr, nd, ns = sr + r, nd + int(d), ns + info["num_snakes"]
        if d:
            h.update(env.reset().tobytes())
    return float(sr), nd, ns, h.hexdigest()


@pytest.mark.parametrize("rules,S,D,sr,nd,ns,sha", SURVEY_HASHES)
def test_oracle_reproduces_survey_hashes(rules, S, D, sr, nd, ns, sha):
    env = so.SnakeOracle(D, S, S, 3, rules, draws=np.random.RandomState(0))
    assert _survey_run(env, S) == (sr, nd, ns, sha)


@pytest.mark.reference
@pytest.mark.parametrize("rules,S,D,sr,nd,ns,sha", SURVEY_HASHES)
def test_reference_reproduces_survey_hashes(rules, S, D, sr, nd, ns, sha):
    import ref_loader
    env = ref_loader.make_env(rules, S, D, np.random.RandomState(0))
    assert _survey_run(env, S) == (sr, nd, ns, sha)


@pytest.mark.reference
@pytest.mark.parametrize("rules,S,D", [("classic", 2, 19), ("classic", 3, 10), ("classic", 1, 10), ("adversarial", 3, 10),
                                        ("adversarial", 2, 7), ("classic", 3, 3), ("classic", 2, 2), ("adversarial", 3, 4)])
def test_lockstep_random_policy(rules, S, D):
    import ref_compare
    steps, episodes = ref_compare.lockstep(rules, S, D, 1500, seed=31 + D, action_seed=77)
    assert episodes > 10


@pytest.mark.reference
@pytest.mark.parametrize("rules,S,D", [("classic", 2, 19), ("classic", 3, 10), ("adversarial", 3, 10), ("classic", 1, 6)])
def test_lockstep_scripted_policy(rules, S, D):
    """Fruit-seeking actions: long bodies, many respawns on crowded boards, self collisions."""
    import ref_compare
    max_len = ref_compare.lockstep_scripted(rules, S, D, 1500, seed=5)
    assert max_len >= 8


@pytest.mark.reference
@pytest.mark.parametrize("S,D", [(2, 19), (3, 10)])
def test_world_view_matches_reference_get_ob_world(S, D):
    """render(mode='rgb_array') source: get_ob_world (snake_multiple_test.py:60-91)."""
    import ref_loader
    ref = ref_loader.make_env("classic", S, D, np.random.RandomState(11))
    orc = so.SnakeOracle(D, S, S, 3, "classic", draws=np.random.RandomState(11))
    ref.reset(); orc.reset()
    rng = np.random.RandomState(12)
    for _ in range(300):
        assert np.array_equal(ref.get_ob_world(), orc.world_view())
        a = rng.randint(0, 5, size=S)
        _, _, d, _ = ref.step(a)
        orc.step(a)
        if d:
            ref.reset(); orc.reset()
