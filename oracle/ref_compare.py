"""Lock-step comparison helpers: reference env (via ref_loader) vs SnakeOracle.  TEST INFRASTRUCTURE."""
import numpy as np

from snake_oracle import SnakeOracle, _U32

_VEL_CODE = {(0, 0): 0, (1, 0): 1, (0, 1): 2, (-1, 0): 3, (0, -1): 4}


def canonical_from_reference(env, S, rules, cap):
    """The reference's list-of-tuples state (snake_multiple_test.py:230) as canonical arrays."""
    snakes, fruits, vels, grow, t = env.state
    V = env.dim + 2
    pid = lambda c: (c[0] + 1) * V + (c[1] + 1)
    body = np.zeros((S, cap), dtype=np.uint16)
    for s in range(S):
        body[s, :len(snakes[s])] = [pid(c) for c in snakes[s]]
    out = {
        "t": np.int32(t), "spare": np.uint32(getattr(env, "spare_fruits", 0) & _U32),
        "len": np.array([len(snakes[s]) for s in range(S)], dtype=np.uint16),
        "grow_to": np.array(grow[:S], dtype=np.uint16),
        "vel": np.array([_VEL_CODE[tuple(vels[s])] for s in range(S)], dtype=np.uint8),
        "body": body,
    }
    if rules == "classic":
        out["fruit"] = np.array([pid(c) for c in fruits], dtype=np.uint16)
    else:
        g = np.zeros(V * V, dtype=np.int64)
        for c in fruits:
            g[pid(c)] += 1
        out["fruit_grid"] = np.minimum(g, 255).astype(np.uint8)
    return out


def assert_same_canonical(a, b, ctx=""):
    assert a.keys() == b.keys(), ctx
    for k in a:
        assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), "%s: state field %r differs\n%r\n%r" % (ctx, k, a[k], b[k])


def lockstep(rules, S, D, steps, seed, action_seed, n_actions=5, load_state=None):
    """Runs reference and oracle side by side from identically seeded RandomStates; asserts
    equality of obs (K=3 as the reference emits), reward, done, num_snakes and full state at
    every step.  Returns (#steps, #episodes, #draws)."""
    import ref_loader
    ref = ref_loader.make_env(rules, S, D, np.random.RandomState(seed))
    orc = SnakeOracle(D, S, S, 3, rules, draws=np.random.RandomState(seed))
    cap = D * D + 1
    ob_r, ob_o = ref.reset(), orc.reset()
    assert np.array_equal(ob_r, ob_o)
    arng = np.random.RandomState(action_seed)
    episodes = 0
    for i in range(steps):
        a = arng.randint(0, n_actions, size=S)
        ob_r, r_r, d_r, info_r = ref.step(a)
        ob_o, r_o, d_o, info_o = orc.step(a)
        ctx = "%s S=%d D=%d step %d" % (rules, S, D, i)
        assert np.array_equal(ob_r, ob_o), ctx
        assert float(r_r) == float(r_o) and bool(d_r) == bool(d_o), ctx
        assert info_r["num_snakes"] == info_o["num_snakes"], ctx
        assert_same_canonical(canonical_from_reference(ref, S, rules, cap), orc.canonical(cap), ctx)
        if d_r:
            episodes += 1
            assert np.array_equal(ref.reset(), orc.reset()), ctx
    return steps, episodes


def lockstep_scripted(rules, S, D, steps, seed):
    """Like lockstep() but with the fruit-seeking script of make_golden.py; returns max body length seen."""
    import ref_loader
    from make_golden import scripted_action
    ref = ref_loader.make_env(rules, S, D, np.random.RandomState(seed))
    orc = SnakeOracle(D, S, S, 3, rules, draws=np.random.RandomState(seed))
    cap = D * D + 1
    assert np.array_equal(ref.reset(), orc.reset())
    arng = np.random.RandomState(seed + 1)
    max_len = 0
    for i in range(steps):
        a = np.array([scripted_action(ref, s, arng) for s in range(S)])
        ob_r, r_r, d_r, info_r = ref.step(a)
        ob_o, r_o, d_o, info_o = orc.step(a)
        ctx = "%s S=%d D=%d step %d" % (rules, S, D, i)
        assert np.array_equal(ob_r, ob_o), ctx
        assert float(r_r) == float(r_o) and bool(d_r) == bool(d_o) and info_r["num_snakes"] == info_o["num_snakes"], ctx
        assert_same_canonical(canonical_from_reference(ref, S, rules, cap), orc.canonical(cap), ctx)
        max_len = max(max_len, max(len(b) for b in ref.state[0]))
        if d_r:
            assert np.array_equal(ref.reset(), orc.reset()), ctx
    return max_len
