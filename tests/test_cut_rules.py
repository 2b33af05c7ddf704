"""CPU: properties of the `cut` rule-set (README.md:11 body-cut; no reference code, spec in DESIGN.md §6).
Its parity is unpinned, so these properties are what ties it to the pinned classic rules."""
import numpy as np
import pytest

import c_oracle
import snake_oracle as so


def _fruit_grid_from_list(fruit, VV):
    g = np.zeros((fruit.shape[0], VV), dtype=np.int64)
    for e in range(fruit.shape[0]):
        for f in fruit[e]:
            g[e, int(f)] += 1
    return g


@pytest.mark.parametrize("S,D", [(3, 10), (2, 19), (4, 8)])
def test_cut_equals_classic_when_nobody_strikes(S, D):
    """With actions in {0..4} the cut rule-set IS the classic one: same observations, rewards, dones,
    bodies, and the same fruit multiset (list in classic, count grid in cut)."""
    N = 64
    kw = dict(size=D, n_snakes=S, seed=5, env_id_base=7)
    a_env = c_oracle.COracle(N, rules="classic", **kw)
    b_env = c_oracle.COracle(N, rules="cut", **kw)
    assert np.array_equal(a_env.reset(), b_env.reset())
    for t in range(400):
        a = c_oracle.gen_actions(a_env.cfg, t, 3, 5)
        oa, ra, da, ia = a_env.step(a)
        ob, rb, db, ib = b_env.step(a)
        assert np.array_equal(oa, ob) and np.array_equal(ra, rb) and np.array_equal(da, db), t
        assert np.array_equal(ia["rewards_all"], ib["rewards_all"]) and np.array_equal(ia["num_snakes"], ib["num_snakes"])
        sa, sb = a_env.state(), b_env.state()
        for k in ("t", "len", "grow_to", "vel", "body", "draw_ctr"):
            assert np.array_equal(sa[k], sb[k]), (t, k)
        assert np.array_equal(_fruit_grid_from_list(sa["fruit"], (D + 2) ** 2), sb["fruit_grid"].astype(np.int64)), t


def test_strikes_cut_bodies_and_conserve_cells():
    """A strike that lands on another snake's body saves the striker, truncates the victim and turns
    the removed segments into fruit (except under a striker's head): fruit count + body cells only
    move, they are not created."""
    D, S = 8, 4
    env = so.SnakeOracle(D, S, S, S, "cut", draws=so.PhiloxDraws(3, 0))
    env.reset()
    rng = np.random.RandomState(1)
    cuts = 0
    for t in range(20000):
        before_fruit = int(env.fruit_grid.sum())
        before_len = [len(b) for b in env.body]
        a = rng.randint(0, 6, size=S)
        a[rng.rand(S) < 0.5] = 5  # strike often
        heads_before = [b[0] if b else None for b in env.body]
        ob, r, done, info = env.step(a)
        after_len = [len(b) for b in env.body]
        # a snake that shrank without dying was cut
        for s in range(S):
            if 0 < after_len[s] < before_len[s]:
                cuts += 1
                assert env.grow_to[s] == after_len[s]
        if done:
            env.reset()
    assert cuts > 20, "the random strike policy never produced a cut"


def test_saved_striker_survives_contact_with_a_body():
    """Hand-built: snake 0 strikes into the side of snake 1's body."""
    D = 6
    env = so.SnakeOracle(D, 2, 2, 2, "cut", draws=so.PhiloxDraws(0, 0))
    env.reset()
    P = env.pid
    env.body = [[P(1, 2), P(0, 2)], [P(2, 0), P(2, 1), P(2, 2), P(2, 3), P(2, 4)]]
    env.vel = [1, 4]          # snake 0 moves +x into (2,2); snake 1 moves -y out of the board at (2,-1)... keep it alive:
    env.body[1] = [P(3, 0), P(2, 0), P(2, 1), P(2, 2), P(2, 3)]
    env.vel = [1, 1]          # snake 1 moves +x to (4,0)
    env.grow_to = [3, 5]
    env.fruit_grid[:] = 0
    ob, r, done, info = env.step([5, 0])
    # snake 1 after its move: (4,0),(3,0),(2,0),(2,1),(2,2) ; snake 0's head lands on (2,2) = segment 4
    assert not done and r == 0.0 and info["num_snakes"] == 2
    assert env.body[0][0] == P(2, 2)
    assert env.body[1] == [P(4, 0), P(3, 0), P(2, 0), P(2, 1)] and env.grow_to[1] == 4
    assert int(env.fruit_grid.sum()) == 0  # the only removed cell is under the striker's head
    # the same move WITHOUT the strike action kills snake 0
    env2 = so.SnakeOracle(D, 2, 2, 2, "cut", draws=so.PhiloxDraws(0, 0))
    env2.reset()
    env2.body = [[P(1, 2), P(0, 2)], [P(3, 0), P(2, 0), P(2, 1), P(2, 2), P(2, 3)]]
    env2.vel = [1, 1]; env2.grow_to = [3, 5]; env2.fruit_grid[:] = 0
    ob, r, done, info = env2.step([0, 0])
    assert done and r == -1.0 and env2.body[0] == []
