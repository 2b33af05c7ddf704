"""GPU: the CUDA path (through the C ABI) against the CPU oracle and the reference recordings.
Everything is bit-exact: obs bytes, rewards, dones, num_snakes, episode stats and full state."""
import ctypes as C
import os
import zlib

import numpy as np
import pytest

import c_oracle
import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import snakes_b200
    return snakes_b200


def _state_equal(dev_state, cpu_state, ctx):
    for k in cpu_state:
        assert np.array_equal(dev_state[k], cpu_state[k]), "%s: state field %r differs" % (ctx, k)


def _compare_step(env, co, actions, ctx, check_state=True):
    import torch
    obs, rew, done, infos = env.step(torch.as_tensor(actions, dtype=torch.int8, device=env.device))
    cobs, crew, cdone, cinfo = co.step(actions)
    assert np.array_equal(rew.cpu().numpy(), crew), ctx + " reward"
    assert np.array_equal(done.cpu().numpy(), cdone), ctx + " done"
    assert np.array_equal(env.num_alive.cpu().numpy(), cinfo["num_snakes"]), ctx + " num_snakes"
    assert np.array_equal(env.rewards_all.cpu().numpy(), cinfo["rewards_all"]), ctx + " rewards_all"
    assert np.array_equal(env.episode_return.cpu().numpy(), cinfo["episode_r"]), ctx + " episode r"
    assert np.array_equal(env.episode_len.cpu().numpy(), cinfo["episode_l"]), ctx + " episode l"
    o = obs.cpu().numpy()
    if not np.array_equal(o, cobs):
        bad = np.flatnonzero((o != cobs).reshape(len(o), -1).any(1))
        raise AssertionError("%s obs differs in envs %s" % (ctx, bad[:8]))
    if check_state:
        _state_equal(env.dump_state(), co.state(), ctx)
    return o


@pytest.mark.parametrize("name", helpers.golden_names())
def test_golden_replay(sb, name):
    """Replays the reference's recorded draws and actions: every step must reproduce the
    reference's rewards / dones / num_snakes / Monitor stats and the digests of its observations
    and states (11k+ episodes for classic_2x19)."""
    g = helpers.load_golden(name)
    N, T = g["N"], g["T"]
    env = sb.SnakeVecEnv(N, size=g["D"], n_snakes=g["S"], n_fruits=g["F"], n_views=g["K"], rules=g["rules"])
    env.set_draw_tape(g["tape_vals"], g["tape_bounds"], g["tape_offsets"])
    obs = env.reset()
    assert zlib.crc32(obs.cpu().numpy().tobytes()) == int(g["reset_obs_crc"])
    import torch
    actions = torch.as_tensor(g["actions"], device=env.device)
    for t in range(T):
        obs, rew, done, infos = env.step(actions[t])
        ctx = "%s step %d" % (name, t)
        assert np.array_equal(rew.cpu().numpy(), g["reward"][t]), ctx
        assert np.array_equal(done.cpu().numpy(), g["done"][t].astype(bool)), ctx
        assert np.array_equal(env.num_alive.cpu().numpy(), g["num_snakes"][t]), ctx
        assert np.array_equal(env.episode_return.cpu().numpy(), g["ep_r"][t]), ctx
        assert np.array_equal(env.episode_len.cpu().numpy(), g["ep_l"][t]), ctx
        assert zlib.crc32(obs.cpu().numpy().tobytes()) == int(g["obs_crc"][t]), ctx
        if t % 25 == 0 or t == T - 1:
            assert helpers.batch_state_crc(env.dump_state()) == int(g["state_crc"][t]), ctx
    env.check_errors()
    st = env.dump_state()
    assert np.array_equal(obs.cpu().numpy(), g["final_obs"])
    for k in ("t", "spare", "len", "grow_to", "vel", "body"):
        assert np.array_equal(st[k], g["final_" + k]), k
    fk = "fruit_grid" if "fruit_grid" in st else "fruit"
    assert np.array_equal(st[fk], g["final_" + fk])
    assert np.array_equal(st["draw_ctr"].astype(np.uint64), np.diff(g["tape_offsets"]))
    env.close()


CONFIGS = [
    # rules, S, D, F, K, N, steps
    ("classic", 2, 19, 2, 2, 1000, 300),     # BASELINE configs[3] geometry, N not a multiple of the CTA tile
    ("classic", 2, 10, 2, 2, 4096, 200),     # configs[1]
    ("classic", 3, 10, 3, 3, 777, 300),
    ("classic", 1, 10, 1, 1, 1, 300),        # configs[0]: a single env
    ("classic", 1, 10, 1, 3, 33, 200),       # more views than snakes
    ("classic", 4, 7, 6, 2, 300, 300),       # F > S, K < S
    ("classic", 2, 10, 4, 2, 500, 300),      # NewMultipleSnakes defaults: 2 snakes, 4 fruits
    ("classic", 2, 10, 0, 2, 200, 200),      # no fruit at all
    ("classic", 4, 19, 4, 4, 300, 200),      # 4 snakes, 4 views on 19x19: 42 KB image per warp
    ("cut", 4, 19, 4, 4, 200, 200),
    ("classic", 2, 2, 2, 2, 300, 200),       # smallest board
    ("adversarial", 3, 32, 3, 3, 64, 150),   # largest board of the lane kernel class
    ("classic", 3, 3, 3, 3, 256, 300),       # crowded board: no-free-cell + alias paths
    ("adversarial", 3, 10, 3, 3, 500, 400),
    ("adversarial", 2, 5, 2, 2, 129, 400),
    ("cut", 3, 10, 3, 3, 512, 400),          # configs[2]
    ("cut", 5, 8, 5, 5, 200, 400),
    ("classic", 2, 30, 2, 2, 64, 200),       # large board, still the lane kernel class
    ("classic", 5, 12, 5, 5, 100, 200),      # S = 5: outside the lane kernel class -> tile kernel
    ("classic", 2, 40, 2, 2, 40, 100),       # D = 40: outside the lane kernel class -> tile kernel
]


@pytest.mark.parametrize("rules,S,D,F,K,N,steps", CONFIGS)
def test_philox_vs_oracle(sb, rules, S, D, F, K, N, steps):
    """Production RNG: device and oracle generate the same draws; compare everything every step."""
    kw = dict(size=D, n_snakes=S, n_fruits=F, n_views=K, rules=rules, seed=1234, env_id_base=5000)
    env = sb.SnakeVecEnv(N, **kw)
    co = c_oracle.COracle(N, **kw)
    assert np.array_equal(env.reset().cpu().numpy(), co.reset())
    nact = env.action_space.n
    for t in range(steps):
        a = c_oracle.gen_actions(co.cfg, t, 99, nact)
        if t < 3:
            assert np.array_equal(env.gen_actions(t, 99).cpu().numpy(), a)
        _compare_step(env, co, a, "%s S%d D%d step %d" % (rules, S, D, t), check_state=(t % 10 == 0 or t == steps - 1))
    env.check_errors()
    assert co.errors() == 0
    stats = env.stats(reduce=False)
    assert np.allclose([stats[k] for k in sb._lib.STAT_NAMES], co.stats())
    env.close()


def test_scripted_long_snakes_vs_oracle(sb):
    """Fruit-seeking actions computed from the oracle state: long bodies, ring wrap-around,
    respawn on crowded boards."""
    N, D, S = 64, 8, 2
    kw = dict(size=D, n_snakes=S, rules="classic", seed=7)
    env = sb.SnakeVecEnv(N, **kw)
    co = c_oracle.COracle(N, **kw)
    assert np.array_equal(env.reset().cpu().numpy(), co.reset())
    V = D + 2
    rng = np.random.RandomState(3)
    max_len = 0
    for t in range(1500):
        st = co.state()
        a = np.zeros((N, S), dtype=np.int8)
        for e in range(N):
            occ = set()
            for s in range(S):
                occ.update(int(p) for p in st["body"][e, s, :st["len"][e, s]])
            for s in range(S):
                L = int(st["len"][e, s])
                if L == 0 or rng.rand() < 0.05:
                    a[e, s] = rng.randint(0, 5)
                    continue
                head, vel = int(st["body"][e, s, 0]), int(st["vel"][e, s])
                best, bd = 0, None
                for act, dl in ((1, V), (2, 1), (3, -V), (4, -1)):
                    if vel and act == ((vel + 1) & 3) + 1:
                        continue
                    p = head + dl
                    x, y = p // V - 1, p % V - 1
                    if not (0 <= x < D and 0 <= y < D) or p in occ:
                        continue
                    d = min(abs(x - (int(f) // V - 1)) + abs(y - (int(f) % V - 1)) for f in st["fruit"][e])
                    if bd is None or d < bd:
                        best, bd = act, d
                a[e, s] = best
        max_len = max(max_len, int(st["len"].max()))
        _compare_step(env, co, a, "scripted step %d" % t, check_state=(t % 20 == 0))
    assert max_len >= 20
    env.close()


def test_known_answers(sb):
    """Hand-built states stepped once by the reference, via load_state."""
    for ka in helpers.known_answers():
        S, D = ka["S"], ka["D"]
        env = sb.SnakeVecEnv(1, size=D, n_snakes=S, n_fruits=len(ka["in"]["fruits"]), n_views=3, rules=ka["rules"], auto_reset=False)
        i = ka["in"]
        env.load_state_blob(helpers.state_blob_from_lists(env.lay, env.cfg, i["snakes"], i["fruits"], i["vels"], i["grow_to"], i["t"], i["spare"]))
        vals = np.array([d[1] for d in ka["draws"]], dtype=np.uint32)
        bounds = np.array([d[0] for d in ka["draws"]], dtype=np.uint32)
        env.set_draw_tape(vals, bounds, np.array([0, len(vals)], dtype=np.uint64))
        obs, rew, done, infos = env.step(np.array([ka["action"]], dtype=np.int8))
        note = ka["note"]
        env.check_errors()
        assert float(rew[0]) == ka["reward"] and bool(done[0]) == ka["done"], note
        assert infos[0]["num_snakes"] == ka["num_snakes"], note
        assert zlib.crc32(obs[0].cpu().numpy().tobytes()) == ka["obs_crc"], note
        st = env.dump_state()
        got = helpers.lists_from_state(st, env.cfg)
        want = dict(ka["out"])
        if ka["rules"] != "classic":
            want["fruits"] = sorted(want["fruits"])
        assert got == want, note
        assert int(st["draw_ctr"][0]) == len(vals), note
        env.close()


def test_rows_kernel_small_fields(sb):
    """The row-chunked CTA-per-env kernel forced onto boards whose rows are 16-byte multiples."""
    for rules, S, D, K, N, steps in (("classic", 4, 14, 4, 60, 150), ("cut", 4, 30, 4, 40, 150), ("adversarial", 3, 10, 4, 50, 150)):
        kw = dict(size=D, n_snakes=S, n_views=K, rules=rules, seed=19)
        env = sb.SnakeVecEnv(N, debug="force_kernel=rows", **kw)
        assert env.launch_info()["kernel"] == "k_step_rows"
        co = c_oracle.COracle(N, **kw)
        assert np.array_equal(env.reset().cpu().numpy(), co.reset())
        for t in range(steps):
            a = c_oracle.gen_actions(co.cfg, t, 5, env.action_space.n)
            _compare_step(env, co, a, "rows %s step %d" % (rules, t), check_state=(t % 10 == 0))
        env.close()


@pytest.mark.parametrize("kernel", ["tile", "dense"])
def test_general_kernels_match_oracle(sb, kernel):
    """The warp-per-env tile kernel and the CTA-per-env dense kernel (used for boards outside the
    lane kernel's class) forced onto small configs."""
    for rules, S, D, N, steps in (("classic", 2, 19, 100, 150), ("adversarial", 3, 10, 100, 200), ("cut", 3, 10, 100, 200)):
        kw = dict(size=D, n_snakes=S, rules=rules, seed=11)
        env = sb.SnakeVecEnv(N, debug="force_kernel=" + kernel, **kw)
        assert env.launch_info()["kernel"] == "k_step_" + kernel
        co = c_oracle.COracle(N, **kw)
        assert np.array_equal(env.reset().cpu().numpy(), co.reset())
        for t in range(steps):
            a = c_oracle.gen_actions(co.cfg, t, 5, env.action_space.n)
            _compare_step(env, co, a, "%s %s step %d" % (kernel, rules, t), check_state=(t % 10 == 0))
        env.close()


def test_tile_kernel_golden_replay(sb, monkeypatch):
    """The tile kernel also replays the reference recording of the headline geometry."""
    monkeypatch.setenv("SNK_DEBUG", "force_kernel=tile")   # the one environment variable snk_create reads
    test_golden_replay(sb, "classic_2x19")
    monkeypatch.delenv("SNK_DEBUG")


@pytest.mark.parametrize("rules,S,D", [("adversarial", 32, 40), ("cut", 16, 24)])
def test_large_field_many_views(sb, rules, S, D):
    """The large-field kernel with 32 views (96-byte pixels, the maximum snake count) and with 16 views on a small board."""
    N = 24
    kw = dict(size=D, n_snakes=S, rules=rules, seed=5)
    env = sb.SnakeVecEnv(N, debug="force_kernel=rows", **kw)
    assert env.launch_info()["kernel"] == "k_step_rows"
    co = c_oracle.COracle(N, **kw)
    assert np.array_equal(env.reset().cpu().numpy(), co.reset())
    for t in range(40):
        a = c_oracle.gen_actions(co.cfg, t, 5, env.action_space.n)
        _compare_step(env, co, a, "%dx%d %s step %d" % (D, D, rules, t), check_state=(t % 10 == 0))
    env.close()


@pytest.mark.parametrize("rules", ["classic", "cut"])
def test_large_field_16_snakes_64x64(sb, rules):
    """BASELINE configs[4] geometry (16 snakes, 64x64, 209 KB of observation per env)."""
    N = 48
    kw = dict(size=64, n_snakes=16, rules=rules, seed=3)
    env = sb.SnakeVecEnv(N, **kw)
    co = c_oracle.COracle(N, **kw)
    assert np.array_equal(env.reset().cpu().numpy(), co.reset())
    for t in range(60):
        a = c_oracle.gen_actions(co.cfg, t, 5, env.action_space.n)
        _compare_step(env, co, a, "64x64 %s step %d" % (rules, t), check_state=(t % 10 == 0))
    env.close()


def test_full_size_properties(sb):
    """BASELINE configs[3] at its per-GPU size (131072 envs): spot-check against the oracle and
    size-independent properties (obs is a pure function of state; rewards/dones consistent)."""
    import torch
    N = 131072
    kw = dict(size=19, n_snakes=2, rules="classic", seed=0)
    env = sb.SnakeVecEnv(N, **kw)
    co = c_oracle.COracle(N, **kw)
    assert np.array_equal(env.reset().cpu().numpy(), co.reset())
    for t in range(40):
        a = env.gen_actions(t, 1)
        obs, rew, done, infos = env.step(a)
        cobs, crew, cdone, cinfo = co.step(a.cpu().numpy(), want_obs=(t % 13 == 0))
        assert np.array_equal(rew.cpu().numpy(), crew) and np.array_equal(done.cpu().numpy(), cdone)
        if t % 13 == 0:
            assert np.array_equal(obs.cpu().numpy(), cobs)
    _state_equal(env.dump_state(), co.state(), "full size")
    # idempotence: re-encoding the observation from the state gives the same bytes (masked reset of nothing)
    before = env.obs.clone()
    env.reset(mask=torch.zeros(N, dtype=torch.bool))
    assert torch.equal(before, env.obs)
    env.close()


def test_shard_invariance(sb):
    """Env i's trajectory depends on its GLOBAL id only: two shards == one handle (multi-GPU sharding)."""
    N, S = 600, 2
    kw = dict(size=10, n_snakes=S, rules="classic", seed=21)
    whole = sb.SnakeVecEnv(N, env_id_base=0, **kw)
    a_env = sb.SnakeVecEnv(250, env_id_base=0, **kw)
    b_env = sb.SnakeVecEnv(350, env_id_base=250, **kw)
    import torch
    cat = lambda x, y: torch.cat([x, y]).cpu().numpy()
    assert np.array_equal(whole.reset().cpu().numpy(), cat(a_env.reset(), b_env.reset()))
    for t in range(100):
        aw = whole.gen_actions(t, 4).clone()
        assert np.array_equal(aw.cpu().numpy(), cat(a_env.gen_actions(t, 4), b_env.gen_actions(t, 4)))
        ow, rw, dw, _ = whole.step(aw)
        oa, ra, da, _ = a_env.step(aw[:250].contiguous())
        ob, rb, db, _ = b_env.step(aw[250:].contiguous())
        assert np.array_equal(ow.cpu().numpy(), cat(oa, ob))
        assert np.array_equal(rw.cpu().numpy(), cat(ra, rb)) and np.array_equal(dw.cpu().numpy(), cat(da, db))
    sw, sa, sb_ = whole.stats(False), a_env.stats(False), b_env.stats(False)
    for k in sw:
        assert abs(sw[k] - (sa[k] + sb_[k])) < 1e-6
    for e in (whole, a_env, b_env):
        e.close()


@pytest.mark.parametrize("variant", ["fused", "ws", "split"])
def test_lane_kernel_variants_match_oracle(sb, variant):
    """The alternative forms of the lane path (warp-specialised single kernel; logic + paint as two
    kernels) produce the same bytes as the default fused form."""
    for rules, S, D, N, steps in (("classic", 2, 19, 1000, 120), ("adversarial", 3, 10, 300, 150), ("cut", 3, 10, 300, 150)):
        kw = dict(size=D, n_snakes=S, rules=rules, seed=13)
        env = sb.SnakeVecEnv(N, debug="lane=" + variant, **kw)
        co = c_oracle.COracle(N, **kw)
        assert np.array_equal(env.reset().cpu().numpy(), co.reset())
        for t in range(steps):
            a = c_oracle.gen_actions(co.cfg, t, 5, env.action_space.n)
            _compare_step(env, co, a, "%s %s step %d" % (variant, rules, t), check_state=(t % 10 == 0))
        env.close()


@pytest.mark.parametrize("thr,variant", [(1, "fused"), (3, "fused"), (1, "ws"), (2, "split")])
def test_lane_restore_unpaint_matches_oracle(sb, thr, variant):
    """Un-paint by zero-fill + border redraw (Params::restore_thr) instead of the second chain walk: same bytes.
    thr = 1 restores after almost every image, thr = 3 mixes both forms inside one launch; covers 1 view (3-byte
    pixels, byte-wide border stores), 2 and 3 views, all TE / LPE shapes and long injected bodies."""
    dbg = "restore_thr=%d,lane=%s" % (thr, variant)
    for rules, S, D, N, steps in (("classic", 2, 19, 1000, 150), ("classic", 1, 10, 500, 100), ("classic", 1, 2, 70, 40),
                                  ("adversarial", 3, 10, 300, 150), ("cut", 3, 14, 300, 150), ("classic", 4, 14, 100, 100)):
        kw = dict(size=D, n_snakes=S, rules=rules, seed=17)
        env = sb.SnakeVecEnv(N, debug=dbg, **kw)
        assert env.launch_info()["kind"] == 0 and env.launch_info()["kernel"] == {
            "fused": "k_step_lane", "ws": "k_step_lane_ws", "split": "k_lane_logic + k_lane_paint2"}[variant], (rules, S, D, env.launch_info())
        co = c_oracle.COracle(N, **kw)
        assert np.array_equal(env.reset().cpu().numpy(), co.reset())
        for t in range(steps):
            a = c_oracle.gen_actions(co.cfg, t, 5, env.action_space.n)
            _compare_step(env, co, a, "restore thr %d %s S%d D%d step %d" % (thr, rules, S, D, t), check_state=(t % 25 == 0))
        env.close()
    # long bodies: 20-70 segments
    N = 96
    kw = dict(size=19, n_snakes=2, rules="classic", seed=31)
    env = sb.SnakeVecEnv(N, debug=dbg, **kw)
    co = c_oracle.COracle(N, **kw)
    rng = np.random.RandomState(9)
    blob = helpers.random_long_snake_states(co.lay, co.cfg, rng)
    co.load_state(blob)
    env.load_state_blob(blob)
    import torch
    env.reset(mask=torch.zeros(N, dtype=torch.bool, device=env.device))
    assert np.array_equal(env.obs.cpu().numpy(), co.observe())
    for t in range(80):
        a = rng.randint(0, 5, size=(N, 2)).astype(np.int8)
        a[rng.rand(N, 2) < 0.6] = 0
        _compare_step(env, co, a, "restore long step %d" % t, check_state=(t % 20 == 0))
    env.close()


@pytest.mark.parametrize("rules,S,D,kernel", [("classic", 2, 19, None), ("cut", 3, 19, None), ("adversarial", 2, 14, None),
                                               ("classic", 2, 19, "tile"), ("cut", 3, 14, "dense"), ("classic", 4, 30, "rows")])
def test_injected_long_snakes(sb, rules, S, D, kernel):
    """Random long bodies (20-70 segments) loaded through snk_load_state into device and oracle, then
    stepped with actions biased to keep moving: multi-word chain codes, carries, cuts of long bodies."""
    N = 96
    kw = dict(size=D, n_snakes=S, rules=rules, seed=31, n_views=4 if kernel == "rows" else None)
    env = sb.SnakeVecEnv(N, debug=("force_kernel=" + kernel) if kernel else "", **kw)
    co = c_oracle.COracle(N, **kw)
    rng = np.random.RandomState(8)
    blob = helpers.random_long_snake_states(co.lay, co.cfg, rng)
    co.load_state(blob)
    env.load_state_blob(blob)
    env.check_errors()
    _state_equal(env.dump_state(), co.state(), "after load")
    import torch
    env.reset(mask=torch.zeros(N, dtype=torch.bool, device=env.device))
    assert np.array_equal(env.obs.cpu().numpy(), co.observe())
    nact = env.action_space.n
    longest = 0
    for t in range(120):
        a = rng.randint(0, nact, size=(N, S)).astype(np.int8)
        a[rng.rand(N, S) < 0.6] = 0  # mostly keep going straight so that long bodies survive a while
        _compare_step(env, co, a, "long %s %s step %d" % (rules, kernel, t), check_state=(t % 5 == 0))
        longest = max(longest, int(co.state()["len"].max()))
    assert longest >= 40
    env.close()


def test_cut_equals_classic_without_strikes_full_size(sb):
    """Size-independent property at BASELINE configs[2] size (65536 envs, 3 snakes, 10x10): with actions
    in {0..4} the cut rule-set produces exactly the classic observations, rewards and dones."""
    import torch
    N = 65536
    kw = dict(size=10, n_snakes=3, seed=2)
    a_env = sb.SnakeVecEnv(N, rules="classic", **kw)
    b_env = sb.SnakeVecEnv(N, rules="cut", **kw)
    assert torch.equal(a_env.reset(), b_env.reset())
    acts = torch.empty((N, 3), dtype=torch.int8, device=a_env.device)
    for t in range(60):
        a_env.gen_actions(t, 9, out=acts)  # classic action space: values 0..4
        oa, ra, da, _ = a_env.step(acts)
        ob, rb, db, _ = b_env.step(acts)
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db), t
    sa, sb_ = a_env.dump_state(), b_env.dump_state()
    for k in ("t", "len", "grow_to", "vel", "body", "draw_ctr"):
        assert np.array_equal(sa[k], sb_[k]), k
    a_env.close(); b_env.close()


def test_scripted_policy_stream_vs_oracle(sb):
    """The on-device fruit-seeking policy (bench.py's second action stream) keeps snakes long; the
    env under it stays bit-exact with the oracle fed the same actions."""
    N = 512
    kw = dict(size=19, n_snakes=2, rules="classic", seed=8)
    env = sb.SnakeVecEnv(N, **kw)
    co = c_oracle.COracle(N, **kw)
    assert np.array_equal(env.reset().cpu().numpy(), co.reset())
    longest = 0
    for t in range(600):
        a = env.gen_scripted_actions(t, seed=3, eps=0.05).cpu().numpy()
        assert a.min() >= 0 and a.max() <= 4
        _compare_step(env, co, a, "scripted stream step %d" % t, check_state=(t % 50 == 0))
        if t % 50 == 0:
            longest = max(longest, int(co.state()["len"].max()))
    assert longest >= 25
    s = env.stats(reduce=False)
    assert s["body_cells"] / s["env_steps"] > 8.0  # random policy: about 4.3
    env.close()


def test_time_limit_episodes(sb):
    """max_steps: done by the step cap (reference: t >= 2000, snake_multiple_test.py:195), reward 0."""
    N = 256
    kw = dict(size=19, n_snakes=2, seed=12, max_steps=6)
    env = sb.SnakeVecEnv(N, **kw)
    co = c_oracle.COracle(N, **kw)
    assert np.array_equal(env.reset().cpu().numpy(), co.reset())
    capped = 0
    for t in range(40):
        a = np.zeros((N, 2), dtype=np.int8)  # nobody moves: only spawn overlaps and the cap end episodes
        _compare_step(env, co, a, "cap step %d" % t)
        d = env._done_u8.cpu().numpy().astype(bool)
        capped += int((d & (env.rewards.cpu().numpy() == 0)).sum())
    assert capped >= N * 5
    env.close()
