"""GPU: the Gym / VecEnv surface (shapes, dtypes, auto-reset contract, infos, host I/O, obs target)."""
import numpy as np
import pytest

import c_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    import snakes_b200
    return snakes_b200


def test_vecenv_surface(sb):
    import torch
    env = sb.make_basic_env("snake-multiple-test-v0", 64, seed=0)
    assert env.num_envs == 64 and env.action_space.n == 5
    assert env.observation_space.shape == (21, 21, 9) and env.observation_space.dtype == np.uint8
    obs = env.reset()
    assert obs.shape == (64, 21, 21, 9) and obs.dtype == torch.uint8 and obs.is_cuda
    # list-of-tuples actions, as MultiModel.multi_step builds them (ppo_multi_agent_new.py:35-37)
    actions = list(zip([1] * 64, [2] * 64))
    env.step_async(actions)
    with pytest.raises(sb.SnkError):
        env.step_async(actions)
    obs, rews, dones, infos = env.step_wait()
    with pytest.raises(sb.SnkError):
        env.step_wait()
    assert rews.shape == (64,) and rews.dtype == torch.float32
    assert dones.shape == (64,) and dones.dtype == torch.bool
    assert len(infos) == 64 and infos[0]["ale.lives"] == 1 and "num_snakes" in infos[0]
    # main view / opponent view slicing used by the learner (ppo_multi_agent_new.py:161-165)
    assert obs[:, :, :, 0:3].shape == (64, 21, 21, 3)
    env.close()


def test_auto_reset_contract_and_episode_info(sb):
    """subproc_vec_env.py:13-16: on done the env resets at once; obs is the first obs of the new
    episode while reward/done/info belong to the terminal step; Monitor adds info['episode']."""
    N = 256
    env = sb.SnakeVecEnv(N, size=10, n_snakes=2, seed=3)
    env.reset()
    seen = 0
    for t in range(60):
        obs, rews, dones, infos = env.step(env.gen_actions(t, 2))
        st = env.dump_state()
        d = dones.cpu().numpy()
        assert (st["t"][d] == 0).all() and (st["ep_len"][d] == 0).all()  # already reset
        assert (st["len"][d] == 1).all()
        eps = infos.episodes()
        assert len(eps) == d.sum()
        for i in np.flatnonzero(d)[:4]:
            info = infos[int(i)]
            assert info["episode"]["l"] >= 1 and info["episode"]["r"] >= -1.0
            seen += 1
        for i in np.flatnonzero(~d)[:2]:
            assert "episode" not in infos[int(i)]
    assert seen > 10
    s = env.stats()
    assert s["env_steps"] == 60 * N and s["episodes"] > 0 and s["length_sum"] > 0
    env.close()


def test_host_io_matches_device_path(sb):
    """snk_step_host (numpy in / numpy out, like SubprocVecEnv.step_wait) == device path."""
    kw = dict(size=10, n_snakes=2, seed=9)
    dev = sb.SnakeVecEnv(128, **kw)
    host = sb.SnakeVecEnv(128, host_io=True, **kw)
    o1, o2 = dev.reset(), host.reset()
    assert isinstance(o2, np.ndarray) and np.array_equal(o1.cpu().numpy(), o2)
    for t in range(50):
        a = dev.gen_actions(t, 7).cpu().numpy()
        od, rd, dd, _ = dev.step(a)
        oh, rh, dh, ih = host.step(a)
        assert isinstance(oh, np.ndarray) and oh.dtype == np.uint8 and rh.dtype == np.float64 and dh.dtype == bool
        assert np.array_equal(od.cpu().numpy(), oh) and np.array_equal(rd.cpu().numpy(), rh) and np.array_equal(dd.cpu().numpy(), dh)
        assert ih[0]["num_snakes"] == int(dev.num_alive[0])
    dev.close(); host.close()


def test_host_io_returns_fresh_arrays(sb):
    """SubprocVecEnv.step_wait returns new np.stack arrays every step and the reference Runner appends them without a
    copy (ppo_multi_agent_new.py:196 `mb_rewards.append(rewards)`): results of earlier steps must not change when the
    env steps again.  host_copy=False (views of the pinned staging buffers) is the documented opt-out for obs only."""
    kw = dict(size=10, n_snakes=2, seed=9)
    host = sb.SnakeVecEnv(256, host_io=True, **kw)
    dev = sb.SnakeVecEnv(256, **kw)
    host.reset(); dev.reset()
    kept, want = [], []
    for t in range(12):
        a = dev.gen_actions(t, 7).cpu().numpy()
        o, r, d, info = host.step(a)
        od, rd, dd, _ = dev.step(a)
        kept.append((o, r, d, info))
        want.append((od.cpu().numpy(), rd.cpu().numpy().astype(np.float64), dd.cpu().numpy(), dev.num_alive.cpu().numpy()))
    assert len({id(k[0]) for k in kept}) == 12 and any(not np.array_equal(kept[0][1], k[1]) for k in kept[1:])
    for (o, r, d, info), (wo, wr, wd, wa) in zip(kept, want):
        assert np.array_equal(o, wo) and np.array_equal(r, wr) and np.array_equal(d, wd)
        assert [info[i]["num_snakes"] for i in range(0, 256, 37)] == [int(wa[i]) for i in range(0, 256, 37)]
    view = sb.SnakeVecEnv(64, host_io=True, host_copy=False, **kw)
    view.reset()
    o1 = view.step(np.zeros((64, 2), dtype=np.int8))[0]
    o2 = view.step(np.ones((64, 2), dtype=np.int8))[0]
    assert o1 is o2 or np.shares_memory(o1, o2)
    assert view.host_numa_node is not None
    for e in (host, dev, view):
        e.close()


def test_host_io_main_view_only(sb):
    """host_views=1: only the main snake's view crosses PCIe (ppo_multi_agent_new.py:181 keeps obs[..., 0:3] alone)."""
    kw = dict(size=19, n_snakes=2, seed=4)
    full = sb.SnakeVecEnv(300, **kw)
    main = sb.SnakeVecEnv(300, host_io=True, host_views=1, **kw)
    assert main.observation_space.shape == (21, 21, 3)
    assert np.array_equal(full.reset().cpu().numpy()[..., 0:3], main.reset())
    for t in range(30):
        a = full.gen_actions(t, 3).cpu().numpy()
        of, rf, df, _ = full.step(a)
        om, rm, dm, _ = main.step(a)
        assert om.shape == (300, 21, 21, 3) and np.array_equal(of.cpu().numpy()[..., 0:3], om), t
        assert np.array_equal(rf.cpu().numpy().astype(np.float64), rm) and np.array_equal(df.cpu().numpy(), dm)
    full.close(); main.close()


def test_step_scalars_pipeline(sb):
    """Observations stay in HBM, actions in / scalars out through two pinned slots with the copies on their own streams:
    every ticket reports its own step (a slot is not reused before its trip is over) and the device state follows the
    same trajectory as the plain device path."""
    kw = dict(size=10, n_snakes=2, seed=5)
    N, T = 2048, 40
    ref = sb.SnakeVecEnv(N, **kw)
    env = sb.SnakeVecEnv(N, **kw)
    ref.reset(); env.reset()
    acts = [ref.gen_actions(t, 11).cpu().numpy() for t in range(T)]
    want = []
    for t in range(T):
        _, r, d, _ = ref.step(acts[t])
        want.append((r.cpu().numpy(), d.cpu().numpy().astype(np.uint8), ref.num_alive.cpu().numpy(),
                     ref.episode_return.cpu().numpy(), ref.episode_len.cpu().numpy()))
    tickets, got = [], {}
    for t in range(T):
        tickets.append(env.step_scalars_async(acts[t]))
        if t >= 1:   # read one step late, as a pipelined learner would
            got[tickets[t - 1]] = tuple(x.copy() for x in env.wait_scalars(tickets[t - 1]))
    with pytest.raises(sb.SnkError):
        env.wait_scalars(tickets[0])          # long gone
    got[tickets[-1]] = tuple(x.copy() for x in env.wait_scalars(tickets[-1]))
    # a second pass that reads as late as the ring allows: no slot is recycled early
    more = [ref.gen_actions(T + t, 11).cpu().numpy() for t in range(12)]
    for t in range(12):
        _, r, d, _ = ref.step(more[t])
        want.append((r.cpu().numpy(), d.cpu().numpy().astype(np.uint8), ref.num_alive.cpu().numpy(),
                     ref.episode_return.cpu().numpy(), ref.episode_len.cpu().numpy()))
        tickets.append(env.step_scalars_async(more[t]))
        if t >= env.SCALAR_SLOTS - 1:
            tk = tickets[-env.SCALAR_SLOTS]
            got[tk] = tuple(x.copy() for x in env.wait_scalars(tk))
    for tk in tickets[-(env.SCALAR_SLOTS - 1):]:
        got[tk] = tuple(x.copy() for x in env.wait_scalars(tk))
    T = T + 12
    for t in range(T):
        r, d, n, er, el = got[tickets[t]]
        assert r.dtype == np.float32 and np.array_equal(r, want[t][0]) and np.array_equal(d, want[t][1]) and np.array_equal(n, want[t][2]), t
        m = d.astype(bool)
        assert np.array_equal(er[m], want[t][3][m]) and np.array_equal(el[m], want[t][4][m]), t
    assert np.array_equal(env.obs.cpu().numpy(), ref.obs.cpu().numpy())
    a, b = env.dump_state(), ref.dump_state()
    assert all(np.array_equal(a[k], b[k]) for k in a)
    ref.close(); env.close()


def test_infos_describe_their_own_step(sb):
    """An Infos object read AFTER the next step still reports its own step (SubprocVecEnv infos are immutable)."""
    N = 512
    env = sb.SnakeVecEnv(N, size=10, n_snakes=2, seed=3)
    env.reset()
    held = []
    for t in range(25):
        _, _, dones, infos = env.step(env.gen_actions(t, 2))
        held.append((infos, dones.cpu().numpy().copy(), env.num_alive.cpu().numpy().copy(),
                     env.episode_return.cpu().numpy().copy(), env.episode_len.cpu().numpy().copy()))
    assert sum(int(h[1].sum()) for h in held) > 50
    for infos, d, alive, ret, length in held:  # all read late
        eps = infos.episodes()
        assert len(eps) == int(d.sum())
        for i in list(np.flatnonzero(d)[:5]) + [0, N - 1]:
            info = infos[int(i)]
            assert info["num_snakes"] == int(alive[i])
            assert ("episode" in info) == bool(d[i])
            if d[i]:
                assert info["episode"]["l"] == int(length[i]) and info["episode"]["r"] == round(float(ret[i]), 6)
    env.close()


def test_load_state_rejects_bad_blobs(sb):
    """snk_load_state validates the blob: too-long bodies, cell ids off the padded grid and broken
    adjacency return SNK_ESTATE instead of writing out of bounds."""
    env = sb.SnakeVecEnv(8, size=10, n_snakes=2, seed=1)
    env.reset()
    env.step(env.gen_actions(0, 1))
    good = env.dump_state_blob()
    for field, value in (("len", 5000), ("body", 60000), ("vel", 9)):
        blob = good.copy()
        st = sb.split_state(blob, env.lay, env.cfg)
        if field == "body":
            st["body"][3, 1, 0] = value
        else:
            st[field][3, 1] = value
        with pytest.raises(sb.SnkError):
            env.load_state_blob(blob)
    blob = good.copy()
    st = sb.split_state(blob, env.lay, env.cfg)
    st["len"][2, 0] = 3
    st["body"][2, 0, :3] = [14, 15, 40]   # 15 -> 40 is not a neighbour step
    with pytest.raises(sb.SnkError):
        env.load_state_blob(blob)
    env.load_state_blob(good)              # a valid blob still loads, and the error flag was consumed
    env.check_errors()
    env.close()


def test_dump_state_range_equals_slice(sb):
    for rules, S, D, N in (("classic", 2, 19, 1000), ("cut", 3, 10, 700), ("adversarial", 5, 12, 90)):
        env = sb.SnakeVecEnv(N, size=D, n_snakes=S, rules=rules, seed=6, env_id_base=100)
        env.reset()
        for t in range(20):
            env.step(env.gen_actions(t, 4))
        whole = env.dump_state()
        for first, count in ((0, 1), (0, N), (N - 33, 33), (N // 3, 200 if N > 500 else 40)):
            part = env.dump_state_range(first, count)
            for k in whole:
                assert np.array_equal(part[k], whole[k][first:first + count]), (rules, first, count, k)
        with pytest.raises(sb.SnkError):
            env.dump_state_range(N - 1, 2)
        env.close()


@pytest.mark.parametrize("mode", ["p2p", "nccl"])
def test_single_rank_comm_and_graph(sb, mode):
    """The per-step statistics reduction with one rank, in both forms (peer-memory pushes fused into the step kernel;
    ncclAllReduce on a side stream): it runs once per step, also inside a CUDA graph, changes no result, and
    stats_global() equals the local sums.  (Two ranks: tests/test_gpu_multirank.py.)"""
    import torch
    N, T = 2048, 24
    env = sb.SnakeVecEnv(N, size=10, n_snakes=2, seed=2)
    ref = sb.SnakeVecEnv(N, size=10, n_snakes=2, seed=2)
    env.reset(); ref.reset()
    assert env.init_comm(mode=mode) == 1
    acts = torch.stack([ref.gen_actions(t, 5).clone() for t in range(T)])
    for t in range(5):
        env.step(acts[t]); ref.step(acts[t])
    assert env.comm_info()["allreduces"] == 5 and env.comm_info()["ranks"] == 1
    assert env.stats_global() == ref.stats(reduce=False)
    g = env.make_graph(acts[5:])
    g.launch()
    for t in range(5, T):
        ref.step(acts[t])
    assert torch.equal(env.obs, ref.obs)
    assert env.comm_info()["allreduces"] == T
    assert env.stats_global() == ref.stats(reduce=False)
    env.step(acts[0]); ref.step(acts[0])
    assert env.stats_global() == ref.stats(reduce=False) and torch.equal(env.obs, ref.obs)
    g.close(); env.close(); ref.close()


def test_main_view_target(sb):
    """snk_set_main_view_target: view 0 of every step, packed [N,H,W,3], lands in the caller's slot while all K views
    stay in the env's own buffer (f1: the learner's rollout keeps the main view only, ppo_multi_agent_new.py:181)."""
    import torch
    cases = [(203, kw) for kw in (dict(size=19, n_snakes=2), dict(size=10, n_snakes=3, rules="cut"), dict(size=19, n_snakes=2, obs_mode="atari84"))]
    # 16-byte aligned slots and whole 16-pixel groups: the vectorised gather (k_extract_main) for 6 / 9 / 12 bytes per pixel
    cases += [(256, dict(size=19, n_snakes=2)), (256, dict(size=10, n_snakes=3, rules="cut")), (256, dict(size=10, n_snakes=4)),
              (272, dict(size=19, n_snakes=2, obs_mode="atari84"))]
    for N, kw in cases:
        env = sb.SnakeVecEnv(N, seed=5, **kw)
        H = env.obs.shape[1]
        slots = torch.zeros((5, N, H, H, 3), dtype=torch.uint8, device=env.device)
        env.set_main_view_target(slots[0])
        obs = env.reset()
        assert torch.equal(slots[0], obs[..., 0:3])
        for t in range(1, 5):
            env.set_main_view_target(slots[t])
            obs, _, _, _ = env.step(env.gen_actions(t, 2))
            assert torch.equal(slots[t], obs[..., 0:3]), t
            assert obs.shape[-1] == 3 * env.K
        env.set_main_view_target(None)
        keep = slots[4].clone()
        env.step(env.gen_actions(9, 2))
        assert torch.equal(slots[4], keep)
        env.close()


def test_monitor_csv(sb, tmp_path):
    """MonitorCSV writes the file baselines' Monitor writes (bench/monitor.py): json header, r,l,t rows per episode."""
    import csv
    import json
    env = sb.SnakeVecEnv(64, size=10, n_snakes=2, seed=1)
    env.reset()
    mon = sb.MonitorCSV(str(tmp_path / "0"), env_id="snake-multiple-test-v0")
    total = 0
    for t in range(40):
        _, _, dones, infos = env.step(env.gen_actions(t, 3))
        mon.write(infos)
        total += int(dones.sum())
    mon.close()
    lines = open(mon.path).read().splitlines()
    assert mon.path.endswith("0.monitor.csv") and lines[0].startswith("#") and json.loads(lines[0][1:])["env_id"] == "snake-multiple-test-v0"
    rows = list(csv.DictReader(lines[1:]))
    assert len(rows) == total == mon.episodes and total > 20
    assert all(int(r["l"]) >= 1 and float(r["r"]) >= -1.0 and float(r["t"]) >= 0 for r in rows)
    env.close()


def test_obs_target_rollout_slot(sb):
    """Observations written straight into a slot of a caller-owned rollout buffer."""
    import torch
    N = 96
    env = sb.SnakeVecEnv(N, size=10, n_snakes=2, seed=1)
    ref = sb.SnakeVecEnv(N, size=10, n_snakes=2, seed=1)
    rollout = torch.zeros((4, N, 12, 12, 6), dtype=torch.uint8, device=env.device)
    env.set_obs_target(rollout[0]); env.reset(); ref.reset()
    assert torch.equal(rollout[0], ref.obs)
    for t in range(3):
        env.set_obs_target(rollout[t + 1])
        a = ref.gen_actions(t, 3)
        env.step(a); ref.step(a)
        assert torch.equal(rollout[t + 1], ref.obs)
    env.close(); ref.close()


def test_single_env_gym_api(sb):
    """gym.make(id) -> reset/step/seed like evaluate_snake.py:52-117 (no auto-reset, numpy out)."""
    env = sb.make("snake-multiple-test-v0")
    env.seed(5)
    ob = env.reset()
    assert isinstance(ob, np.ndarray) and ob.shape == (21, 21, 9) and ob.dtype == np.uint8
    co = c_oracle.COracle(1, size=19, n_snakes=2, n_views=3, seed=5, auto_reset=False)
    assert np.array_equal(co.reset()[0], ob)
    done = False
    steps = 0
    rng = np.random.RandomState(0)
    while not done and steps < 500:
        a = rng.randint(0, 5, size=2)
        ob, r, done, info = env.step(a)
        cob, cr, cd, ci = co.step(a[None])
        assert np.array_equal(ob, cob[0]) and r == float(cr[0]) and done == bool(cd[0])
        assert isinstance(r, float) and isinstance(done, bool) and info["ale.lives"] == 1
        steps += 1
    assert done
    ob2 = env.reset()
    assert np.array_equal(ob2, co.reset()[0])
    # kwargs through a second __init__ call, as utils.py:38 does
    env.__init__(n_snakes=3, n_fruits=3)
    assert env.reset().shape == (12, 12, 9)
    assert env.step(1)[0].shape == (12, 12, 9)  # scalar action is wrapped (snake_multiple_test.py:167-168)
    env.close()
    for env_id in sb.ENV_IDS:
        e = sb.make(env_id)
        e.reset(); e.step([0] * e.n_snakes); e.close()


def test_create_rejects_bad_config(sb):
    with pytest.raises(sb.SnkError):
        sb.SnakeVecEnv(4, size=1)
    with pytest.raises(sb.SnkError):
        sb.SnakeVecEnv(4, n_snakes=40)
    with pytest.raises(ValueError):
        sb.SnakeVecEnv(4, size=(10, 12))
    with pytest.raises(ValueError):
        sb.SnakeVecEnv(4, rules="nope")


@pytest.mark.parametrize("D,S,K,rules", [(19, 2, 2, "classic"), (10, 3, 3, "cut"), (10, 2, 3, "adversarial"), (12, 1, 1, "classic")])
def test_atari84_obs_mode(sb, D, S, K, rules):
    """obs_mode='atari84' = the reference's WarpFrame (utils.py:27-31): exact r x r replication of the
    native image (cv2 INTER_AREA for integer factors, see tests/test_warpframe.py)."""
    import torch
    N = 200
    kw = dict(size=D, n_snakes=S, n_views=K, rules=rules, seed=4)
    env = sb.SnakeVecEnv(N, obs_mode="atari84", **kw)
    co = c_oracle.COracle(N, **kw)
    r = 84 // (D + 2)
    up = lambda o: np.repeat(np.repeat(o, r, axis=1), r, axis=2)
    assert env.observation_space.shape == (84, 84, 3 * K)
    obs = env.reset()
    assert obs.shape == (N, 84, 84, 3 * K)
    assert np.array_equal(obs.cpu().numpy(), up(co.reset()))
    for t in range(40):
        a = c_oracle.gen_actions(co.cfg, t, 8, env.action_space.n)
        obs, rew, done, _ = env.step(a)
        cobs, crew, cdone, _ = co.step(a)
        assert np.array_equal(obs.cpu().numpy(), up(cobs)), t
        assert np.array_equal(rew.cpu().numpy(), crew) and np.array_equal(done.cpu().numpy(), cdone)
    # straight into a rollout slot
    rollout = torch.zeros((2, N, 84, 84, 3 * K), dtype=torch.uint8, device=env.device)
    env.set_obs_target(rollout[1])
    a = c_oracle.gen_actions(co.cfg, 40, 8, env.action_space.n)
    env.step(a)
    assert np.array_equal(rollout[1].cpu().numpy(), up(co.step(a)[0]))
    env.close()
    with pytest.raises(sb.SnkError):
        sb.SnakeVecEnv(4, size=9, obs_mode="atari84")


@pytest.mark.parametrize("D,S,K", [(19, 2, 2), (10, 3, 3), (19, 2, 3)])
def test_atari84_many_envs_per_cta(sb, D, S, K):
    """More envs than resident CTAs: every CTA of k_upscale84 streams several images through its two-part pipeline
    (part B of env e and part A of env e + grid overlap; the next native image travels in registers)."""
    N = 4001
    kw = dict(size=D, n_snakes=S, n_views=K, rules="classic", seed=6)
    env = sb.SnakeVecEnv(N, obs_mode="atari84", **kw)
    nat = sb.SnakeVecEnv(N, **kw)
    r = 84 // (D + 2)
    up = lambda o: o.repeat_interleave(r, dim=1).repeat_interleave(r, dim=2)
    import torch
    assert torch.equal(env.reset(), up(nat.reset()))
    for t in range(5):
        a = nat.gen_actions(t, 8)
        assert torch.equal(env.step(a)[0], up(nat.step(a)[0])), t
    env.close(); nat.close()


@pytest.mark.parametrize("rules,S,D", [("classic", 2, 19), ("adversarial", 3, 10), ("cut", 16, 64)])
def test_checkpoint_resume_is_bit_exact(sb, tmp_path, rules, S, D):
    """save() / load(): the resumed env continues exactly like the original (state incl. RNG counters)."""
    N = 96 if D < 64 else 12
    env = sb.SnakeVecEnv(N, size=D, n_snakes=S, rules=rules, seed=77, env_id_base=1000)
    env.reset()
    for t in range(30):
        env.step(env.gen_actions(t, 6))
    path = str(tmp_path / "ckpt.npz")
    env.save(path)
    twin = sb.SnakeVecEnv.load(path)
    assert np.array_equal(twin.obs.cpu().numpy(), env.obs.cpu().numpy())
    for t in range(30, 60):
        a = env.gen_actions(t, 6).clone()
        o1, r1, d1, _ = env.step(a)
        o2, r2, d2, _ = twin.step(a.cpu().numpy())
        assert np.array_equal(o1.cpu().numpy(), o2.cpu().numpy()), t
        assert np.array_equal(r1.cpu().numpy(), r2.cpu().numpy()) and np.array_equal(d1.cpu().numpy(), d2.cpu().numpy())
    s1, s2 = env.dump_state(), twin.dump_state()
    for k in s1:
        assert np.array_equal(s1[k], s2[k]), k
    env.close(); twin.close()


@pytest.mark.parametrize("obs_mode,N", [("native", 256), ("atari84", 64), ("native", 100)])
def test_rollout_writer_equals_step_by_step(sb, obs_mode, N):
    """snk_rollout: T steps written straight into [T,N,...] buffers == T calls of step()."""
    import torch
    T = 12
    kw = dict(size=19, n_snakes=2, seed=6, obs_mode=obs_mode)
    a_env = sb.SnakeVecEnv(N, **kw)
    b_env = sb.SnakeVecEnv(N, **kw)
    a_env.reset(); b_env.reset()
    acts = torch.stack([a_env.gen_actions(t, 2).clone() for t in range(T)])
    if N % 8 != 0 and obs_mode == "native":
        # 100 envs x 2646 B is not a multiple of 16: slots would be misaligned for the TMA store
        with pytest.raises(sb.SnkError):
            a_env.rollout(acts)
        a_env.close(); b_env.close()
        return
    obs, rews, dones = a_env.rollout(acts)
    for t in range(T):
        o, r, d, _ = b_env.step(acts[t])
        assert torch.equal(obs[t], o), t
        assert torch.equal(rews[t], r) and torch.equal(dones[t], d), t
    assert torch.equal(a_env.obs, b_env.obs) and torch.equal(a_env.rewards, b_env.rewards)
    sa, sb_ = a_env.dump_state(), b_env.dump_state()
    for k in sa:
        assert np.array_equal(sa[k], sb_[k]), k
    a_env.close(); b_env.close()


def test_render_world_view(sb):
    """render(mode='rgb_array') == the oracle's get_ob_world restatement, for every env index asked."""
    import snake_oracle as so
    N, S, D = 8, 3, 10
    env = sb.SnakeVecEnv(N, size=D, n_snakes=S, seed=3)
    orc = so.VecOracle(N, D, S, S, S, "classic", seed=3)
    env.reset(); orc.reset()
    for t in range(25):
        a = env.gen_actions(t, 5).cpu().numpy()
        env.step(a); orc.step(a)
        for i in (0, 5):
            img = env.render("rgb_array", env_index=i)
            assert img.shape == (D + 2, D + 2, 3) and img.dtype == np.uint8
            assert np.array_equal(img, orc.envs[i].world_view()), (t, i)
    env.close()


@pytest.mark.parametrize("T,N,gamma,lam", [(64, 1000, 0.99, 0.95), (5, 33, 0.9, 1.0), (1, 7, 0.99, 0.95), (128, 4096, 0.999, 0.9)])
def test_gae_on_device_is_bit_exact(sb, T, N, gamma, lam):
    """snk_gae == the numpy loop of Runner.run (ppo_multi_agent_new.py:205-218), bit for bit."""
    import torch
    import gae_oracle
    rng = np.random.RandomState(T + N)
    rewards = rng.choice([-1.0, 0.0, 0.0, 0.0, 1.0, 2.0], size=(T, N)).astype(np.float32)
    values = rng.randn(T, N).astype(np.float32) * 3
    dones = rng.rand(T, N) < 0.1
    last_values = rng.randn(N).astype(np.float32)
    last_dones = rng.rand(N) < 0.1
    want_a, want_r = gae_oracle.gae(rewards, values, dones, last_values, last_dones, gamma, lam)
    dev = torch.device("cuda", 0)
    t = lambda x: torch.as_tensor(x, device=dev)
    got_a, got_r = sb.gae(t(rewards), t(values), t(dones), t(last_values), t(last_dones), gamma, lam)
    assert got_a.dtype == torch.float32 and tuple(got_a.shape) == (T, N)
    assert np.array_equal(got_a.cpu().numpy().view(np.uint32), want_a.view(np.uint32))
    assert np.array_equal(got_r.cpu().numpy().view(np.uint32), want_r.view(np.uint32))
