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
        assert isinstance(oh, np.ndarray) and oh.dtype == np.uint8 and rh.dtype == np.float32 and dh.dtype == bool
        assert np.array_equal(od.cpu().numpy(), oh) and np.array_equal(rd.cpu().numpy(), rh) and np.array_equal(dd.cpu().numpy(), dh)
        assert ih[0]["num_snakes"] == int(dev.num_alive[0])
    dev.close(); host.close()


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
