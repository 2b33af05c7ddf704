"""GPU: parity at the BASELINE.json sizes of every configuration.

The small-N parity tests (test_gpu_parity.py) leave grid-stride tails, the `n_batches > grid` paths and the 64-bit
offset arithmetic of the big buffers unexercised.  Here the CUDA path runs at full size and is compared with the C
oracle either on every env (where the oracle finishes in seconds) or on contiguous blocks spread over the batch --
first, last and evenly spaced ones -- that the oracle steps from reset (RNG and actions are keyed by the global env id).
Everything bit-exact: observations, rewards, dones, num_snakes, Monitor r / l, full state."""
import numpy as np
import pytest

import c_oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import snakes_b200
    return snakes_b200


def _check_blocks(env, bo, results, ctx, obs=True, state=False):
    for (first, count, co), (cobs, crew, cdone, cinfo) in zip(bo.blocks, results):
        sl = slice(first, first + count)
        c = "%s block @%d" % (ctx, first)
        assert np.array_equal(env.rewards[sl].cpu().numpy(), crew), c + " reward"
        assert np.array_equal(env._done_u8[sl].cpu().numpy().astype(bool), cdone), c + " done"
        assert np.array_equal(env.num_alive[sl].cpu().numpy(), cinfo["num_snakes"]), c + " num_snakes"
        assert np.array_equal(env.rewards_all[sl].cpu().numpy(), cinfo["rewards_all"]), c + " rewards_all"
        assert np.array_equal(env.episode_return[sl].cpu().numpy(), cinfo["episode_r"]), c + " episode r"
        assert np.array_equal(env.episode_len[sl].cpu().numpy(), cinfo["episode_l"]), c + " episode l"
        if obs:
            o = env.obs[sl].cpu().numpy()
            if not np.array_equal(o, cobs):
                bad = np.flatnonzero((o != cobs).reshape(count, -1).any(1))
                raise AssertionError("%s obs differs in envs %s" % (c, (bad[:8] + first).tolist()))
        if state:
            dev, cpu = env.dump_state_range(first, count), co.state()
            for k in cpu:
                assert np.array_equal(dev[k], cpu[k]), "%s state field %r" % (c, k)


def test_cut_with_strikes_65536_envs(sb):
    """BASELINE configs[2] at its size: 3 snakes on 10x10, cut rules, 65536 envs, the 6-action stream (one action in
    six is a strike), 60 steps; EVERY env against the oracle, observations every step."""
    import torch
    N, steps = 65536, 60
    kw = dict(size=10, n_snakes=3, rules="cut", seed=42, env_id_base=1 << 20)
    env = sb.SnakeVecEnv(N, **kw)
    co = c_oracle.COracle(N, **kw)
    assert np.array_equal(env.reset().cpu().numpy(), co.reset())
    strikes = 0
    for t in range(steps):
        a = env.gen_actions(t, 77)
        an = a.cpu().numpy()
        if t == 0:
            assert np.array_equal(an, c_oracle.gen_actions(co.cfg, t, 77, 6))
        strikes += int((an == 5).sum())
        obs, rew, done, _ = env.step(a)
        cobs, crew, cdone, cinfo = co.step(an)
        ctx = "cut 65536 step %d" % t
        assert np.array_equal(rew.cpu().numpy(), crew), ctx
        assert np.array_equal(done.cpu().numpy(), cdone), ctx
        assert np.array_equal(env.num_alive.cpu().numpy(), cinfo["num_snakes"]), ctx
        assert np.array_equal(env.rewards_all.cpu().numpy(), cinfo["rewards_all"]), ctx
        assert torch.equal(obs.cpu(), torch.from_numpy(cobs)), ctx
    assert strikes > N * steps * 3 // 8
    dev, cpu = env.dump_state(), co.state()
    for k in cpu:
        assert np.array_equal(dev[k], cpu[k]), k
    stats = env.stats(reduce=False)
    assert np.allclose([stats[k] for k in sb._lib.STAT_NAMES], co.stats())
    env.check_errors()
    assert co.errors() == 0
    env.close()


def test_headline_300_steps_every_step(sb):
    """BASELINE configs[3] per-GPU shard (131072 envs of 2x19x19) for 300 steps: rewards / dones / num_snakes of EVERY
    env and the observations of 4096 envs (16 blocks spread over the batch) compared at EVERY step; full state of all
    envs at the end."""
    N, steps = 131072, 300
    kw = dict(size=19, n_snakes=2, rules="classic", seed=0)
    env = sb.SnakeVecEnv(N, **kw)
    co = c_oracle.COracle(N, **kw)
    bo = c_oracle.BlockOracles(c_oracle.BlockOracles.spread(N, 16, 256), **kw)
    env.reset(); co.reset(want_obs=False)
    for (first, count, _), cobs in zip(bo.blocks, bo.reset()):
        assert np.array_equal(env.obs[first:first + count].cpu().numpy(), cobs)
    for t in range(steps):
        a = env.gen_actions(t, 1)
        _, rew, done, _ = env.step(a)
        _, crew, cdone, cinfo = co.step(a.cpu().numpy(), want_obs=False)
        ctx = "headline step %d" % t
        assert np.array_equal(rew.cpu().numpy(), crew) and np.array_equal(done.cpu().numpy(), cdone), ctx
        assert np.array_equal(env.num_alive.cpu().numpy(), cinfo["num_snakes"]), ctx
        _check_blocks(env, bo, bo.step_generated(t, 1, 5), ctx)
    dev, cpu = env.dump_state(), co.state()
    for k in cpu:
        assert np.array_equal(dev[k], cpu[k]), k
    env.check_errors()
    env.close()


@pytest.mark.parametrize("N,steps,n_blocks,count", [(32768, 12, 16, 256), (262144, 3, 16, 256)])
def test_large_field_cut_at_baseline_sizes(sb, N, steps, n_blocks, count):
    """BASELINE configs[4]: 16 snakes on 64x64 with the cut rules at the per-GPU shard of an 8-GPU run (32768 envs) and
    at the whole 262144 envs on one GPU (55 GB of observations, 34 GB of body rings: offsets beyond 2^32).  4096 envs in
    16 blocks (first, last, evenly spaced) are compared with the oracle at every step: observations, rewards, dones,
    num_snakes, Monitor stats, and the full state of the blocks at the end."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    need = N * (209088 + 16 * 4104 * 2 + 4400) * 1.05
    if free < need:
        pytest.skip("needs %.0f GB of free HBM" % (need / 1e9))
    kw = dict(size=64, n_snakes=16, rules="cut", seed=9)
    env = sb.SnakeVecEnv(N, **kw)
    assert env.launch_info()["kernel"] == "k_step_rows"
    bo = c_oracle.BlockOracles(c_oracle.BlockOracles.spread(N, n_blocks, count), **kw)
    env.reset()
    for (first, cnt, _), cobs in zip(bo.blocks, bo.reset()):
        assert np.array_equal(env.obs[first:first + cnt].cpu().numpy(), cobs), "reset obs block @%d" % first
    for t in range(steps):
        env.step(env.gen_actions(t, 3))
        _check_blocks(env, bo, bo.step_generated(t, 3, 6), "64x64 cut N=%d step %d" % (N, t), state=(t == steps - 1))
    env.check_errors()
    s = env.stats(reduce=False)
    assert s["env_steps"] == float(N) * steps
    env.close()


def test_small_batch_4096_envs_graph(sb):
    """BASELINE configs[1] (2 snakes on 10x10, 4096 envs) stepped through ONE CUDA graph launch of 64 steps into rollout
    buffers, against the oracle: the graph path is the bench's step loop."""
    import torch
    N, T = 4096, 64
    kw = dict(size=10, n_snakes=2, rules="classic", seed=5)
    env = sb.SnakeVecEnv(N, **kw)
    co = c_oracle.COracle(N, **kw)
    assert np.array_equal(env.reset().cpu().numpy(), co.reset())
    acts = torch.stack([env.gen_actions(t, 2).clone() for t in range(T)])
    l0 = env.launch_count()
    obs, rews, dones = env.rollout(acts)
    assert env.launch_count() - l0 == T
    an = acts.cpu().numpy()
    for t in range(T):
        cobs, crew, cdone, _ = co.step(an[t])
        assert np.array_equal(obs[t].cpu().numpy(), cobs), t
        assert np.array_equal(rews[t].cpu().numpy(), crew) and np.array_equal(dones[t].cpu().numpy(), cdone), t
    # a second launch of the same (cached) graph continues the trajectories
    obs, rews, dones = env.rollout(acts, obs, rews, dones)
    for t in range(T):
        cobs, crew, cdone, _ = co.step(an[t])
        assert np.array_equal(obs[t].cpu().numpy(), cobs), t
    dev, cpu = env.dump_state(), co.state()
    for k in cpu:
        assert np.array_equal(dev[k], cpu[k]), k
    env.close()
