"""GPU: the on-device self-play PPO loop around the env."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_single_snake_ppo_learns_on_device():
    """policy -> env step -> rollout buffer -> GAE -> PPO2 update with no host round trip: the mean episode
    return of a single snake on 10x10 must rise clearly above the random policy's (about -0.94)."""
    import torch
    import snakes_b200
    from snakes_b200 import selfplay
    torch.manual_seed(0)
    env = snakes_b200.SnakeVecEnv(1024, size=10, n_snakes=1, seed=0)
    model, log = selfplay.learn(env, nsteps=32, total_timesteps=1024 * 32 * 36, log_interval=6, seed=0)
    env.check_errors()
    first, last = log.rows[0], log.rows[-1]
    assert first["eprewmean 100"] < -0.8
    assert last["eprewmean 100"] > first["eprewmean 100"] + 0.5, (first, last)
    assert last["eplenmean"] > 3 * first["eplenmean"]
    # the reference's logger keys (ppo_multi_agent_new.py:357-371)
    for k in ("num_opponents", "serial_timesteps", "nupdates", "total_timesteps", "explained_variance", "eprewmean 100",
              "eplenmean", "time_elapsed", "ep_rew_mean", "policy_loss", "value_loss", "policy_entropy", "approxkl", "clipfrac"):
        assert k in last
    env.close()


def test_two_snake_self_play_pool_and_checkpoints(tmp_path):
    import joblib
    import torch
    import snakes_b200
    from snakes_b200 import selfplay
    torch.manual_seed(0)
    env = snakes_b200.SnakeVecEnv(256, size=10, n_snakes=2, seed=3)
    d = str(tmp_path / "ppo")
    model, log = selfplay.learn(env, nsteps=16, total_timesteps=256 * 16 * 5, log_interval=1, model_dir=d,
                                opponent_save_interval=2, csv_path=str(tmp_path / "ppo.csv"), save_interval=2)
    env.check_errors()
    # pool: initial snapshot + one per 2 updates (ppo_multi_agent_new.py:283-286, :332-341), file names of utils.get_opponent_file
    assert sorted(f for f in os.listdir(d) if f.startswith("opponent")) == ["opponent0_0.pkl", "opponent0_1.pkl", "opponent0_2.pkl"]
    assert log.rows[-1]["num_opponents"] == 3 and len(log.rows) == 5
    assert os.path.exists(os.path.join(d, "snake_model_num2_final.pkl"))
    # a checkpoint is the reference's format (list of float32 ndarrays, TF variable order) and restores the policy
    params = joblib.load(os.path.join(d, "snake_model_num2_final.pkl"))
    assert len(params) == 14 and params[0].shape == (3, 3, 3, 32)
    m2 = selfplay.Model((12, 12, 3), 5, device=env.device, trainable=False)
    m2.load(os.path.join(d, "snake_model_num2_final.pkl"))
    ob = env.obs[..., 0:3]
    assert torch.allclose(m2.net(ob)[0], model.net(ob)[0], atol=1e-5)
    env.close()


def test_runner_rollout_matches_step_by_step_env():
    """Runner.run's buffers hold exactly what the env returned: replay the recorded actions on a twin env."""
    import torch
    import snakes_b200
    from snakes_b200 import selfplay
    torch.manual_seed(0)
    kw = dict(size=10, n_snakes=2, seed=11)
    env, twin = snakes_b200.SnakeVecEnv(128, **kw), snakes_b200.SnakeVecEnv(128, **kw)
    model = selfplay.Model((12, 12, 3), 5, device=env.device)
    runner = selfplay.Runner(env, model, [None], nsteps=8)
    obs, returns, dones, actions, values, neglogp, _, ep = runner.run()
    T, N = 8, 128
    o = twin.reset()
    a_tn = actions.reshape(N, T).t()
    prev_done = torch.zeros(N, dtype=torch.bool, device=env.device)
    rew = []
    for t in range(T):
        assert torch.equal(obs.reshape(N, T, 12, 12, 3)[:, t], o[..., 0:3])
        assert torch.equal(dones.reshape(N, T)[:, t], prev_done)
        full = torch.stack([a_tn[t].to(torch.int8), torch.ones(N, dtype=torch.int8, device=env.device)], 1).contiguous()
        o, r, d, _ = twin.step(full)
        rew.append(r.clone()); prev_done = d.clone()
    # returns = GAE advantages + values over the twin's rewards (checked against the numpy loop of the reference)
    import gae_oracle
    last_v = model.value(o[..., 0:3])
    advs, rets = gae_oracle.gae(torch.stack(rew).cpu().numpy(), values.reshape(N, T).t().cpu().numpy(),
                                          dones.reshape(N, T).t().cpu().numpy(), last_v.cpu().numpy(), prev_done.cpu().numpy(), 0.99, 0.95)
    assert np.allclose(returns.reshape(N, T).t().cpu().numpy(), rets, atol=1e-5)
    env.close(); twin.close()
