"""CPU: host logic of the self-play learner -- network shapes, the PPO2 loss against a
numpy restatement of the reference graph, checkpoint interchange in TF variable order, logger keys, tiling."""
import csv

import numpy as np
import pytest
import torch

from snakes_b200 import selfplay
from snakes_b200.vec_env import tile_images


def test_policy_shapes_and_tf_variable_order():
    # custom_cnn (policies.py:23-31): 4 SAME 3x3 convs -> [H, W, 64] -> fc 512; nature_cnn (:12-20) on 84x84 -> 7x7x64
    p = selfplay.CnnPolicy((12, 12, 3), 5, "custom")
    assert p.conv_out == (12, 12, 64)
    shapes = [a.shape for a in p.to_tf_list()]
    assert shapes == [(3, 3, 3, 32), (32,), (3, 3, 32, 32), (32,), (3, 3, 32, 64), (64,), (3, 3, 64, 64), (64,),
                      (12 * 12 * 64, 512), (512,), (512, 5), (5,), (512, 1), (1,)]
    q = selfplay.CnnPolicy((84, 84, 3), 5, "nature")
    assert q.conv_out == (7, 7, 64)
    assert [a.shape for a in q.to_tf_list()][:6] == [(8, 8, 3, 32), (32,), (4, 4, 32, 64), (64,), (3, 3, 64, 64), (64,)]
    x = torch.randint(0, 256, (5, 12, 12, 3), dtype=torch.uint8)
    logits, v = p(x)
    assert logits.shape == (5, 5) and v.shape == (5,)
    a, v2, nlp = p.step(x)
    assert a.shape == (5,) and a.dtype == torch.int64 and torch.all((a >= 0) & (a < 5))
    assert torch.allclose(nlp, torch.nn.functional.cross_entropy(logits, a, reduction="none"))
    # pi is initialised with scale 0.01: a fresh policy is near-uniform (neglogp ~ ln 5)
    assert abs(float(nlp.mean()) - np.log(5)) < 0.05


def test_checkpoint_round_trip_and_tf_layout(tmp_path):
    torch.manual_seed(1)
    m = selfplay.Model((12, 12, 3), 5, device="cpu")
    m.save(str(tmp_path / "snake_model.pkl"))
    import joblib
    params = joblib.load(str(tmp_path / "snake_model.pkl"))
    assert isinstance(params, list) and all(isinstance(a, np.ndarray) and a.dtype == np.float32 for a in params)
    m2 = selfplay.Model((12, 12, 3), 5, device="cpu")
    m2.load(str(tmp_path / "snake_model.pkl"))
    x = torch.randint(0, 256, (3, 12, 12, 3), dtype=torch.uint8)
    assert torch.equal(m.net(x)[0], m2.net(x)[0])
    # the layout is TF's: conv [kh, kw, in, out] and an fc1 whose rows follow the NHWC flatten of conv_to_fc.
    # Evaluate the network by hand from the exported arrays (numpy, NHWC) and compare.
    h = x.numpy().astype(np.float32) / 255.0
    it = iter(params)
    for _ in range(4):
        w, b = next(it), next(it)
        hp = np.pad(h, ((0, 0), (1, 1), (1, 1), (0, 0)))
        out = np.zeros(h.shape[:3] + (w.shape[3],), dtype=np.float32)
        for i in range(3):
            for j in range(3):
                out += hp[:, i:i + h.shape[1], j:j + h.shape[2], :] @ w[i, j]
        h = np.maximum(out + b, 0)
    w, b = next(it), next(it)
    hid = np.maximum(h.reshape(h.shape[0], -1) @ w + b, 0)
    w, b = next(it), next(it)
    logits = hid @ w + b
    assert np.allclose(logits, m.net(x)[0].detach().numpy(), atol=2e-4)
    with pytest.raises(ValueError):
        m2.net.from_tf_list(params[:-1])


def test_ppo_loss_matches_reference_graph():
    """ppo_multi_agent_new.py:62-77 restated in numpy (float64) on random inputs."""
    rng = np.random.RandomState(0)
    B, A = 257, 5
    logits = rng.randn(B, A).astype(np.float32)
    vpred = rng.randn(B).astype(np.float32)
    actions = rng.randint(0, A, size=B)
    advs = rng.randn(B).astype(np.float32)
    returns = rng.randn(B).astype(np.float32)
    old_v = (vpred + 0.3 * rng.randn(B)).astype(np.float32)
    old_nlp = (1.6 + 0.2 * rng.randn(B)).astype(np.float32)
    clip, ent_coef, vf_coef = 0.2, 0.01, 0.5
    z = logits.astype(np.float64)
    z = z - z.max(1, keepdims=True)
    logp = z - np.log(np.exp(z).sum(1, keepdims=True))
    nlp = -logp[np.arange(B), actions]
    entropy = (-(np.exp(logp) * logp).sum(1)).mean()
    vclip = old_v + np.clip(vpred - old_v, -clip, clip)
    vf_loss = 0.5 * np.maximum((vpred - returns) ** 2, (vclip - returns) ** 2).mean()
    ratio = np.exp(old_nlp - nlp)
    pg_loss = np.maximum(-advs * ratio, -advs * np.clip(ratio, 1 - clip, 1 + clip)).mean()
    approxkl = 0.5 * ((nlp - old_nlp) ** 2).mean()
    clipfrac = (np.abs(ratio - 1) > clip).mean()
    want = pg_loss - entropy * ent_coef + vf_loss * vf_coef
    t = lambda a: torch.from_numpy(np.asarray(a))
    loss, stats = selfplay.ppo_loss(t(logits), t(vpred), t(actions), t(advs), t(returns), t(old_nlp), t(old_v), clip, ent_coef, vf_coef)
    assert abs(float(loss) - want) < 1e-5
    assert np.allclose(stats.numpy(), [pg_loss, vf_loss, entropy, approxkl, clipfrac], atol=1e-5)


def test_train_step_moves_towards_advantage():
    torch.manual_seed(0)
    m = selfplay.Model((12, 12, 3), 5, device="cpu")
    obs = torch.randint(0, 256, (64, 12, 12, 3), dtype=torch.uint8)
    a, v, nlp = m.step(obs)
    returns = v + torch.where(a == 2, 1.0, -1.0)       # action 2 is "good"
    before = torch.softmax(m.net(obs)[0], -1)[:, 2].mean()
    for _ in range(5):
        stats = m.train(1e-3, 0.2, obs, returns, None, a, v, nlp)
    after = torch.softmax(m.net(obs)[0], -1)[:, 2].mean()
    assert after > before and stats.shape == (5,)


def test_sf01_and_explained_variance():
    x = torch.arange(2 * 3 * 4).reshape(2, 3, 4)
    assert np.array_equal(selfplay.sf01(x).numpy(), x.numpy().swapaxes(0, 1).reshape(6, 4))
    y = torch.tensor([1.0, 2.0, 3.0, 4.0])
    assert selfplay.explained_variance(y, y) == 1.0
    assert np.isnan(selfplay.explained_variance(y, torch.ones(4)))


def test_kv_logger_csv(tmp_path):
    log = selfplay.KVLogger(str(tmp_path / "ppo.csv"))
    for u in (1, 10):
        log.logkv("nupdates", u); log.logkv("eprewmean 100", -0.5 * u); log.logkv("policy_loss", 0.1)
        log.dumpkvs()
    log.close()
    rows = list(csv.DictReader(open(str(tmp_path / "ppo.csv"))))
    assert [r["nupdates"] for r in rows] == ["1", "10"] and list(rows[0].keys()) == ["nupdates", "eprewmean 100", "policy_loss"]


def test_tile_images_layout():
    """Same mosaic as baselines/common/tile_images.py: P = ceil(sqrt(N)) rows, Q = ceil(N/P) columns, black padding."""
    imgs = np.stack([np.full((2, 3, 3), i + 1, dtype=np.uint8) for i in range(5)])
    big = tile_images(imgs)
    assert big.shape == (3 * 2, 2 * 3, 3)
    assert big[0, 0, 0] == 1 and big[0, 3, 0] == 2 and big[2, 0, 0] == 3 and big[2, 3, 0] == 4 and big[4, 0, 0] == 5
    assert not big[4:, 3:].any()


def test_lane_kernel_reciprocal_constants_are_exact():
    """group_spawn (snk_lane.cuh) divides padded ids by V and board indices by D with (n * (65536 // d + 1)) >> 16;
    exact for every board the lane family accepts (D <= 32, snk_api.cu launch plan)."""
    for D in range(1, 33):
        V = D + 2
        mV, mD = 65536 // V + 1, 65536 // D + 1
        assert all((n * mV) >> 16 == n // V for n in range(V * V))
        assert all((n * mD) >> 16 == n // D for n in range(D * D))
