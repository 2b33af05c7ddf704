"""CPU: the oracles reproduce the recordings of the reference (tests/golden, oracle/make_golden.py)."""
import zlib

import numpy as np
import pytest

import c_oracle
import helpers
import snake_oracle as so


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kat = [(([0] * 4, [0] * 2), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           (([0xffffffff] * 4, [0xffffffff] * 2), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           (([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for (ctr, key), want in kat:
        assert so.philox4x32_10(ctr, key) == want
        assert c_oracle.philox(ctr, key) == want


@pytest.mark.parametrize("name", helpers.golden_names())
def test_c_oracle_replays_reference_recording(name):
    g = helpers.load_golden(name)
    N, T = g["N"], g["T"]
    co = c_oracle.COracle(N, size=g["D"], n_snakes=g["S"], n_fruits=g["F"], n_views=g["K"], rules=g["rules"], rng_mode=1)
    co.set_draw_tape(g["tape_vals"], g["tape_bounds"], g["tape_offsets"])
    assert zlib.crc32(co.reset().tobytes()) == int(g["reset_obs_crc"])
    for t in range(T):
        obs, rew, done, info = co.step(g["actions"][t])
        assert np.array_equal(rew, g["reward"][t]), (name, t)
        assert np.array_equal(done, g["done"][t].astype(bool)), (name, t)
        assert np.array_equal(info["num_snakes"], g["num_snakes"][t]), (name, t)
        assert np.array_equal(info["episode_r"], g["ep_r"][t]), (name, t)
        assert np.array_equal(info["episode_l"], g["ep_l"][t]), (name, t)
        assert zlib.crc32(obs.tobytes()) == int(g["obs_crc"][t]), (name, t)
        if t % 7 == 0 or t == T - 1:
            assert helpers.batch_state_crc(co.state()) == int(g["state_crc"][t]), (name, t)
    assert co.errors() == 0
    st = co.state()
    assert np.array_equal(obs, g["final_obs"])
    for k in ("t", "spare", "len", "grow_to", "vel", "body"):
        assert np.array_equal(st[k], g["final_" + k]), k
    fk = "fruit_grid" if "fruit_grid" in st else "fruit"
    assert np.array_equal(st[fk], g["final_" + fk])
    # every recorded draw was consumed, none more
    assert np.array_equal(st["draw_ctr"].astype(np.uint64), np.diff(g["tape_offsets"]))


@pytest.mark.parametrize("name", helpers.golden_names())
def test_python_oracle_replays_reference_recording(name):
    g = helpers.load_golden(name)
    lanes = list(range(0, g["N"], max(1, g["N"] // 8)))[:8]
    T = min(g["T"], 200)
    off = g["tape_offsets"].astype(np.int64)
    draws = [so.TapeDraws(g["tape_vals"][off[i]:off[i + 1]], g["tape_bounds"][off[i]:off[i + 1]]) for i in lanes]
    po = so.VecOracle(len(lanes), g["D"], g["S"], g["F"], g["K"], g["rules"], draws=draws)
    vals, bounds, o = helpers.sub_tape(g, lanes)
    co = c_oracle.COracle(len(lanes), size=g["D"], n_snakes=g["S"], n_fruits=g["F"], n_views=g["K"], rules=g["rules"], rng_mode=1)
    co.set_draw_tape(vals, bounds, o)
    assert np.array_equal(po.reset(), co.reset())
    for t in range(T):
        a = g["actions"][t][lanes]
        obs, rew, done, info = po.step(a)
        cobs, crew, cdone, cinfo = co.step(a)
        assert np.array_equal(rew, g["reward"][t][lanes]) and np.array_equal(done, g["done"][t][lanes].astype(bool))
        assert np.array_equal(info["num_snakes"], g["num_snakes"][t][lanes])
        assert np.array_equal(info["episode_r"], g["ep_r"][t][lanes]) and np.array_equal(info["episode_l"], g["ep_l"][t][lanes])
        assert np.array_equal(obs, cobs), (name, t)
        assert np.array_equal(po.state_crcs(g["cap"]), helpers.per_env_state_crcs(co.state())), (name, t)


def test_known_answers_c_oracle():
    """Hand-built states stepped once by the reference (SURVEY.md section 8c list)."""
    for ka in helpers.known_answers():
        S, D = ka["S"], ka["D"]
        co = c_oracle.COracle(1, size=D, n_snakes=S, n_fruits=len(ka["in"]["fruits"]), n_views=3, rules=ka["rules"],
                              rng_mode=1, auto_reset=False)
        i = ka["in"]
        co.load_state(helpers.state_blob_from_lists(co.lay, co.cfg, i["snakes"], i["fruits"], i["vels"], i["grow_to"], i["t"], i["spare"]))
        vals = np.array([d[1] for d in ka["draws"]], dtype=np.uint32)
        bounds = np.array([d[0] for d in ka["draws"]], dtype=np.uint32)
        co.set_draw_tape(vals, bounds, np.array([0, len(vals)], dtype=np.uint64))
        obs, rew, done, info = co.step(np.array([ka["action"]], dtype=np.int8))
        note = ka["note"]
        assert co.errors() == 0, note
        assert float(rew[0]) == ka["reward"] and bool(done[0]) == ka["done"], note
        assert int(info["num_snakes"][0]) == ka["num_snakes"], note
        assert zlib.crc32(obs[0].tobytes()) == ka["obs_crc"], note
        got = helpers.lists_from_state(co.state(), co.cfg)
        want = dict(ka["out"])
        if ka["rules"] != "classic":
            want["fruits"] = sorted(want["fruits"])
        assert got == want, note
        assert int(co.state()["draw_ctr"][0]) == len(vals), note
