"""Generates tests/golden/*.npz and tests/golden/known_answers.json FROM THE REFERENCE ITSELF.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference); the files it
writes are committed so that the GPU box, which has no reference tree, can still check the
oracle and the CUDA path against the reference's behaviour.

    python oracle/make_golden.py            # regenerates every fixture (about 4 minutes)

How a trajectory file is made (one per configuration):
  * N independent reference envs (SnakeEnv / SnakeAdversarial, unmodified, `dim` patched,
    Config.NUM_SNAKES = S) each wrapped in the reference's own `Monitor`
    (baselines/bench/monitor.py) and driven with the SubprocVecEnv worker contract
    (baselines/common/vec_env/subproc_vec_env.py:13-16: step, and on done reset at once and
    return the reset observation);
  * env i's np_random is a recording proxy around RandomState(seed + i): every randint(n) the
    reference makes is appended to env i's tape as (bound n, value);
  * actions: the first half of the lanes are uniform over {0..4}, the second half follow a
    fruit-seeking script with 10 % random moves, so that long bodies, eating, respawns on a
    crowded board and self-collisions are all covered;
  * per step the file stores reward, done, num_snakes, Monitor's episode r / l, and crc32
    digests of all N observations ([N, V, V, 9], the K = 3 views SnakeEnv emits) and of all N
    canonical states (oracle/snake_oracle.py:state_crc), plus full final obs / state arrays.
"""
import json
import os
import sys
import time
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402
from ref_compare import canonical_from_reference  # noqa: E402
from snake_oracle import RecordingDraws, state_crc  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")

CASES = [
    # name,              rules,         S, D,   N,   T
    ("classic_2x19",     "classic",     2, 19, 384, 1000),  # BASELINE configs[3] geometry; >= 10k episodes
    ("classic_2x10",     "classic",     2, 10,  64,  400),  # configs[1]
    ("classic_3x10",     "classic",     3, 10,  64,  400),  # configs[2] geometry, classic rules
    ("classic_1x10",     "classic",     1, 10,  32,  400),  # configs[0]
    ("classic_3x3",      "classic",     3,  3,  32,  300),  # crowded: no-free-cell and OOB-alias paths
    ("adversarial_3x10", "adversarial", 3, 10,  64,  400),
    ("adversarial_2x10", "adversarial", 2, 10,  64,  400),
]

_DIRS = {1: (1, 0), 2: (0, 1), 3: (-1, 0), 4: (0, -1)}
_OPP = {1: 3, 2: 4, 3: 1, 4: 2}
_VEL_CODE = {(0, 0): 0, (1, 0): 1, (0, 1): 2, (-1, 0): 3, (0, -1): 4}


def scripted_action(env, s, rng):
    """Greedy fruit seeking on the reference's own state; generation-time only."""
    snakes, fruits, vels = env.state[0], env.state[1], env.state[2]
    if len(snakes[s]) == 0 or rng.rand() < 0.1:
        return int(rng.randint(0, 5))
    hx, hy = snakes[s][0]
    vel = _VEL_CODE[tuple(vels[s])]
    occupied = set(c for b in snakes for c in b)
    best, best_d = 0, None
    for a in (1, 2, 3, 4):
        if vel and a == _OPP[vel]:
            continue
        nx, ny = hx + _DIRS[a][0], hy + _DIRS[a][1]
        if not (0 <= nx < env.dim and 0 <= ny < env.dim) or (nx, ny) in occupied:
            continue
        d = min([abs(nx - fx) + abs(ny - fy) for fx, fy in fruits] or [0])
        if best_d is None or d < best_d:
            best, best_d = a, d
    return best


def make_case(name, rules, S, D, N, T, seed=1234):
    from baselines.bench import Monitor  # the reference's own wrapper
    K = 3
    V = D + 2
    cap = (D * D + 1 + 7) & ~7
    envs, recs = [], []
    for i in range(N):
        rec = RecordingDraws(np.random.RandomState(seed + i))
        env = ref_loader.make_env(rules, S, D, rec)
        envs.append(Monitor(env, None, allow_early_resets=True))
        recs.append(rec)
    arng = np.random.RandomState(seed ^ 0x5EED)
    obs = np.stack([e.reset() for e in envs])
    reset_obs_crc = zlib.crc32(obs.tobytes())
    actions = np.zeros((T, N, S), dtype=np.int8)
    reward = np.zeros((T, N), dtype=np.float32)
    done = np.zeros((T, N), dtype=np.uint8)
    num_snakes = np.zeros((T, N), dtype=np.uint8)
    ep_r = np.zeros((T, N), dtype=np.float32)
    ep_l = np.zeros((T, N), dtype=np.int32)
    obs_crc = np.zeros(T, dtype=np.uint32)
    state_crcs = np.zeros(T, dtype=np.uint32)
    per_env = np.zeros(N, dtype=np.uint32)
    for t in range(T):
        for i, e in enumerate(envs):
            raw = e.env
            if i < N // 2:
                a = arng.randint(0, 5, size=S)
            else:
                a = np.array([scripted_action(raw, s, arng) for s in range(S)])
            actions[t, i] = a
            ob, r, d, info = e.step(a)
            if d:  # subproc_vec_env.py:14-15
                ob = e.reset()
                ep_r[t, i], ep_l[t, i] = info["episode"]["r"], info["episode"]["l"]
            obs[i], reward[t, i], done[t, i], num_snakes[t, i] = ob, r, d, info["num_snakes"]
            per_env[i] = state_crc(canonical_from_reference(raw, S, rules, cap))
        obs_crc[t] = zlib.crc32(obs.tobytes())
        state_crcs[t] = zlib.crc32(per_env.tobytes())
    offsets = np.zeros(N + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(r.vals) for r in recs])
    vals = np.concatenate([np.asarray(r.vals, dtype=np.uint32) for r in recs])
    bounds = np.concatenate([np.asarray(r.bounds, dtype=np.uint32) for r in recs])
    final = [canonical_from_reference(e.env, S, rules, cap) for e in envs]
    out = dict(
        rules=rules, S=S, D=D, F=S, K=K, N=N, T=T, cap=cap,
        actions=actions, tape_vals=vals, tape_bounds=bounds, tape_offsets=offsets,
        reward=reward, done=done, num_snakes=num_snakes, ep_r=ep_r, ep_l=ep_l,
        reset_obs_crc=np.uint32(reset_obs_crc), obs_crc=obs_crc, state_crc=state_crcs,
        final_obs=obs, final_t=np.array([f["t"] for f in final], dtype=np.int32),
        final_spare=np.array([f["spare"] for f in final], dtype=np.uint32),
        final_len=np.stack([f["len"] for f in final]), final_grow_to=np.stack([f["grow_to"] for f in final]),
        final_vel=np.stack([f["vel"] for f in final]), final_body=np.stack([f["body"] for f in final]),
    )
    if rules == "classic":
        out["final_fruit"] = np.stack([f["fruit"] for f in final])
    else:
        out["final_fruit_grid"] = np.stack([f["fruit_grid"] for f in final])
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **out)
    return int(done.sum()), int(len(vals)), float(reward.sum()), int(np.stack([f["len"] for f in final]).max())


# ---------------------------------------------------------------- hand-built known answers
def known_answer(rules, S, D, snakes, fruits, vels, grow, t, action, draws, spare=0, note=""):
    """One reference step from an injected state (SURVEY.md section 8c list)."""
    class Fixed(object):
        def __init__(self, vals):
            self.vals, self.log = list(vals), []

        def randint(self, n):
            v = self.vals.pop(0)
            assert 0 <= v < n, (v, n)
            self.log.append([int(n), int(v)])
            return v

    rng = Fixed(draws)
    env = ref_loader.make_env(rules, S, D, rng)
    pad = lambda lst, fill: list(lst) + [fill] * (3 - len(lst))
    env.state = [[list(map(tuple, b)) for b in snakes], [tuple(f) for f in fruits],
                 pad([tuple(v) for v in vels], (0, 0)), pad(grow, 3), t]
    if rules == "adversarial":
        env.spare_fruits = spare
    ob, r, d, info = env.step(action if S > 1 else action[0])
    sn, fr, ve, gr, tt = env.state
    return {
        "note": note, "rules": rules, "S": S, "D": D,
        "in": {"snakes": snakes, "fruits": fruits, "vels": vels, "grow_to": grow, "t": t, "spare": spare},
        "action": list(action), "draws": rng.log,
        "out": {"snakes": [[[int(c[0]), int(c[1])] for c in b] for b in sn],
                "fruits": [[int(f[0]), int(f[1])] for f in fr],
                "vels": [[int(v[0]), int(v[1])] for v in ve[:S]], "grow_to": [int(g) for g in gr[:S]], "t": int(tt),
                "spare": int(getattr(env, "spare_fruits", 0))},
        "reward": float(r), "done": bool(d), "num_snakes": int(info["num_snakes"]),
        "obs_crc": zlib.crc32(ob.tobytes()),
    }


def known_answers():
    K = []
    add = lambda *a, **k: K.append(known_answer(*a, **k))
    add("classic", 1, 5, [[[2, 2], [1, 2], [0, 2]]], [[4, 4]], [[1, 0]], [3], 5, [3], [],
        note="reversal ignored: keeps moving +x")
    add("classic", 2, 5, [[[1, 2]], [[3, 2]]], [[0, 0], [4, 4]], [[1, 0], [-1, 0]], [3, 3], 3, [0, 0], [],
        note="head-to-head on (2,2): both die, reward -1, done")
    add("classic", 2, 5, [[[1, 2]], [[2, 2]]], [[0, 0], [4, 4]], [[1, 0], [0, 0]], [3, 3], 3, [0, 0], [],
        note="moving head onto a stationary head: both die")
    add("classic", 1, 5, [[[1, 1], [1, 2], [2, 2], [2, 1]]], [[4, 4]], [[0, -1]], [4], 9, [1], [],
        note="own-tail chase legal when len == grow_to (tail popped first)")
    add("classic", 1, 5, [[[1, 1], [1, 2], [2, 2], [2, 1]]], [[4, 4]], [[0, -1]], [6], 9, [1], [],
        note="own-tail chase fatal while growing")
    add("classic", 2, 5, [[[1, 1]], [[2, 1], [3, 1], [4, 1]]], [[0, 4], [4, 4]], [[1, 0], [0, 1]], [3, 3], 2, [0, 0], [],
        note="entering a body cell of another snake that is not vacated this step: dies")
    add("classic", 2, 5, [[[1, 1]], [[3, 1], [3, 2], [2, 2], [2, 1]]], [[0, 4], [4, 4]], [[1, 0], [1, 0]], [3, 4], 2, [0, 0], [],
        note="entering another snake's vacating tail cell is legal")
    add("classic", 2, 5, [[[0, 3]], [[2, 0], [1, 0]]], [[0, 0], [3, 0]], [[-1, 0], [1, 0]], [3, 3], 4, [0, 0], [13],
        note="OOB alias: snake 0 leaves the board at (-1,3) and masks cell (4,2) for snake 1's respawn")
    add("classic", 2, 5, [[[1, 1]], [[4, 4]]], [[2, 1], [2, 1]], [[1, 0], [0, 0]], [3, 3], 0, [0, 0], [5, 6],
        note="two fruits on one cell: reward 2, grow 3 -> 7, two draws")
    add("classic", 2, 5, [[[1, 1]], [[2, 1], [2, 2]]], [[2, 1], [4, 4]], [[1, 0], [0, -1]], [3, 3], 0, [0, 0], [3],
        note="eat and die in the same step: reward -1 but the draw still happens")
    add("classic", 2, 5, [[[1, 1]], [[4, 4]]], [[0, 0], [3, 3]], [[0, 1], [0, 0]], [3, 3], 1999, [0, 0], [],
        note="t = 1999 -> 2000: done with reward 0")
    add("classic", 2, 5, [[[2, 2]], [[2, 2]]], [[0, 0], [3, 3]], [[0, 0], [0, 0]], [3, 3], 0, [0, 0], [],
        note="same-cell spawn: both die on step 1 without moving")
    add("classic", 2, 5, [[[2, 2]], [[2, 2]]], [[0, 0], [3, 3]], [[0, 0], [0, 0]], [3, 3], 0, [1, 2], [],
        note="same-cell spawn, both move away: both live, bodies overlap on (2,2)")
    add("classic", 1, 5, [[[2, 2], [1, 2]]], [[4, 4]], [[1, 0]], [3], 0, [7], [],
        note="action 7 is a no-op")
    add("classic", 1, 2, [[[1, 0], [1, 1], [0, 1]]], [[0, 0]], [[0, -1]], [9], 0, [3], [],
        note="2x2 board filled after eating: fruit -> (0,0) with no draw")
    add("classic", 3, 5, [[[1, 1]], [[3, 3]], []], [[0, 0], [4, 4], [2, 0]], [[1, 0], [0, 1], [0, 0]], [3, 3, 3], 7, [0, 0, 2], [],
        note="dead snake 2 ignores its action")
    add("classic", 2, 5, [[[4, 2]], [[0, 0]]], [[1, 1], [3, 3]], [[1, 0], [0, 0]], [3, 3], 0, [0, 2], [],
        note="wall: snake 0 leaves at x = 5")
    add("adversarial", 2, 5, [[[1, 1]], [[3, 1], [3, 2], [3, 3]]], [[2, 1], [4, 4]], [[1, 0], [0, -1]], [3, 3], 0, [0, 0], [4],
        spare=0, note="adversarial, spare 0: eaten fruit respawns")
    add("adversarial", 2, 5, [[[1, 1]], [[3, 1], [3, 2], [3, 3]]], [[2, 1], [4, 4]], [[1, 0], [0, -1]], [3, 3], 0, [0, 0], [],
        spare=2, note="adversarial, spare 2: fruit stays under the head, spare -> 1, no draw")
    add("adversarial", 2, 5, [[[1, 1]], [[4, 1], [3, 1], [2, 1]]], [[0, 0], [4, 4]], [[0, 1], [1, 0]], [3, 3], 0, [0, 0], [],
        spare=0, note="adversarial death: body incl. the OOB head becomes fruit, spare += len^2")
    return K


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    only = sys.argv[1:]
    ka = known_answers()
    with open(os.path.join(GOLDEN, "known_answers.json"), "w") as f:
        json.dump(ka, f, indent=1)
    print("known answers: %d" % len(ka))
    for case in CASES:
        if only and case[0] not in only:
            continue
        t0 = time.time()
        episodes, draws, ret, maxlen = make_case(*case)
        print("%-18s episodes=%d draws=%d sum_reward=%.0f max_len=%d  (%.0fs)" % (case[0], episodes, draws, ret, maxlen, time.time() - t0))


if __name__ == "__main__":
    main()
