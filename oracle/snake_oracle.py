"""CPU oracle for the batched multi-snake env step -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may
import this file.  The product package (snakes_b200) never does: it drives the CUDA library
and fails loudly when that library is missing.

This is a from-scratch restatement (own data layout: padded integer cell ids, count grids)
of the reference algorithm, generalised over (D, S, F, K, rules), which the reference
hard-codes.  What each piece follows, relative to /root/reference/src/gym-snake/gym_snake:

  reset draw order           envs/snake_multiple_test.py:219-232, :199-200
  turn / advance / eat / pop envs/snake_multiple_test.py:97-145
  k-th free cell + OOB alias envs/snake_multiple_test.py:202-217
  death test (simultaneous)  envs/snake_multiple_test.py:147-164, :178-185
  reward / t / done / info   envs/snake_multiple_test.py:187-197
  per-view RGB encoding      envs/snake_multiple_test.py:24-58, :93-95 (K=3 hard-coded)
  K = S views                core/new_world.py:206-214
  ctor kwargs                envs/snake_multiple_env_new.py:10-21
  adversarial deltas         envs/snake_adversarial_env.py:14, :137-141, :180-186
  auto-reset + stacking      ../baselines/common/vec_env/subproc_vec_env.py:13-16, :57-61
  episode return / length    ../baselines/bench/monitor.py:57-78

Parity pin: the reference publishes no tests or golden vectors for this path (SURVEY.md
section 4), so the pin is the reference itself, run in the build container through
oracle/ref_loader.py: oracle/make_golden.py records its trajectories into tests/golden/ and
tests/test_oracle_vs_reference.py replays them (and, when /root/reference is present,
cross-checks live).  The `cut` rule-set has no reference code at all (README.md:11 is the
only mention) -- its parity is UNPINNED and is checked kernel-vs-this-oracle only.

Geometry.  D = board edge, V = D + 2.  A cell (x, y) -- x is the FIRST array index of the
observation, as in the reference's ob[x+1][y+1] -- is held as the padded id
pid = (x + 1) * V + (y + 1), which also represents a head that is one step out of bounds.
"""
import zlib

import numpy as np

RULES = ("classic", "adversarial", "cut")
MAX_STEPS = 2000  # snake_multiple_test.py:195
OPPOSITE = (0, 3, 4, 1, 2)  # velocity codes equal the action numbers 1..4; 0 = not moving yet

COL_FRUIT = (255, 0, 0)
COL_WALL = (255, 255, 255)
COL_SELF_BODY, COL_SELF_HEAD = (0, 204, 0), (191, 242, 191)
COL_OTHER_BODY, COL_OTHER_HEAD = (0, 51, 204), (128, 154, 230)


# ----------------------------------------------------------------------------- draws
class TapeUnderrun(RuntimeError):
    pass


class TapeDraws(object):
    """Replays recorded reference draws; checks the requested bound against the recorded one."""

    def __init__(self, vals, bounds=None):
        self.vals = np.asarray(vals, dtype=np.uint32)
        self.bounds = None if bounds is None else np.asarray(bounds, dtype=np.uint32)
        self.ctr = 0

    def randint(self, n):
        if self.ctr >= len(self.vals):
            raise TapeUnderrun("draw %d requested, tape has %d" % (self.ctr, len(self.vals)))
        if self.bounds is not None and int(self.bounds[self.ctr]) != n:
            raise ValueError("draw %d: bound %d requested, tape recorded %d" % (self.ctr, n, int(self.bounds[self.ctr])))
        v = int(self.vals[self.ctr])
        self.ctr += 1
        return v


_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
_U32 = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    """Philox4x32-10 (Salmon et al., SC'11), the production counter-based generator."""
    c0, c1, c2, c3 = [int(c) & _U32 for c in ctr]
    k0, k1 = [int(k) & _U32 for k in key]
    for r in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _U32, p1 & _U32, ((p0 >> 32) ^ c3 ^ k1) & _U32, p0 & _U32
        k0 = (k0 + _W0) & _U32
        k1 = (k1 + _W1) & _U32
    return c0, c1, c2, c3


STREAM_ENV, STREAM_ACTION = 0, 1


def philox_bounded(seed, env_id, stream, index, n):
    """Draw `index` of (seed, env_id, stream) mapped to [0, n) by hi32(u32 * n).  One Philox block
    (counter = index >> 2) serves four consecutive draws (word = index & 3)."""
    blk = index >> 2
    x = philox4x32_10((blk & _U32, (blk >> 32) & _U32, stream, (seed >> 32) & _U32), (seed & _U32, env_id & _U32))[index & 3]
    return (x * n) >> 32


class PhiloxDraws(object):
    """Production RNG: key (seed, env_id), counter = per-env draw index.  `randint(n)` has the
    reference's np_random signature so the REFERENCE env can be driven by it too."""

    def __init__(self, seed, env_id, ctr=0):
        self.seed, self.env_id, self.ctr = int(seed), int(env_id), int(ctr)

    def randint(self, n):
        v = philox_bounded(self.seed, self.env_id, STREAM_ENV, self.ctr, int(n))
        self.ctr += 1
        return v


def philox_actions(seed, env_ids, step, S, n_actions=5):
    """Synthetic uniform action stream used by bench/tests: counter = step * S + snake."""
    env_ids = np.asarray(env_ids)
    out = np.empty((len(env_ids), S), dtype=np.int8)
    for i, e in enumerate(env_ids):
        for s in range(S):
            out[i, s] = philox_bounded(seed, int(e), STREAM_ACTION, step * S + s, n_actions)
    return out


class RecordingDraws(object):
    """np_random proxy for the REFERENCE env: forwards randint(n) and records (n, value)."""

    def __init__(self, inner):
        self.inner = inner
        self.bounds = []
        self.vals = []

    def randint(self, n):
        v = int(self.inner.randint(n))
        self.bounds.append(int(n))
        self.vals.append(v)
        return v


# ----------------------------------------------------------------------------- one env
class SnakeOracle(object):
    """One environment instance with the reference's Gym call signatures."""

    def __init__(self, size=10, n_snakes=2, n_fruits=None, n_views=None, rules="classic",
                 max_steps=MAX_STEPS, draws=None):
        if rules not in RULES:
            raise ValueError("rules must be one of %r" % (RULES,))
        self.D = int(size[0] if hasattr(size, "__len__") else size)
        self.V = self.D + 2
        self.S = int(n_snakes)
        self.F = self.S if n_fruits is None else int(n_fruits)
        self.K = self.S if n_views is None else int(n_views)
        self.rules = rules
        self.max_steps = int(max_steps)
        self.draws = draws
        self.n_actions = 6 if rules == "cut" else 5
        self.delta = (0, self.V, 1, -self.V, -1)
        self.spare_fruits = 0  # adversarial only; survives reset() like the reference's attribute
        self.use_grid = rules != "classic"
        self.t = 0
        self.body = [[] for _ in range(self.S)]
        self.vel = [0] * self.S
        self.grow_to = [3] * self.S
        self.fruit = []
        self.fruit_grid = np.zeros(self.V * self.V, dtype=np.int64)

    # -- geometry helpers
    def pid(self, x, y):
        return (x + 1) * self.V + (y + 1)

    def xy(self, pid):
        return pid // self.V - 1, pid % self.V - 1

    def in_bounds(self, pid):
        x, y = self.xy(pid)
        return 0 <= x < self.D and 0 <= y < self.D

    # -- reset: snake_i.x, snake_i.y, fruit_i.x, fruit_i.y interleaved, no overlap checks
    def reset(self):
        D = self.D
        self.body = [[] for _ in range(self.S)]
        self.fruit = []
        self.fruit_grid[:] = 0
        for i in range(max(self.S, self.F)):
            if i < self.S:
                x = self.draws.randint(D)
                y = self.draws.randint(D)
                self.body[i] = [self.pid(x, y)]
            if i < self.F:
                x = self.draws.randint(D)
                y = self.draws.randint(D)
                self._add_fruit(self.pid(x, y))
        self.vel = [0] * self.S
        self.grow_to = [3] * self.S
        self.t = 0
        return self.observation()

    def _add_fruit(self, pid):
        if self.use_grid:
            self.fruit_grid[pid] += 1
        else:
            self.fruit.append(pid)

    # -- k-th free cell in y-major order; bodies are NOT bounds-checked (aliasing reproduced)
    def _spawn_cell(self):
        D = self.D
        used = np.zeros(D * D, dtype=bool)
        for b in self.body:
            for p in b:
                x, y = self.xy(p)
                idx = y * D + x
                if 0 <= idx < D * D:
                    used[idx] = True
        free = np.flatnonzero(~used)
        if len(free) == 0:
            return self.pid(0, 0)  # no draw is consumed
        idx = int(free[self.draws.randint(len(free))])
        return self.pid(idx % D, idx // D)

    def _advance(self, s, action, strike):
        body = self.body[s]
        if not body:
            return 0
        v = self.vel[s]
        if 1 <= action <= 4 and v != OPPOSITE[action]:
            v = action
        if self.rules == "cut" and action == 5:
            strike[s] = True
        if v == 0:
            return 0
        head = body[0] + self.delta[v]
        if self.use_grid:
            n_eat = int(self.fruit_grid[head])
            hit = []
        else:
            hit = [i for i, f in enumerate(self.fruit) if f == head]
            n_eat = len(hit)
        grow = self.grow_to[s] + 2 * n_eat
        if len(body) >= grow:
            body.pop()
        body.insert(0, head)
        if self.rules == "classic":
            for i in hit:
                self.fruit[i] = self._spawn_cell()
        else:
            for _ in range(n_eat):
                if self.rules == "adversarial" and self.spare_fruits > 0:
                    self.spare_fruits -= 1  # the fruit stays where it is
                else:
                    self.fruit_grid[head] -= 1
                    self.fruit_grid[self._spawn_cell()] += 1
        self.vel[s] = v
        self.grow_to[s] = grow
        self.moved[s] = True
        return n_eat

    def step(self, action):
        S = self.S
        if not hasattr(action, "__len__"):
            action = [action]
        was_alive = [len(b) > 0 for b in self.body]
        strike = [False] * S
        self.moved = [False] * S
        eaten = [self._advance(s, int(action[s]), strike) for s in range(S)]

        # death test on the post-move bodies of ALL snakes, before anything is cleared
        dead = [False] * S
        saved_heads = set()
        cut_at = [None] * S
        for i in range(S):
            if not self.body[i]:
                dead[i] = True
                continue
            head = self.body[i][0]
            contacts = [(j, k) for j in range(S) for k, p in enumerate(self.body[j])
                        if p == head and not (j == i and k == 0)]
            if not self.in_bounds(head):
                dead[i] = True
            elif contacts:
                if (self.rules == "cut" and strike[i] and self.moved[i]
                        and all(j != i and k >= 1 for j, k in contacts)):
                    saved_heads.add(head)
                    for j, k in contacts:
                        cut_at[j] = k if cut_at[j] is None else min(cut_at[j], k)
                else:
                    dead[i] = True
        if self.rules == "cut":
            for j in range(S):
                if cut_at[j] is not None:
                    for p in self.body[j][cut_at[j]:]:
                        if p not in saved_heads:
                            self.fruit_grid[p] += 1
                    del self.body[j][cut_at[j]:]
                    self.grow_to[j] = cut_at[j]
        if self.rules == "adversarial":
            for i in range(S):
                if dead[i]:
                    n = len(self.body[i])
                    for p in self.body[i]:
                        self.fruit_grid[p] += 1
                        self.spare_fruits += n
        for i in range(S):
            if dead[i]:
                self.body[i] = []

        main_dead = not self.body[0]
        reward = -1.0 if main_dead else float(eaten[0])
        rewards_all = [reward] + [(-1.0 if was_alive[s] else 0.0) if dead[s] else float(eaten[s]) for s in range(1, S)]
        self.t += 1
        done = self.t >= self.max_steps or main_dead
        info = {"ale.lives": 1, "num_snakes": S - sum(dead), "rewards_all": rewards_all}
        return self.observation(), reward, done, info

    # -- [V, V, 3K] uint8: fruits, then snakes in index order (body then head), then the border
    def observation(self):
        V, K = self.V, self.K
        ob = np.zeros((K, V * V, 3), dtype=np.uint8)
        if self.use_grid:
            ob[:, np.flatnonzero(self.fruit_grid > 0)] = COL_FRUIT
        else:
            for f in self.fruit:
                ob[:, f] = COL_FRUIT
        for k in range(K):
            for s, b in enumerate(self.body):
                if b:
                    ob[k, b] = COL_SELF_BODY if s == k else COL_OTHER_BODY
                    ob[k, b[0]] = COL_SELF_HEAD if s == k else COL_OTHER_HEAD
        ob = ob.reshape(K, V, V, 3)
        ob[:, 0, :] = COL_WALL
        ob[:, V - 1, :] = COL_WALL
        ob[:, :, 0] = COL_WALL
        ob[:, :, V - 1] = COL_WALL
        return np.ascontiguousarray(ob.transpose(1, 2, 0, 3)).reshape(V, V, 3 * K)

    # -- get_ob_world (snake_multiple_test.py:60-91): one [V, V, 3] image, a colour pair per snake index
    WORLD_COLOURS = (((0, 204, 0), (191, 242, 191)), ((0, 51, 204), (128, 154, 230)),
                     ((204, 0, 119), (230, 128, 188)), ((119, 0, 204), (188, 128, 230)))

    def world_view(self):
        V = self.V
        ob = np.zeros((V * V, 3), dtype=np.uint8)
        if self.use_grid:
            ob[np.flatnonzero(self.fruit_grid > 0)] = COL_FRUIT
        else:
            for f in self.fruit:
                ob[f] = COL_FRUIT
        for s, b in enumerate(self.body):
            if b:
                body_c, head_c = self.WORLD_COLOURS[s]  # the reference has 4 colour pairs (KeyError beyond)
                ob[b] = body_c
                ob[b[0]] = head_c
        ob = ob.reshape(V, V, 3)
        ob[0, :] = ob[V - 1, :] = COL_WALL
        ob[:, 0] = ob[:, V - 1] = COL_WALL
        return ob

    # -- canonical arrays (same layout as snk_dump_state, include/snk.h)
    def canonical(self, cap):
        S = self.S
        length = np.array([len(b) for b in self.body], dtype=np.uint16)
        body = np.zeros((S, cap), dtype=np.uint16)
        for s, b in enumerate(self.body):
            body[s, :len(b)] = b
        out = {
            "t": np.int32(self.t), "spare": np.uint32(self.spare_fruits & _U32),
            "len": length, "grow_to": np.array(self.grow_to, dtype=np.uint16),
            "vel": np.array(self.vel, dtype=np.uint8), "body": body,
        }
        if self.use_grid:
            out["fruit_grid"] = np.minimum(self.fruit_grid, 255).astype(np.uint8)
        else:
            out["fruit"] = np.array(self.fruit, dtype=np.uint16)
        return out


def state_crc(c):
    """crc32 of one env's canonical arrays; the per-step fingerprint stored in tests/golden/."""
    h = zlib.crc32(np.int32(c["t"]).tobytes())
    h = zlib.crc32(np.uint32(c["spare"]).tobytes(), h)
    for key in ("len", "grow_to", "vel"):
        h = zlib.crc32(np.ascontiguousarray(c[key]).tobytes(), h)
    for s in range(len(c["len"])):
        h = zlib.crc32(np.ascontiguousarray(c["body"][s, :int(c["len"][s])]).tobytes(), h)
    if "fruit_grid" in c:
        h = zlib.crc32(np.ascontiguousarray(c["fruit_grid"]).tobytes(), h)
    else:
        h = zlib.crc32(np.ascontiguousarray(c["fruit"]).tobytes(), h)
    return h & _U32


# ----------------------------------------------------------------------------- batch
class VecOracle(object):
    """N instances with SubprocVecEnv's auto-reset contract and Monitor's episode stats."""

    def __init__(self, num_envs, size=10, n_snakes=2, n_fruits=None, n_views=None, rules="classic",
                 max_steps=MAX_STEPS, draws=None, seed=0, env_id_base=0, auto_reset=True):
        self.N = int(num_envs)
        if draws is None:
            draws = [PhiloxDraws(seed, env_id_base + i) for i in range(self.N)]
        self.envs = [SnakeOracle(size, n_snakes, n_fruits, n_views, rules, max_steps, draws[i]) for i in range(self.N)]
        e = self.envs[0]
        self.D, self.V, self.S, self.F, self.K = e.D, e.V, e.S, e.F, e.K
        self.auto_reset = auto_reset
        self.ep_ret = np.zeros(self.N, dtype=np.float64)
        self.ep_len = np.zeros(self.N, dtype=np.int64)

    def reset(self):
        self.ep_ret[:] = 0
        self.ep_len[:] = 0
        return np.stack([e.reset() for e in self.envs])

    def step(self, actions):
        N, S = self.N, self.S
        obs = np.empty((N, self.V, self.V, 3 * self.K), dtype=np.uint8)
        rew = np.empty(N, dtype=np.float32)
        rew_all = np.empty((N, S), dtype=np.float32)
        done = np.empty(N, dtype=bool)
        alive = np.empty(N, dtype=np.uint8)
        fin_ret = np.zeros(N, dtype=np.float32)
        fin_len = np.zeros(N, dtype=np.int32)
        for i, e in enumerate(self.envs):
            ob, r, d, info = e.step(actions[i])
            self.ep_ret[i] += r
            self.ep_len[i] += 1
            if d:
                fin_ret[i], fin_len[i] = self.ep_ret[i], self.ep_len[i]
                if self.auto_reset:
                    ob = e.reset()
                    self.ep_ret[i] = 0
                    self.ep_len[i] = 0
            obs[i], rew[i], done[i], alive[i] = ob, r, d, info["num_snakes"]
            rew_all[i] = info["rewards_all"]
        return obs, rew, done, {"num_snakes": alive, "rewards_all": rew_all,
                                "episode_r": fin_ret, "episode_l": fin_len}

    def state_crcs(self, cap=None):
        cap = cap or self.D * self.D + 1
        return np.array([state_crc(e.canonical(cap)) for e in self.envs], dtype=np.uint32)
