"""ctypes binding of oracle/libsnake_oracle.so (the C CPU oracle).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
RULES = {"classic": 0, "adversarial": 1, "cut": 2}
NSTATS = 8


class Config(C.Structure):
    """Mirror of snk_config (include/snk.h)."""
    _fields_ = [("size", C.c_int32), ("n_snakes", C.c_int32), ("n_fruits", C.c_int32), ("n_views", C.c_int32),
                ("rules", C.c_int32), ("max_steps", C.c_int32), ("auto_reset", C.c_int32), ("obs_mode", C.c_int32),
                ("device", C.c_int32), ("rng_mode", C.c_int32), ("num_envs", C.c_int64), ("env_id_base", C.c_int64),
                ("seed", C.c_uint64)]


class Layout(C.Structure):
    """Mirror of snk_state_layout (include/snk.h)."""
    _fields_ = [(n, C.c_size_t) for n in ("total_bytes", "off_t", "off_spare", "off_draw_ctr", "off_ep_ret",
                                          "off_ep_len", "off_len", "off_grow_to", "off_vel", "off_body", "off_fruit")]
    _fields_ += [("cap", C.c_int32), ("fruit_is_grid", C.c_int32)]


def build(force=False):
    so = os.path.join(_HERE, "libsnake_oracle.so")
    src = os.path.join(_HERE, "snake_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libsnake_oracle.so"])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.so_create.restype = C.c_void_p
        L.so_create.argtypes = [C.POINTER(Config)]
        L.so_destroy.argtypes = [C.c_void_p]
        L.so_check_errors.restype = C.c_uint32
        L.so_check_errors.argtypes = [C.c_void_p]
        for name in ("so_set_draw_tape", "so_dump_state", "so_load_state", "so_get_stats", "so_reset_stats",
                     "so_reset", "so_step", "so_step_range", "so_observe", "so_state_layout_of"):
            getattr(L, name).restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_config(num_envs, size=10, n_snakes=2, n_fruits=None, n_views=None, rules="classic", max_steps=2000,
                auto_reset=True, seed=0, env_id_base=0, rng_mode=0, obs_mode=0, device=0):
    return Config(int(size), int(n_snakes), int(n_snakes if n_fruits is None else n_fruits),
                  int(n_snakes if n_views is None else n_views), RULES[rules] if isinstance(rules, str) else int(rules),
                  int(max_steps), int(bool(auto_reset)), int(obs_mode), int(device), int(rng_mode), int(num_envs),
                  int(env_id_base), int(seed))


def split_state(blob, lay, cfg):
    """Views of a canonical state blob (snk_dump_state / so_dump_state) as named numpy arrays."""
    N, S, F, V = cfg.num_envs, cfg.n_snakes, cfg.n_fruits, cfg.size + 2
    b = np.frombuffer(blob, dtype=np.uint8)

    def arr(off, dtype, shape):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        return b[off:off + n].view(dtype).reshape(shape)

    out = {"t": arr(lay.off_t, np.int32, (N,)), "spare": arr(lay.off_spare, np.uint32, (N,)),
           "draw_ctr": arr(lay.off_draw_ctr, np.uint32, (N,)), "ep_ret": arr(lay.off_ep_ret, np.float32, (N,)),
           "ep_len": arr(lay.off_ep_len, np.int32, (N,)), "len": arr(lay.off_len, np.uint16, (N, S)),
           "grow_to": arr(lay.off_grow_to, np.uint16, (N, S)), "vel": arr(lay.off_vel, np.uint8, (N, S)),
           "body": arr(lay.off_body, np.uint16, (N, S, lay.cap))}
    if lay.fruit_is_grid:
        out["fruit_grid"] = arr(lay.off_fruit, np.uint8, (N, V * V))
    else:
        out["fruit"] = arr(lay.off_fruit, np.uint16, (N, F))
    return out


class COracle(object):
    """N envs stepped by the C oracle; same call shape as snake_oracle.VecOracle."""

    def __init__(self, num_envs, **kw):
        self.cfg = make_config(num_envs, **kw)
        self.L = lib()
        self.h = self.L.so_create(C.byref(self.cfg))
        if not self.h:
            raise ValueError("so_create rejected the configuration")
        self.lay = Layout()
        self.L.so_state_layout_of(C.byref(self.cfg), C.byref(self.lay))
        c = self.cfg
        self.N, self.S, self.F, self.K, self.D, self.V = c.num_envs, c.n_snakes, c.n_fruits, c.n_views, c.size, c.size + 2
        self.obs = np.zeros((self.N, self.V, self.V, 3 * self.K), dtype=np.uint8)
        self.reward = np.zeros(self.N, dtype=np.float32)
        self.reward_all = np.zeros((self.N, self.S), dtype=np.float32)
        self.done = np.zeros(self.N, dtype=np.uint8)
        self.num_alive = np.zeros(self.N, dtype=np.uint8)
        self.fin_ret = np.zeros(self.N, dtype=np.float32)
        self.fin_len = np.zeros(self.N, dtype=np.int32)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.so_destroy(C.c_void_p(self.h))
            self.h = None

    def set_draw_tape(self, vals, bounds, offsets):
        vals = np.ascontiguousarray(vals, dtype=np.uint32)
        bounds = None if bounds is None else np.ascontiguousarray(bounds, dtype=np.uint32)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        assert len(offsets) == self.N + 1
        rc = self.L.so_set_draw_tape(C.c_void_p(self.h), _p(vals), _p(bounds), _p(offsets))
        assert rc == 0, rc

    def reset(self, mask=None, want_obs=True):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        rc = self.L.so_reset(C.c_void_p(self.h), _p(m), _p(self.obs) if want_obs else None)
        assert rc == 0, rc
        return self.obs

    def step(self, actions, want_obs=True):
        a = np.ascontiguousarray(actions, dtype=np.int8).reshape(self.N, self.S)
        rc = self.L.so_step(C.c_void_p(self.h), _p(a), _p(self.obs) if want_obs else None, _p(self.reward),
                            _p(self.reward_all), _p(self.done), _p(self.num_alive), _p(self.fin_ret), _p(self.fin_len))
        assert rc == 0, rc
        return self.obs, self.reward, self.done.astype(bool), {"num_snakes": self.num_alive, "rewards_all": self.reward_all,
                                                              "episode_r": self.fin_ret, "episode_l": self.fin_len}

    def observe(self):
        self.L.so_observe(C.c_void_p(self.h), _p(self.obs))
        return self.obs

    def dump_state(self):
        blob = np.zeros(self.lay.total_bytes, dtype=np.uint8)
        rc = self.L.so_dump_state(C.c_void_p(self.h), _p(blob), C.c_size_t(blob.nbytes))
        assert rc == 0, rc
        return blob

    def load_state(self, blob):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        rc = self.L.so_load_state(C.c_void_p(self.h), _p(blob), C.c_size_t(blob.nbytes))
        assert rc == 0, rc

    def state(self):
        return split_state(self.dump_state(), self.lay, self.cfg)

    def stats(self):
        out = np.zeros(NSTATS, dtype=np.float64)
        self.L.so_get_stats(C.c_void_p(self.h), _p(out))
        return out

    def errors(self):
        return int(self.L.so_check_errors(C.c_void_p(self.h)))


def philox(ctr, key):
    out = np.zeros(4, dtype=np.uint32)
    lib().so_philox(_p(np.asarray(ctr, dtype=np.uint32)), _p(np.asarray(key, dtype=np.uint32)), _p(out))
    return tuple(int(x) for x in out)


def gen_actions(cfg, step, seed, n_actions=5):
    a = np.zeros((cfg.num_envs, cfg.n_snakes), dtype=np.int8)
    lib().so_gen_actions(C.byref(cfg), _p(a), C.c_uint64(step), C.c_uint64(seed), C.c_int32(n_actions))
    return a


class BlockOracles(object):
    """C oracles for a few contiguous blocks of a large batch (parity at full size without stepping every env on the
    CPU): block b covers envs [first_b, first_b + count_b) of a batch whose env 0 has global id `env_id_base`.  RNG and
    the synthetic action stream are keyed by the GLOBAL env id, so a block oracle started from reset follows exactly the
    trajectories of those envs inside the big batch.  Blocks step in parallel threads (ctypes drops the GIL)."""

    def __init__(self, blocks, env_id_base=0, **kw):
        self.blocks = [(int(f), int(c), COracle(int(c), env_id_base=env_id_base + int(f), **kw)) for f, c in blocks]
        from concurrent.futures import ThreadPoolExecutor
        self._pool = ThreadPoolExecutor(max_workers=min(len(self.blocks), os.cpu_count() or 1))

    @staticmethod
    def spread(n_envs, n_blocks, count):
        """n_blocks blocks of `count` envs: the first, the last (incl. a ragged tail) and evenly spaced ones."""
        count = min(count, n_envs)
        if n_blocks <= 1 or count >= n_envs:
            return [(0, count)]
        firsts = sorted({int(round(i * (n_envs - count) / float(n_blocks - 1))) for i in range(n_blocks)})
        return [(f, count) for f in firsts]

    def reset(self):
        return list(self._pool.map(lambda b: b[2].reset().copy(), self.blocks))

    def step_generated(self, step, seed, n_actions, want_obs=True):
        """Every block plays the synthetic action stream (so_gen_actions, keyed by global env id)."""
        def one(b):
            co = b[2]
            a = gen_actions(co.cfg, step, seed, n_actions)
            return co.step(a, want_obs=want_obs)
        return list(self._pool.map(one, self.blocks))

    def step(self, actions, want_obs=True):
        """`actions`: int8 [N_big, S] of the whole batch; each block takes its rows."""
        return list(self._pool.map(lambda b: b[2].step(actions[b[0]:b[0] + b[1]], want_obs=want_obs), self.blocks))
