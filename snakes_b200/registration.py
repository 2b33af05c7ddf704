"""Env ids and constructors -- the reference's Gym surface for this path.

Mirrors gym_snake/__init__.py:3-26 (the six `register(...)` calls) and
`utils.make_basic_env` (src/utils.py:34-49).  gym is not a dependency: ids live in an internal
registry, and are ALSO registered with gym / gymnasium when one of them is importable, so
`gym.make('snake-multiple-test-v0')` keeps working where it exists.
"""
import numpy as np

from .spaces import Box, Discrete  # noqa: F401
from .vec_env import SnakeVecEnv

# id -> default kwargs.  Defaults follow the reference classes: SnakeEnv dim=19 and 3 emitted views
# (snake_multiple_test.py:14, :93-95), SnakeAdversarial dim=10 (snake_adversarial_env.py:11),
# NewMultipleSnakes size=(10,10), n_snakes=2, n_fruits=4, one view per snake
# (snake_multiple_env_new.py:10, core/new_world.py:206-214).  The three ids whose classes do not
# run at the reference's HEAD are kept as aliases of the classic rules.
ENV_IDS = {
    "snake-multiple-test-v0": dict(size=(19, 19), n_snakes=2, n_views=3, rules="classic"),
    "snake-adversarial-v0": dict(size=(10, 10), n_snakes=2, n_views=3, rules="adversarial"),
    "snake-new-multiple-v0": dict(size=(10, 10), n_snakes=2, n_fruits=4, rules="classic"),
    "snake-multiple-v0": dict(size=(10, 10), n_snakes=2, rules="classic"),
    "snake-single-v0": dict(size=(10, 10), n_snakes=1, rules="classic"),
    "snake-competitive-v0": dict(size=(10, 10), n_snakes=2, rules="classic"),
    "snake-cut-v0": dict(size=(10, 10), n_snakes=3, rules="cut"),  # README.md:11 body-cut action (no reference class)
}


class SnakeGymEnv(object):
    """One env with the per-instance Gym API of the reference (reset / step / seed / render /
    close, `action_space`), as evaluate_snake.py:52-117 uses it.  numpy in, numpy out."""

    metadata = {"render.modes": ["rgb_array"]}
    reward_range = (-1.0, float("inf"))
    spec = None

    def __init__(self, size=(10, 10), n_snakes=2, n_fruits=None, n_views=None, rules="classic", screen_res=300,
                 seed=0, device=0, max_steps=2000, obs_mode="native"):
        # the reference passes kwargs by RE-CALLING __init__ on a made env (utils.py:38): allowed here too
        if getattr(self, "_venv", None) is not None:
            self._venv.close()
        self._kw = dict(size=size, n_snakes=n_snakes, n_fruits=n_fruits, n_views=n_views, rules=rules,
                        screen_res=screen_res, device=device, max_steps=max_steps, obs_mode=obs_mode)
        self._seed = seed
        self._venv = SnakeVecEnv(1, seed=seed, auto_reset=False, **self._kw)
        self.action_space = self._venv.action_space
        self.observation_space = self._venv.observation_space
        self.n_snakes = self._venv.S

    def seed(self, seed=None):
        if seed is None:
            seed = int(np.random.SeedSequence().generate_state(1)[0])
        self._seed = int(seed)
        self._venv.close()
        self._venv = SnakeVecEnv(1, seed=self._seed, auto_reset=False, **self._kw)
        return [self._seed]

    def reset(self):
        return self._venv.reset()[0].cpu().numpy()

    def step(self, action):
        if not hasattr(action, "__len__"):  # snake_multiple_test.py:167-168
            action = [action]
        a = np.zeros((1, self.n_snakes), dtype=np.int8)
        a[0, :len(action)] = np.asarray(action, dtype=np.int64).clip(-128, 127)
        obs, rew, done, infos = self._venv.step(a)
        info = infos[0]
        info["rewards_all"] = self._venv.rewards_all[0].cpu().tolist()
        return obs[0].cpu().numpy(), float(rew[0]), bool(done[0]), info

    def render(self, mode="rgb_array"):
        return self._venv.render(mode, 0)

    def close(self):
        if self._venv is not None:
            self._venv.close()

    @property
    def unwrapped(self):
        return self


_registry = {}


def register(id, **kwargs):
    _registry[id] = dict(kwargs)


def make(id, num_envs=None, **kwargs):
    """gym.make(id): one `SnakeGymEnv`, or a `SnakeVecEnv` of `num_envs` when that is given."""
    if id not in _registry:
        raise KeyError("No registered env with id: %s" % id)
    kw = dict(_registry[id])
    kw.update(kwargs)
    if num_envs is None:
        return SnakeGymEnv(**kw)
    return SnakeVecEnv(num_envs, **kw)


def make_basic_env(env_id, num_env, seed, start_index=0, n_snakes=None, device=0, **kwargs):
    """utils.make_basic_env (src/utils.py:34-49): `num_env` envs, env i seeded by (seed, start_index + i).
    Returns the batched env; Monitor's episode accounting is built in."""
    kw = dict(kwargs)
    if n_snakes is not None:  # Config.NUM_SNAKES of the reference (utils.py:38)
        kw.update(n_snakes=n_snakes, n_fruits=n_snakes)
    return make(env_id, num_envs=num_env, seed=seed, env_id_base=start_index, device=device, **kw)


def _register_all():
    for env_id, kw in ENV_IDS.items():
        register(env_id, **kw)
    for modname in ("gym", "gymnasium"):
        try:
            mod = __import__(modname + ".envs.registration", fromlist=["register"])
            for env_id, kw in ENV_IDS.items():
                try:
                    mod.register(id=env_id, entry_point="snakes_b200.registration:SnakeGymEnv", kwargs=kw)
                except Exception:
                    pass
        except ImportError:
            pass


_register_all()
