"""Import the UNMODIFIED reference environment: from /root/reference (build container) or from the
offline install of it under baseline/_ref/ (oracle/stage_reference.py; git-ignored, travels to the GPU box).

TEST / BENCH INFRASTRUCTURE ONLY.  Used by oracle/make_golden.py, by the CPU tests that are skipped
when no reference is present, and by bench.py's CPU legs (`--impl reference`, `cpu_baseline`), which
time the reference's own SubprocVecEnv path when it is available.  The product never imports it.

Reference entry points loaded:
  src/gym-snake/gym_snake/envs/snake_multiple_test.py:12    SnakeEnv          (canonical rules)
  src/gym-snake/gym_snake/envs/snake_adversarial_env.py:9   SnakeAdversarial  (dead body -> fruit)
  src/config.py:3                                           Config.NUM_SNAKES read in reset()
  src/baselines/common/vec_env/{subproc,dummy}_vec_env.py, src/baselines/bench/monitor.py
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("SNK_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "gym_shim")
STAGED_ROOT = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")   # pip --target install: flat packages


def source():
    """'tree' (the reference checkout), 'staged' (baseline/_ref) or None."""
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "gym-snake", "gym_snake")):
        return "tree"
    if os.path.isdir(os.path.join(STAGED_ROOT, "gym_snake", "envs")):
        return "staged"
    return None


def available():
    return source() is not None


def _install_paths():
    if source() == "tree":
        paths = [_SHIM, os.path.join(REFERENCE_ROOT, "src", "gym-snake"), os.path.join(REFERENCE_ROOT, "src")]
    else:
        paths = [_SHIM, STAGED_ROOT]
    for p in reversed(paths):
        if p not in sys.path:
            sys.path.insert(0, p)


def load():
    """Returns (SnakeEnv, SnakeAdversarial, Config) classes of the reference."""
    if not available():
        raise RuntimeError("reference not present at %s nor staged at %s" % (REFERENCE_ROOT, STAGED_ROOT))
    _install_paths()
    import gym_snake  # noqa: F401  (runs the reference's register() calls against the shim)
    from gym_snake.envs.snake_multiple_test import SnakeEnv
    from gym_snake.envs.snake_adversarial_env import SnakeAdversarial
    from config import Config
    return SnakeEnv, SnakeAdversarial, Config


def make_env(rules, S, D, rng):
    """A reference env instance with dim patched to D, NUM_SNAKES=S and `rng` as np_random.

    rules: 'classic' -> SnakeEnv, 'adversarial' -> SnakeAdversarial.  S must be 1..3: the
    reference hard-codes 3-long grow_to/vels lists (snake_multiple_test.py:227-228).
    """
    SnakeEnv, SnakeAdversarial, Config = load()
    assert 1 <= S <= 3
    Config.set_num_snakes(S)
    env = {"classic": SnakeEnv, "adversarial": SnakeAdversarial}[rules]()
    env.dim = D
    env.np_random = rng
    return env


def _thunk(rules, S, D, seed, rank, atari84):
    """What utils.make_basic_env's make_env(rank) builds (src/utils.py:35-45), minus the
    `env.__init__(n_snakes=...)` line that SnakeEnv.__init__(self) rejects at HEAD:
    SnakeEnv -> seed(seed + rank) -> Monitor(env, None, allow_early_resets=True) [-> WarpFrame]."""
    def make():
        SnakeEnv, SnakeAdversarial, Config = load()
        from baselines.bench import Monitor
        Config.set_num_snakes(S)
        env = {"classic": SnakeEnv, "adversarial": SnakeAdversarial}[rules]()
        env.dim = D
        env.seed(seed + rank)
        env = Monitor(env, None, allow_early_resets=True)
        if atari84:
            import importlib
            env = importlib.import_module("utils").WarpFrame(env)
        return env
    return make


def make_subproc_vec_env(n_procs, rules="classic", S=2, D=19, seed=0, atari84=False):
    """The reference's vectorised CPU path: its own SubprocVecEnv (baselines/common/vec_env/subproc_vec_env.py:31)
    over `n_procs` worker processes, one Monitor(SnakeEnv) each."""
    load()
    from baselines.common.vec_env.subproc_vec_env import SubprocVecEnv
    return SubprocVecEnv([_thunk(rules, S, D, seed, i, atari84) for i in range(n_procs)])
