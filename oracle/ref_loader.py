"""Import the UNMODIFIED reference environment from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so nothing that
runs there (pytest -m gpu, smoke(), bench.py) may import this module; it is used by
oracle/make_golden.py and by the CPU tests that are skipped when the tree is absent.

Reference entry points loaded:
  src/gym-snake/gym_snake/envs/snake_multiple_test.py:12    SnakeEnv          (canonical rules)
  src/gym-snake/gym_snake/envs/snake_adversarial_env.py:9   SnakeAdversarial  (dead body -> fruit)
  src/config.py:3                                           Config.NUM_SNAKES read in reset()
  src/baselines/common/vec_env/{subproc,dummy}_vec_env.py, src/baselines/bench/monitor.py
"""
import os
import sys

REFERENCE_ROOT = os.environ.get("SNK_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gym_shim")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "gym-snake", "gym_snake"))


def _install_paths():
    paths = [_SHIM, os.path.join(REFERENCE_ROOT, "src", "gym-snake"), os.path.join(REFERENCE_ROOT, "src")]
    for p in reversed(paths):
        if p not in sys.path:
            sys.path.insert(0, p)


def load():
    """Returns (SnakeEnv, SnakeAdversarial, Config) classes of the reference."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
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
