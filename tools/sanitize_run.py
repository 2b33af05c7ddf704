"""Small run over every kernel family / rule-set for compute-sanitizer (memcheck, racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, snakes_b200
def run(N, steps, force=None, lane=None, **kw):
    for k, v in (("SNK_FORCE_KERNEL", force), ("SNK_LANE", lane)):
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = v
    env = snakes_b200.SnakeVecEnv(N, **kw)
    env.reset()
    for t in range(steps):
        env.step(env.gen_actions(t, 3))
    env.dump_state()
    torch.cuda.synchronize()
    env.check_errors()
    print("ok", env.launch_info()["kernel"], force, lane, kw)
    env.close()
run(333, 40, size=19, n_snakes=2)
run(333, 40, lane="ws", size=19, n_snakes=2)
run(333, 40, lane="split", size=19, n_snakes=2)
run(200, 60, size=10, n_snakes=3, rules="cut")
run(200, 60, size=10, n_snakes=3, rules="adversarial", n_views=3)
run(100, 40, force="tile", size=10, n_snakes=3, rules="adversarial")
run(100, 40, force="dense", size=10, n_snakes=3, rules="cut")
run(40, 30, size=64, n_snakes=16, rules="cut")
run(100, 30, size=19, n_snakes=2, obs_mode="atari84")
run(60, 30, size=12, n_snakes=5, rules="classic")
