import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, snakes_b200
def probe(N, steps=100, warm=20, **kw):
    env = snakes_b200.SnakeVecEnv(N, **kw)
    env.reset()
    acts = [env.gen_actions(t, 1).clone() for t in range(8)]
    for t in range(warm): env.step(acts[t % 8])
    torch.cuda.synchronize(); env.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(steps): env.step_async(acts[t % 8]); env._pending = False
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    st = env.stats(False); sl = st["body_cells"] / max(st["env_steps"], 1)
    ab = env.algorithmic_bytes_per_step(sl); rate = N / (ms * 1e-3)
    print("N=%d %s %.1f us/step %.3e env-steps/s algB=%.0f %.0f GB/s frac=%.3f" % (N, kw, ms * 1e3, rate, ab, ab * rate / 1e9, ab * rate / 1e9 / 6548.2))
    env.close()
probe(131072, size=19, n_snakes=2, obs_mode="atari84")
probe(32768, size=19, n_snakes=2, obs_mode="atari84")
probe(65536, size=10, n_snakes=3, obs_mode="atari84")
