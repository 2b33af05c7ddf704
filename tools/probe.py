"""Quick throughput probe of the step kernel (CUDA events; not the bench contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snakes_b200

def probe(N, steps=200, warm=50, **kw):
    env = snakes_b200.SnakeVecEnv(N, **kw)
    env.reset()
    acts = [env.gen_actions(t, 1).clone() for t in range(16)]
    for t in range(warm):
        env.step(acts[t % 16])
    torch.cuda.synchronize()
    env.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(steps):
        env.step_async(acts[t % 16]); env._pending = False
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    st = env.stats(False)
    sl = st["body_cells"] / max(st["env_steps"], 1)
    ab = env.algorithmic_bytes_per_step(sl)
    rate = N / (ms * 1e-3)
    print("N=%d %s  %.1f us/step  %.3e env-steps/s  %.3e agent-steps/s  sumL=%.2f  algB=%.0f  %.1f GB/s  frac=%.3f  %s" % (
        N, kw, ms * 1e3, rate, rate * env.S, sl, ab, ab * rate / 1e9, ab * rate / 1e9 / 6548.2, env.launch_info()))
    env.close()

if __name__ == "__main__":
    probe(131072, size=19, n_snakes=2)
    probe(1048576, size=19, n_snakes=2)
    probe(4096, size=10, n_snakes=2)
    probe(65536, size=10, n_snakes=3, rules="cut")
    probe(4096, size=64, n_snakes=16, rules="cut", steps=20, warm=5)
