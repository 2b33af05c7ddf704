"""Timing of the large-field kernel (16 snakes on 64x64) for one build / env setting."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, snakes_b200
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for rules in sys.argv[2:] or ("classic", "cut"):
    env = snakes_b200.SnakeVecEnv(N, size=64, n_snakes=16, rules=rules); env.reset()
    acts = [env.gen_actions(t, 1).clone() for t in range(8)]
    for t in range(10): env.step(acts[t % 8])
    env.reset_stats()
    T = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(T): env.step_async(acts[t % 8]); env._pending = False
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / T * 1e3
    st = env.stats(False); sl = st["body_cells"] / max(st["env_steps"], 1)
    ab = env.algorithmic_bytes_per_step(sl)
    print("%s N=%d: %.0f us/step  %.3e env-steps/s  frac %.3f  %s" % (rules, N, us, N / us * 1e6, ab * N / us / 1e3 / 6548.2, env.launch_info()), flush=True)
    env.close()
