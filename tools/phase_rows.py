"""Per-phase cycles of k_step_rows (tools/build_variant.sh phase -DSNK_PHASE_TIMING): producer warp vs consumers."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, snakes_b200
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for rules in ("classic", "cut"):
    env = snakes_b200.SnakeVecEnv(N, size=64, n_snakes=16, rules=rules); env.reset()
    acts = [env.gen_actions(t, 1).clone() for t in range(8)]
    for t in range(10): env.step(acts[t % 8])
    env.reset_stats()
    T = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(T): env.step_async(acts[t % 8]); env._pending = False
    e1.record(); torch.cuda.synchronize()
    st = env.stats(False); info = env.launch_info(); ctas = info["grid"]
    per_env = lambda k: st[k] / (N * T)
    print("%s N=%d: %.0f us/step; per CTA per launch total %.0f cyc; per env: producer logic %.0f, grid build %.0f | consumers %.0f (TMA read-wait %.0f); envs per CTA %.1f" % (
        rules, N, e0.elapsed_time(e1) / T * 1e3, st["length_sum"] / ctas / T, per_env("fruits"), per_env("deaths"), per_env("body_cells"), per_env("draws"), N / ctas))
    env.close()
