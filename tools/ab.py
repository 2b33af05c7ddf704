"""A/B timing of one library build (SNK_LIB=...): short bodies (random actions), long bodies (scripted policy), optional extra shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, snakes_b200

def timed(env, acts, T):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for t in range(T):
        env.step_async(acts[t % len(acts)]); env._pending = False
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / T * 1e3

def report(tag, env, us, N):
    st = env.stats(False)
    sl = st["body_cells"] / max(st["env_steps"], 1)
    ab = env.algorithmic_bytes_per_step(sl)
    es = max(st["env_steps"], 1)
    print("%-28s %7.1f us  sumL %5.1f  frac %.3f  (per env-step: %.3f fruits, %.3f draws, %.4f episodes)" % (tag, us, sl, ab * N / us / 1e3 / 6548.2, st["fruits"] / es, st["draws"] / es, st["episodes"] / es), flush=True)

def short(N, reps=3, **kw):
    env = snakes_b200.SnakeVecEnv(N, **kw); env.reset()
    acts = [env.gen_actions(t, 1).clone() for t in range(16)]
    timed(env, acts, 300)
    env.reset_stats()
    us = min(timed(env, acts, 500) for _ in range(reps))
    report("short N=%d %s" % (N, kw), env, us, N); env.close()

def long_(N, warm=400, **kw):
    env = snakes_b200.SnakeVecEnv(N, **kw); env.reset()
    for t in range(warm):
        env.step(env.gen_scripted_actions(t, 7))
    T = 100
    acts = torch.empty((T, N, env.S), dtype=torch.int8, device="cuda")
    blob = env.dump_state_blob()
    for t in range(T):
        env.gen_scripted_actions(warm + t, 7, out=acts[t]); env.step_async(acts[t]); env._pending = False
    best = 1e9
    for _ in range(3):
        env.load_state_blob(blob); env.reset_stats()
        best = min(best, timed(env, acts, T))
    report("long  N=%d %s" % (N, kw), env, best, N); env.close()

if __name__ == "__main__":
    print("lib:", os.environ.get("SNK_LIB", "default"))
    which = sys.argv[1:] or ["short", "long"]
    if "short" in which: short(131072, size=19, n_snakes=2)
    if "long" in which: long_(131072, size=19, n_snakes=2)
    if "c3" in which: short(65536, size=10, n_snakes=3, rules="cut")
    if "c3l" in which: long_(65536, size=10, n_snakes=3, rules="cut")
    if "c2" in which: short(4096, size=10, n_snakes=2)
    if "1m" in which: short(1048576, size=19, n_snakes=2, reps=1)
