import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, snakes_b200
N = 131072
env = snakes_b200.SnakeVecEnv(N, size=19, n_snakes=2)
env.reset()
for t in range(450):
    env.step(env.gen_scripted_actions(t, 7))
L = env.dump_state()["len"].astype(np.int64)
print("mean len per snake %.1f  p50 %d p90 %d p99 %d max %d" % (L.mean(), np.percentile(L, 50), np.percentile(L, 90), np.percentile(L, 99), L.max()))
per_env = L.sum(1)
w = per_env.reshape(-1, 32)
print("per-warp (32 envs): mean of max single snake %.1f, mean of max sumL %.1f; per image (8 envs) mean of max snake %.1f" % (
    L.reshape(-1, 32, 2).max((1, 2)).mean(), w.max(1).mean(), L.reshape(-1, 8, 2).max((1, 2)).mean()))
