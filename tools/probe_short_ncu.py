import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, snakes_b200
N = int(os.environ.get("PROBE_N", 131072))
env = snakes_b200.SnakeVecEnv(N, size=19, n_snakes=2)
env.reset()
for t in range(60):
    env.step(env.gen_actions(t, 1))
torch.cuda.synchronize()
