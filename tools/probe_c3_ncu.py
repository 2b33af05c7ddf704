import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, snakes_b200
N = int(os.environ.get("PROBE_N", 262144))
env = snakes_b200.SnakeVecEnv(N, size=10, n_snakes=3, rules=os.environ.get("PROBE_RULES", "cut"))
env.reset()
for t in range(40):
    env.step(env.gen_actions(t, 1))
torch.cuda.synchronize()
