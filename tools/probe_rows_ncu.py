import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, snakes_b200
rules = sys.argv[1] if len(sys.argv) > 1 else "cut"
env = snakes_b200.SnakeVecEnv(16384, size=64, n_snakes=16, rules=rules); env.reset()
for t in range(14): env.step(env.gen_actions(t, 1))
torch.cuda.synchronize()
