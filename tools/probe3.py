import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, snakes_b200
env = snakes_b200.SnakeVecEnv(131072, size=19, n_snakes=2)
env.reset()
acts = [env.gen_actions(t, 1).clone() for t in range(8)]
for t in range(12):
    env.step(acts[t % 8])
torch.cuda.synchronize()
