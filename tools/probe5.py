import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, snakes_b200
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
env = snakes_b200.SnakeVecEnv(N, size=19, n_snakes=2)
env.reset()
acts = [env.gen_actions(t, 1).clone() for t in range(4)]
for t in range(8):
    env.step(acts[t % 4])
torch.cuda.synchronize()
