import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, snakes_b200
N = 131072
env = snakes_b200.SnakeVecEnv(N, size=19, n_snakes=2)
env.reset()
for t in range(400):
    env.step(env.gen_scripted_actions(t, 7))
for t in range(6):
    env.step(env.gen_scripted_actions(400 + t, 7))
torch.cuda.synchronize()
