import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, snakes_b200
env = snakes_b200.SnakeVecEnv(131072, size=19, n_snakes=2, obs_mode="atari84")
env.reset()
for t in range(12):
    env.step(env.gen_actions(t, 1))
torch.cuda.synchronize()
