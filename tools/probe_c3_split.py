import sys, os
sys.path.insert(0, os.getcwd())
import torch, snakes_b200
env = snakes_b200.SnakeVecEnv(65536, size=10, n_snakes=3, rules="cut"); env.reset()
acts = [env.gen_actions(t, 1).clone() for t in range(8)]
for t in range(40): env.step(acts[t % 8])
torch.cuda.synchronize()
print(env.launch_info())
