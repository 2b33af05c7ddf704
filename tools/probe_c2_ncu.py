import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, snakes_b200
env = snakes_b200.SnakeVecEnv(4096, size=10, n_snakes=2)
env.reset()
print(env.launch_info())
for t in range(40):
    env.step(env.gen_actions(t, 1))
torch.cuda.synchronize()
