"""Self-play PPO demo on the device: python tools/train_demo.py [n_snakes] [size] [num_envs] [nsteps] [updates]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snakes_b200
from snakes_b200 import selfplay

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1
D = int(sys.argv[2]) if len(sys.argv) > 2 else 10
N = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
T = int(sys.argv[4]) if len(sys.argv) > 4 else 32
U = int(sys.argv[5]) if len(sys.argv) > 5 else 40
torch.manual_seed(0)
torch.backends.cudnn.allow_tf32 = True; torch.backends.cuda.matmul.allow_tf32 = True
env = snakes_b200.SnakeVecEnv(N, size=D, n_snakes=S, seed=0)
t0 = time.time()
model, log = selfplay.learn(env, nsteps=T, total_timesteps=N * T * U, log_interval=max(U // 10, 1), echo=True,
                            opponent_save_interval=5)
torch.cuda.synchronize()
dt = time.time() - t0
print("%d updates, %d agent-steps in %.1f s = %.3e agent-steps/s (learner included)" % (U, N * T * U * S, dt, N * T * U * S / dt))
