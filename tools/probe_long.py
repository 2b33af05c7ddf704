"""Step-kernel time under the scripted policy (long snakes), policy kernel excluded."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, snakes_b200
N = 131072
env = snakes_b200.SnakeVecEnv(N, size=19, n_snakes=2)
env.reset()
for t in range(400):
    env.step(env.gen_scripted_actions(t, 7))
blob = env.dump_state_blob()
T = 100
acts = torch.empty((T, N, 2), dtype=torch.int8, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for t in range(T):
    env.gen_scripted_actions(400 + t, 7, out=acts[t])
    env.step_async(acts[t]); env._pending = False
e1.record(); torch.cuda.synchronize()
both = e0.elapsed_time(e1) / T * 1e3
env.load_state_blob(blob)
env.reset_stats()
torch.cuda.synchronize(); e0.record()
for t in range(T):
    env.step_async(acts[t]); env._pending = False
e1.record(); torch.cuda.synchronize()
step_only = e0.elapsed_time(e1) / T * 1e3
st = env.stats(False)
sl = st["body_cells"] / st["env_steps"]
ab = env.algorithmic_bytes_per_step(sl)
print("policy+step %.1f us, step only %.1f us (sumL %.1f, algB %.0f -> %.0f GB/s, frac %.3f)" % (both, step_only, sl, ab, ab * N / step_only / 1e3, ab * N / step_only / 1e3 / 6548.2))
