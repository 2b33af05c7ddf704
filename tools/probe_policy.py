"""How long does the scripted-policy kernel take by itself?  200 back-to-back launches on a long-body state (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, snakes_b200
N = 131072
env = snakes_b200.SnakeVecEnv(N, size=19, n_snakes=2); env.reset()
g = env.make_scripted_graph(400, 0, 7); g.launch(); g.close()
torch.cuda.synchronize()
env.step_async(env.gen_scripted_actions(400, 7)); env._pending = False   # eager step: obs_current is known
torch.cuda.synchronize()
for tag, invalidate in (("occupancy from the observations", False), ("occupancy by walking the bodies", True)):
    if invalidate:
        blob = env.dump_state_blob(); env.load_state_blob(blob)
    for _ in range(20): env.gen_scripted_actions(401, 7)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(200): env.gen_scripted_actions(401, 7)
    e1.record(); torch.cuda.synchronize()
    print("policy kernel, %s: %.2f us per launch (200 back to back)" % (tag, e0.elapsed_time(e1) / 200 * 1e3))
