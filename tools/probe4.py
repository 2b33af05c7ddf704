import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, snakes_b200
def run(tag):
    env = snakes_b200.SnakeVecEnv(131072, size=19, n_snakes=2)
    env.reset()
    m = torch.zeros(131072, dtype=torch.bool, device="cuda")
    for _ in range(5): env.reset(mask=m)     # masked reset of nothing = logic kernel (trivial) + paint kernel
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): env.reset(mask=m)
    e1.record(); torch.cuda.synchronize()
    print("%-40s %.1f us per (trivial logic + paint)" % (tag, e0.elapsed_time(e1) / 50 * 1e3))
    env.close()
for store in ("tma", "stg"):
    for dbg in ("0", "1", "2", "3"):
        os.environ["SNK_STORE"] = store; os.environ["SNK_DEBUG"] = dbg
        run("store=%s debug=%s (1=no paint, 2=no store)" % (store, dbg))
