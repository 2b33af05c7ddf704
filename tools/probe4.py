import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, snakes_b200
os.environ["SNK_LANE"] = "split"
def run(tag, N=131072):
    env = snakes_b200.SnakeVecEnv(N, size=19, n_snakes=2)
    env.reset()
    L = snakes_b200._lib.lib()
    import ctypes as C
    # MODE_OBSERVE is not exposed; a masked reset of nothing = trivial logic kernel + paint kernel
    m = torch.zeros(N, dtype=torch.bool, device="cuda")
    for _ in range(5): env.reset(mask=m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): env.reset(mask=m)
    e1.record(); torch.cuda.synchronize()
    print("%-60s %.1f us" % (tag, e0.elapsed_time(e1) / 50 * 1e3))
    env.close()
for store in ("tma", "stg"):
    for dbg in ("3", "1", "17"):
        os.environ["SNK_STORE"] = store; os.environ["SNK_DEBUG"] = dbg
        run("store=%s debug=%s (3: nothing; 1: store only; 17: store only, no read-wait)" % (store, dbg))
for l2 in ("0", "1"):
    os.environ["SNK_STORE"] = "tma"; os.environ["SNK_DEBUG"] = "17"; os.environ["SNK_L2"] = l2
    run("tma no-wait store-only SNK_L2=%s" % l2)
