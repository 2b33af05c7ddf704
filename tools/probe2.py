import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.probe import probe
for variant, lw in (("fused", "3"), ("ws", "3"), ("ws", "2")):
    os.environ["SNK_LANE"] = variant; os.environ["SNK_LOGIC_WARPS"] = lw
    print("variant", variant, "logic warps", lw)
    probe(131072, size=19, n_snakes=2)
    probe(1048576, size=19, n_snakes=2, steps=50, warm=10)
    probe(65536, size=10, n_snakes=3, rules="cut")
