import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.probe import probe
for variant in ("fused", "ws", "split"):
    for mode in ("tma", "stg"):
        os.environ["SNK_STORE"] = mode
        os.environ["SNK_LANE"] = variant
        print("variant", variant, "store mode", mode)
        probe(131072, size=19, n_snakes=2)
        probe(1048576, size=19, n_snakes=2, steps=50, warm=10)
