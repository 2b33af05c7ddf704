import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.probe import probe
for variant in ("fused", "ws", "split"):
    os.environ["SNK_LANE"] = variant
    print("variant", variant)
    probe(65536, size=10, n_snakes=3, rules="cut")
    probe(65536, size=10, n_snakes=3, rules="classic")
    probe(4096, size=10, n_snakes=2)
    probe(1048576, size=10, n_snakes=3, rules="cut", steps=30, warm=5)
