import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.probe import probe
for l2 in ("0", "1", "2", "3"):
    os.environ["SNK_L2"] = l2
    print("SNK_L2", l2, "(bit0 obs evict-first, bit1 records evict-last)")
    probe(131072, size=19, n_snakes=2)
    probe(1048576, size=19, n_snakes=2, steps=50, warm=10)
