import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.probe import probe
for kb, blk in ((12, 128), (12, 64), (12, 256)):
    os.environ["SNK_ROWS_KB"] = str(kb); os.environ["SNK_ROWS_BLOCK"] = str(blk)
    print("rows KB", kb, "block", blk)
    probe(16384, size=64, n_snakes=16, rules="classic", steps=10, warm=3)
    probe(16384, size=64, n_snakes=16, rules="cut", steps=10, warm=3)
