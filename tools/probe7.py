import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.probe import probe
probe(16384, size=64, n_snakes=16, rules="classic", steps=10, warm=3)
probe(16384, size=64, n_snakes=16, rules="cut", steps=10, warm=3)
probe(16384, size=64, n_snakes=16, rules="adversarial", steps=10, warm=3)
