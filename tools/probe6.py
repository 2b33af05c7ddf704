import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.probe import probe
probe(65536, size=10, n_snakes=3, rules="cut")
probe(65536, size=10, n_snakes=3, rules="adversarial")
probe(65536, size=10, n_snakes=3, rules="classic")
probe(1048576, size=10, n_snakes=3, rules="cut", steps=50, warm=10)
probe(4096, size=10, n_snakes=2, rules="classic")
