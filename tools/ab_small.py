"""Small-image configurations under the three forms of the lane path (SNK_DEBUG=lane=...): which one should be the default."""
import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch, snakes_b200
    from tools.ab import short
    short(65536, size=10, n_snakes=3, rules="cut")
    short(65536, size=10, n_snakes=3)
    short(65536, size=10, n_snakes=2)
    short(131072, size=10, n_snakes=1)
    short(65536, size=10, n_snakes=3, rules="adversarial")
    short(1048576, size=10, n_snakes=3, rules="cut", reps=1)
else:
    for v in ("ws", "split", "fused"):
        print("== lane=%s" % v, flush=True)
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, SNK_DEBUG="lane=" + v))
