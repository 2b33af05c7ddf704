import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import ab
for N in (65536, 262144, 1048576):
    ab.short(N, reps=1, size=10, n_snakes=3, rules="cut")
