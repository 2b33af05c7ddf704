"""Per-phase warp cycles of k_step_lane (tools/build_variant.sh phase -DSNK_PHASE_TIMING; SNK_LIB=variants/libsnk_phase.so): logic / paint / store+wait / un-paint."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, snakes_b200
N = 131072
def show(tag, env, T):
    st = env.stats(False)
    info = env.launch_info(); warps = info["grid"] * info["block"] // 32
    tot, A, B, Cc, D = (st[k] / warps / T for k in ("length_sum", "fruits", "deaths", "body_cells", "draws"))
    print("%s: per warp per launch: total %.0f cyc | logic %.0f  paint %.0f  store+wait %.0f  unpaint %.0f  (other %.0f)" % (tag, tot, A, B, Cc, D, tot - A - B - Cc - D))
env = snakes_b200.SnakeVecEnv(N, size=19, n_snakes=2); env.reset()
acts = [env.gen_actions(t, 1).clone() for t in range(16)]
for t in range(100): env.step(acts[t % 16])
env.reset_stats()
for t in range(100): env.step(acts[t % 16])
show("short", env, 100)
for t in range(400): env.step(env.gen_scripted_actions(t, 7))
env.reset_stats()
for t in range(50): env.step(env.gen_scripted_actions(400 + t, 7))
show("long ", env, 50)
