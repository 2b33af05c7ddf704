"""Per-phase warp cycles of k_step_lane (tools/build_variant.sh X - -DSNK_PHASE_TIMING [-DSNK_PHASE_LOGIC]; SNK_LIB=variants/libsnk_X.so).
Plain build flag: logic / paint / store+wait / un-paint.  With -DSNK_PHASE_LOGIC the four slots hold the parts of the logic
instead: move+push / respawns / death test / tail+reset."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, snakes_b200
N = 131072
names = ("up to death test", "clear+bitmap", "outputs", "reset") if "fine" in os.environ.get("SNK_LIB", "") else ("move+push", "respawns", "death test", "tail+reset") if ("phl" in os.environ.get("SNK_LIB", "") or "_x_" in os.environ.get("SNK_LIB", "")) else ("logic", "paint", "store+wait", "unpaint")
def show(tag, env, T):
    st = env.stats(False)
    info = env.launch_info(); warps = info["grid"] * info["block"] // 32
    tot, A, B, Cc, D = (st[k] / warps / T for k in ("length_sum", "fruits", "deaths", "body_cells", "draws"))
    print("%s %s: per warp per launch: total %.0f cyc | %s %.0f  %s %.0f  %s %.0f  %s %.0f  (other %.0f)" % (
        os.path.basename(os.environ.get("SNK_LIB", "default")), tag, tot, names[0], A, names[1], B, names[2], Cc, names[3], D, tot - A - B - Cc - D), flush=True)
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
