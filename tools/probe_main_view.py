"""What the main-view rollout target costs (snk_set_main_view_target: a gather kernel behind the step kernel).
Eager steps, CUDA events, 131 072 envs of 2 snakes on 19x19; with / without the target, native and atari84."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, snakes_b200
N, T = 131072, 200
for mode in ("native", "atari84"):
    env = snakes_b200.SnakeVecEnv(N, size=19, n_snakes=2, obs_mode=mode)
    env.reset()
    acts = [env.gen_actions(t, 1).clone() for t in range(16)]
    side = env.obs.shape[1]
    main = torch.empty((N, side, side, 3), dtype=torch.uint8, device="cuda")
    for with_target in (False, True):
        env.set_main_view_target(main if with_target else None)
        for t in range(20):
            env.step_async(acts[t % 16]); env._pending = False
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for t in range(T):
            env.step_async(acts[t % 16]); env._pending = False
        e1.record(); torch.cuda.synchronize()
        print("%s main_view_target=%s: %.1f us per step" % (mode, with_target, e0.elapsed_time(e1) / T * 1e3))
    if mode == "native":
        assert torch.equal(main, env.obs[..., 0:3])
    env.set_main_view_target(None)
    env.close()
