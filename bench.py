"""bench.py -- agent-steps/s of the batched multi-snake env step (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path
    python bench.py --impl reference [--gpus N] --steps K ...    # the reference's CPU path (oracle port)
    torchrun --nproc-per-node N bench.py --gpus N ...            # one rank per GPU (weak scaling)

Workload: BASELINE.json configs[3] -- 2 snakes on 19x19, classic rules, 1M envs sharded over 8
GPUs, i.e. 131072 envs per GPU (weak scaling), uniform random actions from the Philox action
stream (already resident in HBM), K = S = 2 views of 21x21x3 uint8 per env.  One "step" = one
pass of the fused step kernel over every env of the rank.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ENVS_PER_GPU = 131072
SIZE, N_SNAKES, RULES = 19, 2, "classic"
WORKLOAD = "2-snake 19x19 classic, %d envs per GPU (BASELINE configs[3]: 1M envs over 8 GPUs), K=S=2 views" % ENVS_PER_GPU
METRIC, UNIT = "agent-steps/sec", "agent-steps/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


# ------------------------------------------------------------------ CPU reference path (oracle port)
def _cpu_worker(remote, seed, rank, inner_obs):
    """SubprocVecEnv worker (subproc_vec_env.py:7-28): one env instance per process; step, reset
    at once on done, send (ob, reward, done, info) back through the pipe."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import snake_oracle as so
    env = so.SnakeOracle(SIZE, N_SNAKES, N_SNAKES, N_SNAKES, RULES, draws=so.PhiloxDraws(seed, rank))
    ep_r, ep_l = 0.0, 0
    while True:
        cmd, data = remote.recv()
        if cmd == "step":
            ob, r, d, info = env.step(data)
            ep_r += r
            ep_l += 1
            if d:
                info["episode"] = {"r": ep_r, "l": ep_l}
                ob = env.reset()
                ep_r, ep_l = 0.0, 0
            remote.send((ob, r, d, info))
        elif cmd == "reset":
            remote.send(env.reset())
        else:
            remote.close()
            break


class CpuSubprocVecEnv(object):
    """The reference's vectorised CPU path: utils.make_basic_env -> SubprocVecEnv of Monitor(SnakeEnv)
    (utils.py:34-49), with the oracle port standing in for gym-snake (the reference tree does not
    travel to the GPU box)."""

    def __init__(self, n_procs, seed=0):
        import numpy as np
        self.np = np
        ctx = mp.get_context("fork")
        self.remotes, self.ps = [], []
        for i in range(n_procs):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_cpu_worker, args=(b, seed, i, True), daemon=True)
            p.start()
            b.close()
            self.remotes.append(a)
            self.ps.append(p)
        self.n = n_procs

    def reset(self):
        for r in self.remotes:
            r.send(("reset", None))
        return self.np.stack([r.recv() for r in self.remotes])

    def step(self, actions):
        for r, a in zip(self.remotes, actions):
            r.send(("step", a))
        obs, rews, dones, infos = zip(*[r.recv() for r in self.remotes])
        return self.np.stack(obs), self.np.stack(rews), self.np.stack(dones), infos

    def close(self):
        for r in self.remotes:
            r.send(("close", None))
        for p in self.ps:
            p.join(timeout=5)


def time_cpu_path(seconds, n_procs=None, steps=None, warmup_steps=20):
    """env-steps/s of the SubprocVecEnv-style CPU path over `n_procs` workers (default: all cores)."""
    import numpy as np
    n_procs = n_procs or os.cpu_count() or 1
    venv = CpuSubprocVecEnv(n_procs)
    venv.reset()
    rng = np.random.RandomState(0)
    for _ in range(warmup_steps):
        venv.step(rng.randint(0, 5, size=(n_procs, N_SNAKES)))
    t0 = time.perf_counter()
    n = 0
    while True:
        venv.step(rng.randint(0, 5, size=(n_procs, N_SNAKES)))
        n += 1
        if steps is not None:
            if n >= steps:
                break
        elif time.perf_counter() - t0 >= seconds:
            break
    dt = time.perf_counter() - t0
    venv.close()
    return n * n_procs / dt, n, dt, n_procs


def time_c_oracle(seconds=3.0, n=4096):
    """The C restatement on one core (context only: a far stronger CPU baseline than the reference's Python)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle
    co = c_oracle.COracle(n, size=SIZE, n_snakes=N_SNAKES, rules=RULES)
    co.reset()
    acts = [c_oracle.gen_actions(co.cfg, t, 1) for t in range(16)]
    t0 = time.perf_counter()
    k = 0
    while time.perf_counter() - t0 < seconds:
        co.step(acts[k % 16])
        k += 1
    return k * n / (time.perf_counter() - t0)


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # one host: rank 0 alone measures the CPU path
    cores = os.cpu_count() or 1
    # calibrate so that warmup + K steps end within a couple of minutes
    rate, _, _, _ = time_cpu_path(3.0, cores)
    budget = 60.0
    inner = max(1, int(budget * rate / (cores * max(args.steps, 1))))
    venv_steps = inner * args.steps
    for _ in range(1):
        time_cpu_path(0, cores, steps=max(1, inner * min(args.warmup, 3)))
    value, n, dt, _ = time_cpu_path(0, cores, steps=venv_steps)
    agent = value * N_SNAKES
    sample = "%d SubprocVecEnv-style steps over %d worker processes (1 env each), %.1f s; CPU %s" % (n, cores, dt, cpu_model())
    line = {
        "impl": "reference", "metric": METRIC, "value": agent, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "envs": cores, "reference_step": "%d vec-steps x %d envs" % (inner, cores)},
        "cpu_baseline": {"value": agent, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": agent, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.005):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.stop_flag, self.max_mhz = [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ------------------------------------------------------------------ this repo's CUDA path
def run_ours(args):
    # stdout carries exactly ONE JSON line: everything libraries print (the NCCL version banner goes
    # to fd 1) is sent to stderr, the line is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import snakes_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("SNK_NCCL_DEBUG", "WARN")  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    N = ENVS_PER_GPU
    env = snakes_b200.SnakeVecEnv(N, size=SIZE, n_snakes=N_SNAKES, rules=RULES, seed=0, device=local, env_id_base=rank * N)
    S = env.S
    env.reset()
    K, W = args.steps, args.warmup
    n_act = min(K, 256)  # distinct action batches resident in HBM, cycled
    acts = torch.empty((n_act, N, S), dtype=torch.int8, device=dev)
    for t in range(n_act):
        env.gen_actions(t, 1, out=acts[t])
    for t in range(W):
        env.step_async(acts[t % n_act]); env.step_wait()
    env.reset_stats()
    l0 = env.launch_count()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(K):
        env.step_async(acts[t % n_act]); env.step_wait()
    e1.record()
    barrier()
    clocks = sampler.result()
    launches = env.launch_count() - l0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    stats = env.stats(reduce=True)  # NCCL all-reduce of the 8 episode-stat doubles: the path's only collective

    # ---- second action stream: fruit-seeking policy computed on the device every step
    # (one extra small kernel per step, inside the timed region); snakes get long, resets get rare
    Ks = max(50, K // 4)
    for t in range(300):
        env.step_async(env.gen_scripted_actions(t, 7)); env.step_wait()
    env.reset_stats()
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for t in range(Ks):
        env.step_async(env.gen_scripted_actions(300 + t, 7)); env.step_wait()
    s1.record()
    barrier()
    ms_s = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_s, op=dist.ReduceOp.MAX)
    ms_s = float(ms_s.item())
    stats_s = env.stats(reduce=True)
    sum_len_s = stats_s["body_cells"] / max(stats_s["env_steps"], 1.0)
    alg_s = env.algorithmic_bytes_per_step(sum_len_s)
    scripted = {"value": float(N) * world * Ks * S / (ms_s * 1e-3), "unit": UNIT, "ms_per_step": ms_s / Ks, "steps": Ks,
                "mean_sum_len": sum_len_s, "algorithmic_bytes_per_env_step": alg_s,
                "episodes_per_env_step": stats_s["episodes"] / max(stats_s["env_steps"], 1.0),
                "note": "scripted fruit-seeking policy kernel + step kernel per step, both inside the timed region"}
    env.check_errors()
    mean_sum_len = stats["body_cells"] / max(stats["env_steps"], 1.0)
    alg_bytes = env.algorithmic_bytes_per_step(mean_sum_len)
    total_env_steps = float(N) * world * K
    value = total_env_steps * S / (ms * 1e-3)

    # ---- e2e: the same step through the host-buffer entry point (numpy in / numpy out, like
    # SubprocVecEnv.step_wait): pinned actions H2D, kernel, obs + reward + done + num_snakes D2H
    Ke = max(3, min(K, args.e2e_steps))
    henv = snakes_b200.SnakeVecEnv(N, size=SIZE, n_snakes=N_SNAKES, rules=RULES, seed=0, device=local,
                                   env_id_base=rank * N, host_io=True)
    henv.reset()
    h_acts = [acts[t % n_act].cpu().numpy() for t in range(8)]
    for t in range(3):
        henv.step(h_acts[t % 8])
    barrier()
    t0 = time.perf_counter()
    for t in range(Ke):
        henv.step(h_acts[t % 8])
    torch.cuda.synchronize(dev)
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    # obs stays in HBM for an on-device learner: actions H2D + reward/done D2H only
    h_a = torch.as_tensor(h_acts[0]).pin_memory()
    barrier()
    t0 = time.perf_counter()
    for t in range(Ke * 4):
        env.step_async(h_a)
        _, rew, done, _ = env.step_wait()
        rew_h, done_h = rew.cpu(), done.cpu()
    dt2 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dist.all_reduce(dt2, op=dist.ReduceOp.MAX)
    e2e_value = float(N) * world * Ke * S / float(dt.item())
    e2e_resident = float(N) * world * Ke * 4 * S / float(dt2.item())
    h2d = N * S
    d2h = N * env.V * env.V * 3 * env.K + N * 4 + 2 * N

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
    launch_s = ms * 1e-3 / K
    achieved = alg_bytes * N / launch_s / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_total": N * world, "rules": RULES, "size": SIZE, "n_snakes": S,
                       "obs": "uint8 [N,21,21,6]", "actions": "Philox uniform{0..4}, %d batches resident in HBM" % n_act,
                       "l2": "per-step working set (347 MB obs stream + state) exceeds the 126 MB L2; no explicit flush",
                       "kernel": env.launch_info(), "mean_sum_len": mean_sum_len},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": Ke, "path": "snk_step_host: pinned host actions -> H2D -> fused kernel -> obs+reward+done+num_snakes D2H"},
            "e2e_obs_resident": {"value": e2e_resident, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": N * 5,
                                 "note": "obs stays in HBM for an on-device learner; actions H2D, reward+done D2H every step"},
            "gpu_launches": int(launches) * world,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": env.launch_info()["kernel"],
                         "algorithmic_bytes_per_env_step": alg_bytes, "env_steps_per_launch": N,
                         "launch_us": launch_s * 1e6},
            "episode_stats": stats,
            "scripted_policy": scripted,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            v, n, dts, _ = time_cpu_path(args.cpu_seconds, cores)
            line["cpu_baseline"] = {
                "value": v * S, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "%d SubprocVecEnv-style steps over %d worker processes (1 env each, oracle port of gym-snake), %.1f s; CPU %s"
                          % (n, cores, dts, cpu_model()),
                "c_oracle_1core": time_c_oracle(2.0) * S,
            }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    env.close(); henv.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # launched without torchrun: re-exec under it, one rank per GPU
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
