"""CPU: the bench's CPU legs -- the reference arm's harness around the reference's own SubprocVecEnv of Monitor(SnakeEnv)
(src/utils.py:34-49, src/baselines/common/vec_env/subproc_vec_env.py:31) when a reference tree is on the box, the
repo's port under the same harness otherwise -- run, report which one they timed, and print the contract's line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_cpu_path_times_reference_or_port():
    import bench
    r = bench.time_cpu_path(1.0, n_procs=2, warmup_steps=3)
    assert r["kind"] in ("reference", "port") and r["procs"] == 2
    assert r["env_steps_per_s"] > 100 and r["vec_steps"] > 10
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_loader
    assert (r["kind"] == "reference") == ref_loader.available()
    inproc = bench.time_reference_inprocess(0.5)
    assert (inproc is None) == (not ref_loader.available())
    if inproc is not None:
        assert inproc > 100
    assert bench.time_c_oracle(0.3, n=256) > 1e4


def test_reference_arm_line_on_other_ranks_is_silent():
    """Under torchrun only rank 0 measures and prints; the other ranks exit 0 without work."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"],
                         env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
