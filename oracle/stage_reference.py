"""Stage the UNMODIFIED reference into baseline/_ref/ so that it travels to the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY.  /root/reference exists in the build container only; `baseline/_ref/` is git-ignored
(never in history) but not gpurun-ignored, so `bench.py --impl reference` and the `cpu_baseline` leg can time the
reference's OWN CPU path -- SubprocVecEnv of Monitor(SnakeEnv), src/utils.py:34-49 -- on the GPU box's host cores.
This is the one offline install the bench contract allows: `pip install --no-index --no-build-isolation --no-deps
--target baseline/_ref <copy of /root/reference/src>`.  The reference ships a setup.py for gym-snake only
(src/gym-snake/setup.py); the copy under /tmp gets a generated one that also names `baselines`, `config` and `utils`,
all of which stay byte-identical to the reference's files.  Nothing is copied into tracked paths.

    python oracle/stage_reference.py          # no-op when /root/reference is absent or the install is current
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("SNK_REFERENCE_ROOT", "/root/reference")
TARGET = os.path.join(ROOT, "baseline", "_ref")

SETUP = '''from setuptools import setup, find_packages
setup(name="snake_selfplay_reference", version="0.0.0",
      packages=["gym_snake"] + ["gym_snake." + p for p in find_packages("gym-snake/gym_snake")] +
               ["baselines"] + ["baselines." + p for p in find_packages("baselines")],
      package_dir={"gym_snake": "gym-snake/gym_snake", "baselines": "baselines"},
      py_modules=["config", "utils"])
'''


def staged():
    return os.path.isdir(os.path.join(TARGET, "gym_snake", "envs")) and os.path.isdir(os.path.join(TARGET, "baselines", "common", "vec_env"))


def stage(force=False, quiet=True):
    src = os.path.join(REFERENCE, "src")
    if not os.path.isdir(os.path.join(src, "gym-snake", "gym_snake")):
        return False
    if staged() and not force:
        return True
    tmp = tempfile.mkdtemp(prefix="snk_ref_")
    try:
        work = os.path.join(tmp, "src")
        shutil.copytree(src, work, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        with open(os.path.join(work, "setup.py"), "w") as f:
            f.write(SETUP)
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        os.makedirs(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--target", TARGET, work]
        subprocess.check_call(cmd, stdout=subprocess.DEVNULL if quiet else None)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return staged()


if __name__ == "__main__":
    ok = stage(force="--force" in sys.argv, quiet=False)
    print("reference staged at %s" % TARGET if ok else "reference tree not available: nothing staged")
