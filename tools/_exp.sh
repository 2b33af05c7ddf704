python -m pytest tests/test_gpu_api.py -m gpu -x -q -k "scalars" > gpurun_out/s13_pytest.txt 2>&1; tail -2 gpurun_out/s13_pytest.txt
python bench.py --steps 400 --warmup 40 --no-cpu-baseline --no-configs > gpurun_out/s13_bench.json 2> gpurun_out/s13_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/s13_bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step')}, d['roofline']['frac'])
for k in ('e2e', 'e2e_main_view', 'e2e_obs_resident'): print(k, d[k]['value'])
PY
