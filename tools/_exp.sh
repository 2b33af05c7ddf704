python bench.py --steps 400 --warmup 40 --no-configs --cpu-seconds 4 > gpurun_out/s19_bench.json 2> gpurun_out/s19_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/s19_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/s19_bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step')}, d['roofline']['frac'])
print(d['with_action_gen']); print(d['cpu_baseline'])
PY
