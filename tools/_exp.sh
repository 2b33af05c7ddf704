python -m pytest tests/test_gpu_api.py -m gpu -x -q -k "scalars or host_io" > gpurun_out/s5_pytest.txt 2>&1; tail -3 gpurun_out/s5_pytest.txt
python bench.py --steps 400 --warmup 40 --no-cpu-baseline --no-configs > gpurun_out/s5_bench.json 2> gpurun_out/s5_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/s5_bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step')}, d['roofline']['frac'])
for k in ('e2e', 'e2e_main_view', 'e2e_obs_resident'): print(k, d[k]['value'])
PY
python tools/probe_main_view.py 2>&1 | tee gpurun_out/s5_main_view.txt
# configs[2]: launch list, then one full capture of each of its kernels
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -s 60 -c 4 --csv --log-file gpurun_out/s5_c3_launches.csv python tools/probe_c3_split.py > gpurun_out/s5_c3_l.log 2>&1; echo "c3 list rc=$?"
ncu --set full --clock-control none --import-source on -s 60 -c 2 -o gpurun_out/s5_c3_full -f python tools/probe_c3_split.py > gpurun_out/s5_c3_f.log 2>&1; echo "c3 full rc=$?"
grep -v "^==" gpurun_out/s5_c3_launches.csv | cut -d, -f5,9,10,13,15 | head -20
