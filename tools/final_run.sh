#!/bin/bash
# One GPU session that produces everything profiles/ cites for the final state of a round:
# parity tests, the bench line (ours + reference arm), the ncu launch list and one full capture per hot kernel.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/final_pytest_gpu.log 2>&1; tail -2 gpurun_out/final_pytest_gpu.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/final_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/final_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3 > gpurun_out/final_ncu1.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_step_lane -s 55 -c 1 -o gpurun_out/final_lane -f python tools/probe_short_ncu.py > gpurun_out/final_ncu2.log 2>&1; echo "lane full rc=$?"
cat > /tmp/rows_probe.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch, snakes_b200
env = snakes_b200.SnakeVecEnv(16384, size=64, n_snakes=16, rules="cut"); env.reset()
for t in range(12): env.step(env.gen_actions(t, 1))
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:k_step_rows -s 10 -c 1 -o gpurun_out/final_rows -f python /tmp/rows_probe.py > gpurun_out/final_ncu3.log 2>&1; echo "rows full rc=$?"
python tools/probe.py > gpurun_out/final_probe.txt 2>&1; cat gpurun_out/final_probe.txt | cut -c1-170
python tools/ab_rows.py 32768 > gpurun_out/final_rows_32768.txt 2>&1; cat gpurun_out/final_rows_32768.txt | cut -c1-120
# long-body regime (fruit-seeking policy): one full capture of the lane kernel, and what the box sustains for pure writes
python tools/probe_long_ncu.py && ncu --set full --clock-control none --import-source on -k regex:k_step_lane -s 403 -c 1 -o gpurun_out/final_long -f python tools/probe_long_ncu.py > gpurun_out/final_ncu_long.log 2>&1; echo "long full rc=$?"
python tools/bw_probe.py > gpurun_out/final_bw_probe.txt 2>&1; tail -2 gpurun_out/final_bw_probe.txt
python tools/ab.py short long c3 > gpurun_out/final_ab.txt 2>&1; grep -E "^short|^long" gpurun_out/final_ab.txt | cut -c1-100
