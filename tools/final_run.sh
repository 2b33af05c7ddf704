#!/bin/bash
# One GPU session that produces everything profiles/ cites for the final state of a round (prefix $1, default r02_final):
# parity tests, the bench line (ours + reference arm), the ncu launch list of the bench command and one full capture per
# hot kernel (headline fused kernel; the two kernels of the long-body form).
P=${1:-r02_final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${P}_pytest_gpu.txt 2>&1; tail -2 gpurun_out/${P}_pytest_gpu.txt
python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/${P}_bench.json
python bench.py --action-batches 256 --no-cpu-baseline --no-configs > gpurun_out/${P}_bench_256batches.json 2> /dev/null; echo "bench256 rc=$?"; cut -c1-200 gpurun_out/${P}_bench_256batches.json
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs > gpurun_out/${P}_bench_20steps.json 2> /dev/null; echo "bench20 rc=$?"; cut -c1-200 gpurun_out/${P}_bench_20steps.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${P}_bench_ref.json 2> gpurun_out/${P}_bench_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/${P}_bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs --e2e-steps 3 > gpurun_out/${P}_ncu1.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_step_lane -s 55 -c 1 -o gpurun_out/${P}_lane -f python tools/probe_short_ncu.py > gpurun_out/${P}_ncu2.log 2>&1; echo "lane full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_lane_ -s 400 -c 2 -o gpurun_out/${P}_long -f python tools/probe_long_ncu.py > gpurun_out/${P}_ncu3.log 2>&1; echo "long full rc=$?"
python tools/ab.py short long long40 c3 c2 1m atari rows > gpurun_out/${P}_ab.txt 2>&1; cut -c1-120 gpurun_out/${P}_ab.txt
python tools/bw_probe.py > gpurun_out/${P}_bw_probe.txt 2>&1; tail -2 gpurun_out/${P}_bw_probe.txt
