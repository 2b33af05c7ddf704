P=r02_head
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs --e2e-steps 3 > gpurun_out/${P}_ncu1.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_step_lane -s 55 -c 1 -o gpurun_out/${P}_lane -f python tools/probe_short_ncu.py > gpurun_out/${P}_ncu2.log 2>&1; echo "lane full rc=$?"
