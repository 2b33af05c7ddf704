timeout 1400 python tools/selfplay_curve.py 2 19 32 1e7 gpurun_out/selfplay_2x19_1gpu > gpurun_out/selfplay_2x19_1gpu.log 2>&1
echo rc=$?
tail -3 gpurun_out/selfplay_2x19_1gpu.log | cut -c1-600
