python -m pytest tests -m gpu -q > gpurun_out/r02_final4_pytest_gpu.txt 2>&1; tail -2 gpurun_out/r02_final4_pytest_gpu.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final4_bench_20steps.json 2> gpurun_out/r02_final4_bench.err; echo "bench rc=$?"; cut -c1-240 gpurun_out/r02_final4_bench_20steps.json
