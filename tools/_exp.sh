python -m pytest tests -m gpu -q > gpurun_out/r02_final3_pytest_gpu.txt 2>&1; tail -2 gpurun_out/r02_final3_pytest_gpu.txt
python bench.py > gpurun_out/r02_final3_bench.json 2> gpurun_out/r02_final3_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r02_final3_bench.json
python -c "import __graft_entry__ as g; g.smoke()"
