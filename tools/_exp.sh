python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rows or large or 64" > gpurun_out/s15_pytest.txt 2>&1; tail -2 gpurun_out/s15_pytest.txt
SNK_LIB=variants/libsnk_base.so python tools/ab.py rows > gpurun_out/s15_ab_base.txt 2>&1; cut -c1-130 gpurun_out/s15_ab_base.txt
python tools/ab.py rows > gpurun_out/s15_ab_new.txt 2>&1; cut -c1-130 gpurun_out/s15_ab_new.txt
