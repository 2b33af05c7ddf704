python -m pytest tests/test_gpu_parity.py tests/test_gpu_forms.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/s18_pytest.txt 2>&1; tail -2 gpurun_out/s18_pytest.txt
SNK_LIB=variants/libsnk_base.so python tools/ab.py short c2 > gpurun_out/s18_ab_base.txt 2>&1; cut -c1-110 gpurun_out/s18_ab_base.txt
python tools/ab.py short c2 c3 long > gpurun_out/s18_ab_new.txt 2>&1; cut -c1-110 gpurun_out/s18_ab_new.txt
