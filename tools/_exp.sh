python -m pytest tests -m gpu -x -q > gpurun_out/exp8_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/exp8_pytest.txt
tail -5 gpurun_out/exp8_pytest.txt
