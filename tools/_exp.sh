for m in 5 7 3; do SNK_DEBUG=restore_mode=$m python tools/ab.py long > gpurun_out/s8_mode$m.txt 2>&1; cut -c1-110 gpurun_out/s8_mode$m.txt; done
