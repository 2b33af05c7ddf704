#!/bin/bash
# ncu_quick.sh LIB TAG [short|long]: time / instructions / issue / DRAM bytes of the lane kernel for one build
export SNK_LIB=$1
reg=${3:-long}
if [ $reg = long ]; then skip=403; script=tools/probe_long_ncu.py; else skip=55; script=tools/probe_short_ncu.py; fi
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum --clock-control none -k regex:k_step_lane -s $skip -c 1 --csv --log-file gpurun_out/q_$2_$reg.csv python $script > /dev/null 2>&1
grep k_step_lane gpurun_out/q_$2_$reg.csv | awk -F'","' '{printf "%s %s %s %s %s | ", "'$2'", "'$reg'", $(NF-2), $(NF-1), $NF} END {print ""}' | tr -d '"'
