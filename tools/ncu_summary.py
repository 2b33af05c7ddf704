"""Key metrics of each kernel instance in an ncu report: python tools/ncu_summary.py <rep>"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'sm__inst_executed_pipe_lsu.sum']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print("%-70s %-12s %s" % (w, units[i], [r[i][:40] for r in rows[2:]]))
print("---- stall reasons (pct of warp-active), instance 0")
st = []
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('per_warp_active.pct'):
        st.append((float(rows[2][i]), h.replace('smsp__warps_issue_stalled_', '').replace('_per_warp_active.pct', '')))
for v, h in sorted(st, reverse=True)[:12]:
    print("  %6.1f  %s" % (v, h))
