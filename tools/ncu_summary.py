"""Key metrics of each kernel instance in an ncu report: python tools/ncu_summary.py <rep> [> profiles/rNN_ncu_*.txt]

Per instance: the launch / throughput / memory figures bench.py's roofline cites, then the stall table -- cycles a warp
spends stalled per instruction issued (`smsp__average_warps_issue_stalled_*_per_issue_active.ratio`, what the ncu UI shows
as "warp state") and the share of the PC samples per reason (`smsp__pcsamp_warps_issue_stalled_*`)."""
import csv, subprocess, sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__warps_active.avg.per_cycle_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'sm__inst_executed_pipe_lsu.sum']


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


for n, r in enumerate(rows[2:]):
    print("==== instance %d" % n)
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("%-70s %-12s %s" % (w, units[i], r[i][:90]))
    st = [(num(r[i]), h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')])
          for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
    if st:
        print("---- warp cycles per issued instruction, by state (sum = %.2f)" % sum(v for v, _ in st))
        for v, h in sorted(st, reverse=True)[:10]:
            print("  %6.2f  %s" % (v, h))
    pc = [(num(r[i]), h[len('smsp__pcsamp_warps_issue_stalled_'):])
          for i, h in enumerate(hdr) if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued')]
    tot = sum(v for v, _ in pc)
    if tot > 0:
        print("---- PC samples by stall reason (%d samples)" % tot)
        for v, h in sorted(pc, reverse=True)[:10]:
            print("  %5.1f %%  %s" % (100.0 * v / tot, h))
