"""Summarise an .ncu-rep (read here, no GPU needed): python scripts/ncu_summary.py gpurun_out/x.ncu-rep [out.txt]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_fmaheavy.sum', 'sm__inst_executed_pipe_fmalite.sum', 'sm__inst_executed_pipe_lsu.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_wait_per_warp_active.pct', 'smsp__warp_issue_stalled_not_selected_per_warp_active.pct',
        'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct',
        'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct', 'smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct']
out = []
for r in rows[2:]:
    out.append('-----')
    for w in want:
        if w in idx:
            out.append('  %-78s %s %s' % (w, r[idx[w]][:110], units[idx[w]]))
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(txt + "\n")
