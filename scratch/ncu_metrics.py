"""Print key metrics per kernel from an .ncu-rep (raw page)."""
import csv, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'Grid Size', 'gpu__time_duration.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print('-----')
    for w in want:
        if w in idx:
            print('%-82s %18s %s' % (w, r[idx[w]], units[idx[w]]))
