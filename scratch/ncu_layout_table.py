"""Per-kernel table from the csv of scratch/ncu_layout.sh."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, im, iv = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value')
iid = hdr.index('ID')
data = {}
for r in rows[1:]:
    key = (int(r[iid]), r[ik].split('(')[0])
    data.setdefault(key, {})[r[im]] = float(r[iv].replace(',', ''))
print('# %s' % sys.argv[2])
print('%-34s %8s %7s %7s %9s %9s %6s %6s %6s %6s %5s' % (
    'kernel', 'us', 'rd MB', 'wr MB', 'ld sec/rq', 'st sec/rq', 'L1hit', 'L2hit', 'longsb', 'lanes', 'regs'))
for (i, name), m in sorted(data.items()):
    g = lambda k: m.get(k, 0.0)
    ld = g('l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum') / max(g('l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum'), 1)
    st = g('l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum') / max(g('l1tex__t_requests_pipe_lsu_mem_global_op_st.sum'), 1)
    print('%-34s %8.1f %7.1f %7.1f %9.2f %9.2f %6.1f %6.1f %6.2f %6.2f %5d' % (
        name[:34], g('gpu__time_duration.sum') / 1e3, g('dram__bytes_read.sum') / 1e6,
        g('dram__bytes_write.sum') / 1e6, ld, st, g('l1tex__t_sector_hit_rate.pct'),
        g('lts__t_sector_hit_rate.pct'),
        g('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'),
        g('smsp__thread_inst_executed_per_inst_executed.ratio'), g('launch__registers_per_thread')))
