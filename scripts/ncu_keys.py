"""Prints the key raw metrics of every kernel in an .ncu-rep (reads `ncu --page raw --csv` from stdin)."""
import csv, sys
r = list(csv.reader(sys.stdin))
h, u = r[0], r[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active',
        'lts__t_sectors.sum', 'lts__t_bytes.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'sm__cycles_elapsed.max',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']
for row in r[2:]:
    for i, k in enumerate(h):
        if k in keys or ('issue_stalled' in k and k.endswith('per_issue_active.ratio') and float(row[i] or 0) > 0.15):
            print('%-90s %-14s %s' % (k, u[i], row[i][:120]))
    print('-' * 40)
