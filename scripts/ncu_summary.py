"""Prints the handful of ncu raw metrics we track for a kernel (reads a .ncu-rep here, no GPU)."""
import csv, subprocess, sys
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'launch__grid_size', 'launch__block_size', 'sm__cycles_elapsed.max', 'launch__occupancy_limit_registers',
        'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_fmaheavy.sum',
        'sm__inst_executed_pipe_cbu.sum', 'sm__inst_executed_pipe_adu.sum', 'sm__inst_executed_pipe_uniform.sum',
        'lts__t_sectors.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg']
for path in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print('===', path, vals[hdr.index('Kernel Name')][:60] if 'Kernel Name' in hdr else '')
        for h, u, v in zip(hdr, units, vals):
            if h in want or ('issue_stalled' in h and h.endswith('per_issue_active.ratio') and float(v or 0) > 0.05):
                print('  %-78s %-10s %s' % (h, u, v))
