"""Prints the handful of counters the profiles/ summaries quote from an `ncu --page raw --csv` dump."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__grid_size', 'launch__block_size', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__waves_per_multiprocessor', 'sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_op_branch.sum', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active']
rows = list(csv.reader(open(sys.argv[1])))
h, u = rows[0], rows[1]
for v in rows[2:]:
    print('---', v[h.index('Kernel Name')][:80] if 'Kernel Name' in h else '')
    for i, n in enumerate(h):
        try:
            x = float(v[i].replace(',', ''))
        except ValueError:
            continue
        if n in WANT or ('issue_stalled' in n and 'per_issue_active' in n and x > 0.2):
            print(f'{n} [{u[i]}] = {v[i]}')
