"""Key metrics of one kernel from `ncu -i X.ncu-rep --page raw --csv` (development aid)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'sm__throughput.avg.pct', 'smsp__issue_active.avg.pct', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct', 'sm__inst_executed_pipe_fma', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct', 'lts__throughput.avg.pct',
        'sm__warps_active.avg.per_cycle_active', 'launch__occupancy_limit', 'launch__registers', 'launch__waves', 'sm__maximum_warps', 'launch__grid_size', 'launch__block_size',
        'sm__ctas_launched', 'achieved_occupancy', 'sm__cycles_active.avg', 'l1tex__m_xbar2l1tex_read_bytes.sum.per_second', 'smsp__cycles_active.avg']
for h, u, v in zip(hdr, units, vals):
    if any(w in h for w in want):
        try:
            if float(v.replace(',', '')) == 0: continue
        except ValueError:
            pass
        print(f"{h:95s} {u:12s} {v}")
