#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: python tools/ncu_summary.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    print("kernel:", d.get("Kernel Name", ("", "?"))[1][:100])
    keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
            'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
            'sm__cycles_elapsed.avg', 'smsp__inst_executed.sum',
            'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
            'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
            'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
            'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
            'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
            'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum',
            'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
            'lts__throughput.avg.pct_of_peak_sustained_elapsed',
            'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum']
    for k in keys:
        if k in d:
            print(f"  {k:72s} {d[k][0]:16s} {d[k][1]}")
    print("  stalls (warps per issue-active cycle):")
    st = []
    for h in hdr:
        if 'average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h:
            try:
                st.append((float(d[h][1]), h.split('stalled_')[1].replace('_per_issue_active.ratio', '')))
            except ValueError:
                pass
    for v, n in sorted(st, reverse=True)[:8]:
        print(f"    {n:28s} {v:.3f}")
