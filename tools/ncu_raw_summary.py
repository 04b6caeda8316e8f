#!/usr/bin/env python3
"""Filter the metrics quoted in profiles/README.md out of `ncu -i x.ncu-rep --page raw --csv`.
usage: tools/ncu_raw_summary.py raw.csv > profiles/<name>_raw_metrics.txt"""
import csv, re, sys
KEYS = [r'^gpu__time_duration\.sum$', r'^launch__(grid_size|block_size|registers_per_thread|shared_mem_per_block_static|shared_mem_per_block_dynamic|waves_per_multiprocessor|occupancy_limit_\w+)$',
        r'^smsp__inst_executed\.sum$', r'^smsp__issue_active\.avg\.per_cycle_active$', r'^sm__inst_executed_pipe_(fma|alu|xu|fp64|lsu|fmaheavy|fmalite)\.avg\.pct_of_peak_sustained_active$',
        r'^sm__pipe_(fp64|fma|alu|xu)_cycles_active\.avg\.pct_of_peak_sustained_active$',
        r'^dram__bytes_(read|write)\.sum$', r'^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$', r'^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$',
        r'^smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio$', r'^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$',
        r'^l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$', r'^smsp__inst_executed_op_shared_(ld|st)\.sum$',
        r'^sm__cycles_active\.(avg|min|max)$', r'^sm__warps_active\.avg\.per_cycle_active$', r'^smsp__cycles_active\.avg$',
        r'^lts__t_sector_hit_rate\.pct$', r'^l1tex__t_sector_hit_rate\.pct$', r'^sm__cycles_elapsed\.max$']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
    print(f"# kernel: {name}")
    for h, u, v in sorted(zip(hdr, units, vals)):
        if any(re.search(k, h) for k in KEYS) and not (h.startswith('smsp__average_warps_issue_stalled') and float(v or 0) < 0.005):
            print(f"{h:95s} {u:16s} {v}")
