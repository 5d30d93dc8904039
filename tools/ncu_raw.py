#!/usr/bin/env python
"""Selected metrics per launch from `ncu -i REP --page raw --csv` output (file argument)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'launch__grid_size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__occupancy_limit_shared_mem', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']
want += sys.argv[2:]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(w, units[i], [r[i][:28] for r in data])
for i, h in enumerate(hdr):
    if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct') is False and 'ratio' in h and 'not_issued' not in h:
        vals = [float(r[i].replace(',', '')) if r[i] else 0 for r in data]
        if max(vals) > 0.4:
            print(h.split('issue_stalled_')[1].replace('.ratio', ''), [round(v, 2) for v in vals])
