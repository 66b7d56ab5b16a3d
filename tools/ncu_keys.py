#!/usr/bin/env python
"""Print selected metrics of an `ncu --page raw --csv` export, one column per profiled launch.
usage: ncu -i rep.ncu-rep --page raw --csv > raw.csv; python tools/ncu_keys.py raw.csv [extra-metric-regex]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__cycles_active.avg',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor', 'launch__grid_size',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.pct', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_drain_per_issue_active.ratio',
        ]
extra = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
for i, k in enumerate(hdr):
    if k in KEYS or (extra and extra.search(k)):
        print("%-90s %-12s %s" % (k, units[i], " | ".join(r[i][:48] for r in data)))
