#!/usr/bin/env python3
"""Print the roofline-relevant metrics of every kernel in an .ncu-rep (reads `ncu -i <rep> --page raw --csv`).

    python bench/ncu_summary.py gpurun_out/mx_ncu.ncu-rep > profiles/ncu/<name>.txt
"""
import csv
import io
import subprocess
import sys

KEEP = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg.per_second', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed.sum', 'launch__grid_size', 'launch__block_size', 'launch__cluster_dim_x', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.pct', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio']


def main():
    rep = sys.argv[1]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    names, units = rows[head], rows[head + 1]
    col = {n: i for i, n in enumerate(names)}
    for r in rows[head + 2:]:
        if len(r) < len(names):
            continue
        print('## %s  (id %s)' % (r[col['Kernel Name']][:110], r[col['ID']]))
        for k in KEEP:
            if k in col:
                print('%-76s %-18s %s' % (k, units[col[k]], r[col[k]]))
        print()


if __name__ == '__main__':
    main()
