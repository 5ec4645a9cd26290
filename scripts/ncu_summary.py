#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of metrics profiles/*.md quote."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, unit = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, unit))
        print('KERNEL', d['Kernel Name'][:70], d.get('Block Size'), d.get('Grid Size'))
        for k in KEYS:
            print(f'   {k:62s} {d.get(k)} {u.get(k, "")}')
        st = [(float(v or 0), h) for h, v in d.items() if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h]
        st.sort(reverse=True)
        tot = sum(v for v, _ in st) or 1
        print('   stalls: ' + ', '.join(f'{h.split("stalled_")[1]} {100 * v / tot:.1f}%' for v, h in st[:8]))


if __name__ == '__main__':
    main(sys.argv[1])
