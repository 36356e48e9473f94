#!/usr/bin/env python
"""Condense an .ncu-rep into a small CSV of the metrics the profiles/ README quotes:
    python benchmarks/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_ncu_summary.csv"""
import csv
import subprocess
import sys

KEEP = ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput',
        'launch__registers_per_thread', 'launch__occupancy_limit', 'launch__grid_size', 'launch__block_size',
        'launch__waves_per_multiprocessor', 'sm__warps_active.avg.pct', 'sm__throughput.avg.pct',
        'nvlrx__bytes.sum', 'nvltx__bytes.sum', 'nvlrx__bytes_data_user.sum', 'nvltx__bytes_data_user.sum',
        'nvlink__bandwidth', 'nvlink__count_physical', 'lts__t_bytes.sum', 'lts__t_sector_hit_rate',
        'sm__pipe_tensor', 'sm__inst_executed_pipe_tensor', 'smsp__cycles_active.avg',
        'launch__shared_mem_per_block', 'sm__cycles_elapsed.max', 'l1tex__data_pipe_lsu_wavefronts_mem_shared')


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [i for i, h in enumerate(hdr) if any(k in h for k in KEEP)]
    name_i = hdr.index('Kernel Name')
    with open(out, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['launch', 'kernel', 'metric', 'value', 'unit'])
        for n, r in enumerate(rows[2:]):
            for i in cols:
                if r[i] != '':
                    w.writerow([n, r[name_i][:60], hdr[i], r[i], units[i]])
    print('wrote', out, len(rows) - 2, 'launches')


if __name__ == '__main__':
    main()
