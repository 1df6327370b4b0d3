#!/usr/bin/env python
"""Reduce an `ncu --set full` report to the columns DESIGN.md / bench.py cite.

  python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/ncu_x_summary.csv
"""
import csv
import io
import subprocess
import sys

COLS = ['ID', 'Kernel Name', 'launch__grid_size', 'launch__block_size', 'launch__cluster_dim_x',
        'launch__registers_per_thread', 'gpu__time_duration.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__cycles_active.avg', 'sm__cycles_elapsed.max']


def main():
  rep, out = sys.argv[1], sys.argv[2]
  raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
  rows = list(csv.reader(io.StringIO(raw)))
  hdr = rows[0]
  idx = [hdr.index(c) for c in COLS if c in hdr]
  with open(out, 'w', newline='') as f:
    w = csv.writer(f)
    for r in rows:
      w.writerow([r[i] for i in idx])
  print(f'{out}: {len(rows) - 2} launches, {len(idx)} columns')


if __name__ == '__main__':
  main()
