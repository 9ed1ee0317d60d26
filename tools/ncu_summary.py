"""Summaries of ncu outputs for profiles/ (run here, on the CPU box).
  python tools/ncu_summary.py launches gpurun_out/launches_r1.csv
  python tools/ncu_summary.py raw gpurun_out/prof.ncu-rep
"""
import collections
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'smsp__warps_eligible.avg.per_cycle_active']


def launches(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if 'Kernel Name' in r:
            hdr, start = r, i
            break
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    gi = hdr.index('Grid Size')
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split('(')[0].split('::')[-1]
        agg.setdefault(name, []).append((float(r[vi].replace(',', '')), r[gi]))
    tot = sum(sum(v for v, _ in l) for l in agg.values())
    print(f"{'kernel':34s} {'n':>4s} {'total us':>10s} {'avg us':>9s} {'share':>7s}  grid")
    for k, l in sorted(agg.items(), key=lambda kv: -sum(v for v, _ in kv[1])):
        s = sum(v for v, _ in l)
        print(f"{k:34s} {len(l):4d} {s / 1e3:10.1f} {s / len(l) / 1e3:9.1f} {s / tot * 100:6.1f}%  {l[0][1]}")


def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    seen = set()
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')].split('(')[0].split('::')[-1]
        if name in seen:
            continue
        seen.add(name)
        print('----', name)
        for k in KEYS:
            if k in hdr:
                print(f"  {k:75s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
        for i, k in enumerate(hdr):
            if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio'):
                try:
                    if float(r[i]) >= 0.2:
                        print(f"  stall {k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:30s} {float(r[i]):8.2f}")
                except ValueError:
                    pass


if __name__ == '__main__':
    {'launches': launches, 'raw': raw}[sys.argv[1]](sys.argv[2])
