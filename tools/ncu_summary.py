#!/usr/bin/env python
"""Text summary of the ncu captures under gpurun_out/prof_final_*.ncu-rep -> profiles/."""
import csv
import glob
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ['launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(unit, 1)


def to_sec(v, unit):
    v = float(v.replace(',', ''))
    return v * {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1}.get(unit, 1)


out = []
for rep in sorted(glob.glob(os.path.join(ROOT, 'gpurun_out', 'prof_final_*.ncu-rep'))):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        out.append('=' * 100)
        out.append(os.path.basename(rep) + '   ' + r[idx['Kernel Name']])
        for w in WANT:
            if w in idx:
                out.append(f'  {w:80s} {r[idx[w]]:>18s} {units[idx[w]]}')
        rd = to_bytes(r[idx['dram__bytes_read.sum']], units[idx['dram__bytes_read.sum']])
        wr = to_bytes(r[idx['dram__bytes_write.sum']], units[idx['dram__bytes_write.sum']])
        t = to_sec(r[idx['gpu__time_duration.sum']], units[idx['gpu__time_duration.sum']])
        out.append(f'  -> DRAM traffic {(rd + wr) / 1e6:.3f} MB in {t * 1e6:.1f} us = {(rd + wr) / t / 1e9:.0f} GB/s '
                   f'({(rd + wr) / t / 6451.5e9 * 100:.1f} % of 6451.5 GB/s)')
        st = [(h, r[idx[h]]) for h in hdr if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
        st = sorted(st, key=lambda x: -float(x[1].replace(',', '') or 0))[:5]
        out.append('  top stalls: ' + ', '.join(
            h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '') + ' ' + v for h, v in st))
txt = '\n'.join(out) + '\n'
dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'profiles', 'r01_ncu_final_kernels.txt')
open(dst, 'w').write(txt)
print(txt)
