#!/usr/bin/env python
"""Text summary of ncu captures:  python tools/ncu_summary.py <out.txt> <rep> [<rep> ...]
With --traffic n m first_row: also rewrites profiles/update_kernel_traffic.json from the FIRST report (one
launch of tile_dmma_kernel<UPDATE> covering local rows first_row .. n-1 of an n x n, block m problem on one
GPU): measured DRAM bytes against the algorithmic bytes of that launch (C read + write of the stored
trapezoid, 16 B per element, + the panel columns it reads, 8 m (n - first_row) B)."""
import csv
import io
import json
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
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum']


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(unit, 1)


def to_sec(v, unit):
    v = float(v.replace(',', ''))
    return v * {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1, 'usecond': 1e-6, 'msecond': 1e-3, 'nsecond': 1e-9,
                'second': 1}.get(unit, 1)


def main():
    args = sys.argv[1:]
    traffic = None
    if '--traffic' in args:
        i = args.index('--traffic')
        traffic = [int(x) for x in args[i + 1:i + 4]]
        del args[i:i + 4]
    dst, reps = args[0], args[1:]
    out, first = [], None
    for rep in reps:
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
            st = [(h, r[idx[h]]) for h in hdr
                  if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
            st = sorted(st, key=lambda x: -float(x[1].replace(',', '') or 0))[:5]
            out.append('  top stalls: ' + ', '.join(
                h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '') + ' ' + v
                for h, v in st))
            if first is None:
                dm = 'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active'
                first = {'dram': rd + wr, 't': t, 'dmma': float(r[idx[dm]].replace(',', '')) if dm in idx else None,
                         'grid': r[idx['launch__grid_size']], 'rep': os.path.basename(rep)}
    txt = '\n'.join(out) + '\n'
    open(dst, 'w').write(txt)
    print(txt)
    if traffic and first:
        n, m, row0 = traffic
        rows = n - row0
        elems = rows * (n - row0) - rows * (rows - 1) / 2.0
        # 64-row tiles start at the column tile of their first row: stored tiles, not only stored elements
        alg = 16.0 * elems + 8.0 * m * (n - row0)
        tj = {'kernel': 'tile_dmma_kernel<UPDATE>', 'n': n, 'm': m,
              'launch': f'bulk update of step 0 (local rows {row0}..{n - 1}, grid {first["grid"]}), look-ahead off',
              'dram_bytes_per_launch': first['dram'], 'algorithmic_bytes_per_launch': alg,
              'traffic_over_algorithmic': first['dram'] / alg, 'duration_ms_under_ncu': first['t'] * 1e3,
              'tflops_under_ncu': 2.0 * m * elems / first['t'] / 1e12, 'dmma_pipe_active_pct': first['dmma'],
              'source': 'profiles/' + os.path.basename(dst) + ' (ncu --set full --clock-control none, ' + first['rep'] + ')'}
        json.dump(tj, open(os.path.join(ROOT, 'profiles', 'update_kernel_traffic.json'), 'w'), indent=1)
        print(json.dumps(tj, indent=1))


if __name__ == '__main__':
    main()
