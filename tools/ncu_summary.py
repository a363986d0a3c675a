#!/usr/bin/env python3
"""Summarise an .ncu-rep: per-kernel headline metrics (+ optionally the hottest SASS of one kernel).

    python tools/ncu_summary.py REPORT.ncu-rep [--source KERNEL_REGEX] [--top N]
"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed_op_shared_atom.sum']
STALLS = ['short_scoreboard', 'long_scoreboard', 'barrier', 'wait', 'branch_resolving', 'mio_throttle', 'lg_throttle',
          'math_pipe_throttle', 'not_selected', 'no_instruction', 'dispatch_stall', 'sleeping', 'membar', 'drain', 'imc_miss']


def run(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    src = sys.argv[sys.argv.index('--source') + 1] if '--source' in sys.argv else None
    top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 40
    rows = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('----', r[hdr.index('Kernel Name')][:100])
        for w in WANT:
            if w in hdr:
                print(f'  {w:75s} {r[hdr.index(w)][:40]} {units[hdr.index(w)]}')
        st = []
        for s in STALLS:
            k = f'smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio'
            if k in hdr:
                st.append((float(r[hdr.index(k)] or 0), s))
        print('  stalls/issue:', ', '.join(f'{s} {v:.2f}' for v, s in sorted(st, reverse=True)[:7]))
    if src:
        rows = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + src]))))
        hdr = rows[1]
        ia, isrc, iav, ist = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('Avg. Threads Executed'), hdr.index('# Samples')
        data = [(int(r[ia]), r[isrc], r[iav], int(r[ist])) for r in rows[2:] if len(r) > ia and r[ia].isdigit()]
        tot, ts = sum(d[0] for d in data), sum(d[3] for d in data)
        print(f'source: {len(data)} SASS, {tot} warp-instr, {ts} samples')
        print('regions of 40 SASS with > 1.5 % of instructions or samples:')
        for s in range(0, len(data), 40):
            c, t = sum(d[0] for d in data[s:s + 40]), sum(d[3] for d in data[s:s + 40])
            if c > tot * 0.015 or t > ts * 0.015:
                print(f'  [{s:5d}] {c / 1e6:8.1f} M instr {100 * c / tot:5.1f}%   samples {100 * t / ts:5.1f}%')
        print(f'top {top} SASS by samples:')
        for i in sorted(range(len(data)), key=lambda i: -data[i][3])[:top]:
            d = data[i]
            print(f'  [{i:5d}] exec {d[0] / 1e3:9.0f}k thr {d[2]:>3s} samples {d[3]:6d}  {d[1][:90]}')


if __name__ == '__main__':
    main()
