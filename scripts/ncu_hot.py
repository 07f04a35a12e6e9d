#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep: headline raw metrics + the hottest stall sites (address order).
usage: ncu_hot.py REP KERNEL_REGEX [min_pct]"""
import csv, subprocess, sys, io
rep, kre = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.8
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv', '--kernel-name', 'regex:' + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'sm__cycles_active.avg', 'sm__cycles_elapsed.max',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']
for r in rows[2:]:
    print('---', r[hdr.index('Kernel Name')][:90])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w[:62]:62s} {r[i][:40]} {units[i]}")
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
i_src, i_s, i_ex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stalls = [(j, h) for j, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
seen, order = set(), []
for r in rows[2:]:
    if len(r) <= max(i_s, i_ex) or r[0] in seen or r[0] == 'Address':
        continue
    seen.add(r[0])
    try:
        int(r[i_s])
    except ValueError:
        continue
    order.append(r)
tot = sum(int(r[i_s]) for r in order)
print('total samples', tot, 'instructions', len(order))
for r in order:
    n = int(r[i_s])
    if 100.0 * n / tot > minp:
        st = sorted([(int(r[j]), h[6:]) for j, h in stalls if r[j] not in ('', '0')], reverse=True)[:2]
        print(f"{n:6d} {100*n/tot:5.1f}% {r[0][-5:]} ex={r[i_ex]:>8s} {r[i_src].strip()[:64]:64s} {st}")
