#!/usr/bin/env python
"""Summarise an ncu report exported with `--page raw --csv` and `--page source --csv`:
headline counters per launch, stall-reason ranking, opcode mix, hottest SASS lines."""
import collections
import csv
import sys

raw, src = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__grid_size', 'sm__cycles_elapsed.max',
        'launch__shared_mem_per_block_dynamic', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    print('--- launch', r[hdr.index('ID')], r[hdr.index('Kernel Name')][:60])
    for w in want:
        if w in hdr:
            print(f'  {w} = {r[hdr.index(w)]} {units[hdr.index(w)]}')
r = rows[2]
stall = [(h, r[hdr.index(h)]) for h in hdr if 'pcsamp_warps_issue_stalled' in h and not h.endswith('_not_issued')]
stall = [(h.replace('smsp__pcsamp_warps_issue_stalled_', ''), float(v.replace(',', ''))) for h, v in stall if v not in ('', 'n/a')]
tot = sum(v for _, v in stall) or 1
print('--- warp stall samples (first launch)')
for h, v in sorted(stall, key=lambda x: -x[1])[:10]:
    print(f'  {h}: {v:.0f} ({100 * v / tot:.1f}%)')

rows = list(csv.reader(open(src)))
hdr = rows[1]
si, sc, ie = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
recs = []
for r in rows[2:]:
    if r and r[0] == 'Kernel Name':
        break
    try:
        recs.append((int(r[si]), r[sc].strip(), int(r[ie])))
    except (ValueError, IndexError):
        pass
tot = sum(x[2] for x in recs)
ts = sum(x[0] for x in recs) or 1
print(f'--- SASS: {len(recs)} instructions, {tot} warp-level executions, {ts} samples')
op = collections.Counter()
for s, t, e in recs:
    o = t.split()
    name = o[1] if o[0].startswith('@') else o[0]
    op[name.split('.')[0]] += e
print('  opcode mix: ' + ', '.join(f'{k} {100 * v / tot:.1f}%' for k, v in op.most_common(14)))
print('--- hottest SASS lines by stall samples')
for i in sorted(sorted(range(len(recs)), key=lambda i: -recs[i][0])[:14]):
    print(f'  #{i} samples={recs[i][0]} ({100 * recs[i][0] / ts:.1f}%) exec={recs[i][2]}  {recs[i][1][:80]}')
