#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[2]
iL, iS, iN, iE = hdr.index("Line No"), 1, hdr.index("# Samples"), hdr.index("Instructions Executed")
recs = []
for r in rows[3:]:
    if len(r) > iE and r[iL].isdigit():
        try:
            recs.append((int(r[iL]), r[iS].strip(), int(r[iN]), int(r[iE])))
        except ValueError:
            pass
tot_e = sum(x[3] for x in recs) or 1
tot_s = sum(x[2] for x in recs) or 1
print(f"total warp-instr {tot_e}, samples {tot_s}")
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
print("--- by executed instructions")
for l, s, n, e in sorted(recs, key=lambda x: -x[3])[:top]:
    print(f"{l:5d} {100*e/tot_e:5.1f}% instr {100*n/tot_s:5.1f}% stall  {s[:110]}")
print("--- by stall samples")
for l, s, n, e in sorted(recs, key=lambda x: -x[2])[:25]:
    print(f"{l:5d} {100*e/tot_e:5.1f}% instr {100*n/tot_s:5.1f}% stall  {s[:110]}")
