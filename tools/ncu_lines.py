#!/usr/bin/env python
"""Per-source-line instruction / stall-sample share of one kernel from an ncu report (needs -lineinfo + --import-source)."""
import csv, subprocess, sys, io
rep, pat = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.8
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
idx = [i for i, r in enumerate(rows) if r and r[0] == "Function Name"] + [len(rows)]
def I(x):
    try: return int(x)
    except Exception: return 0
alld = []
for k in range(len(idx) - 1):
    if pat not in rows[idx[k]][1]: continue
    hdr = rows[idx[k] + 1]
    ci, si, ti = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
    alld += [r for r in rows[idx[k] + 2:idx[k + 1]] if len(r) == len(hdr) and r[0].isdigit()]
tot = sum(I(r[ci]) for r in alld); ts = sum(I(r[si]) for r in alld)
print("warp instructions", tot, "samples", ts)
print(" line  inst%  smpl%  thr/inst  source")
for r in alld:
    p = 100 * I(r[ci]) / max(tot, 1); q = 100 * I(r[si]) / max(ts, 1)
    if p >= thr or q >= thr: print(r[0].rjust(5), f"{p:6.1f} {q:6.1f} {I(r[ti]) / max(I(r[ci]), 1):6.1f}  ", r[1][:110])
if len(sys.argv) > 4:   # extra args: line-range boundaries -> share per range
    b = [int(x) for x in sys.argv[4:]]
    acc = {}
    for r in alld:
        ln = int(r[0]); k = max([x for x in b if x <= ln], default=0)
        a = acc.setdefault(k, [0, 0]); a[0] += I(r[ci]); a[1] += I(r[si])
    for k in sorted(acc): print(f"from line {k:5d}: inst {100*acc[k][0]/tot:5.1f}%  samples {100*acc[k][1]/max(ts,1):5.1f}%")
