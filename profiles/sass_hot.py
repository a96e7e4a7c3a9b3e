#!/usr/bin/env python
"""Hottest SASS regions of a captured kernel: reads `ncu -i REPORT --page source --csv --print-source sass` and prints,
per contiguous address window, the stall samples, executed instructions and dominant stall reasons.
Usage: python profiles/sass_hot.py REPORT [window]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
win = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(h)]
ci = {n: h.index(n) for n in h}
S = ci["Warp Stall Sampling (All Samples)"]; E = ci["Instructions Executed"]
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot_s = sum(float(r[S] or 0) for r in body); tot_e = sum(float(r[E] or 0) for r in body)
print(f"instructions {len(body)}  samples {tot_s:.0f}  executed {tot_e:.0f}")
for a in range(0, len(body), win):
    blk = body[a:a + win]
    s = sum(float(r[S] or 0) for r in blk); e = sum(float(r[E] or 0) for r in blk)
    if s < 0.01 * tot_s and e < 0.01 * tot_e:
        continue
    reasons = {n: sum(float(r[ci[n]] or 0) for r in blk) for n in stalls}
    top = sorted(reasons.items(), key=lambda kv: -kv[1])[:4]
    ops = {}
    for r in blk:
        op = r[ci["Source"]].split()[0] if r[ci["Source"]] else "?"
        if op.startswith("@"):
            op = r[ci["Source"]].split()[1]
        ops[op] = ops.get(op, 0) + 1
    topops = sorted(ops.items(), key=lambda kv: -kv[1])[:5]
    print(f"[{a:5d}] samples {100 * s / tot_s:5.1f}%  exec {100 * e / tot_e:5.1f}%  " + " ".join(f"{k[6:]}={v:.0f}" for k, v in top) + "   " + " ".join(f"{k}x{v}" for k, v in topops))
