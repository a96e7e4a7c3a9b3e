#!/usr/bin/env python
"""Per-role instruction / stall-sample breakdown of the tcgen05 kernel from an .ncu-rep.
Roles are recognised by how often an instruction executed (loader lines run once per warp-chunk,
epilogue lines once per warp-tile, ...).  Usage: python profiles/sass_roles.py REPORT TILES CHUNKS_PER_TILE"""
import collections
import csv
import io
import subprocess
import sys

rep, tiles, nb = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
body = [r for r in rows[2:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
loader, epi, chunk1 = tiles * nb * 4, tiles * 4, tiles * nb
cls, inst = collections.Counter(), collections.Counter()
for r in body:
    n = float(r[ix["Instructions Executed"]] or 0)
    s = float(r[ix["# Samples"]] or 0)
    src = r[ix["Source"]]
    def near(a, b):
        return abs(a - b) <= 0.02 * b
    if "NANOSLEEP" in src or "SYNCS.PHASECHK" in src:
        c = "barrier waits"
    elif n == 0:
        c = "not executed"
    elif near(n, loader) or near(n, 2 * loader):
        c = f"loader lines (x{loader})"
    elif near(n, epi) or near(n, epi * 19) or near(n, epi * 20):
        c = f"epilogue lines (x{epi})"
    elif near(n, chunk1) or near(n, chunk1 * 32) or near(n, chunk1 * 8) or near(n, chunk1 * 16):
        c = f"class-sum/MMA per-chunk lines (x{chunk1})"
    elif n > loader * 2:
        c = "high-count lines (summation inner paths / spins)"
    elif n < epi:
        c = "rare"
    else:
        c = "other"
    cls[c] += s
    inst[c] += n
tot_s, tot_i = sum(cls.values()), sum(inst.values())
print(f"total warp instructions {tot_i/1e6:.1f}M, samples {tot_s:.0f}")
for c, s in cls.most_common():
    print(f"  {c:52s} samples {100*s/tot_s:5.1f}%   inst {inst[c]/1e6:8.2f}M ({100*inst[c]/tot_i:4.1f}%)")
