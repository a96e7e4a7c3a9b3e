#!/usr/bin/env python
"""Per-role instruction / stall-sample split of fused_tc_kernel from an .ncu-rep (round-2 kernel).
The role branches appear in the SASS in source order and each begins with its setmaxnreg (USETMAXREG): prologue +
converters' entry | converters | summers | epilogue, sorter, MMA issuer, producer and teardown (the last warpgroup keeps
its registers, so that block is split further by execution count: epilogue lines run once per tile and warp).
Usage: python profiles/sass_roles_r2.py REPORT TILES"""
import csv, io, subprocess, sys

rep, tiles = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
body = [r for r in rows[hi + 1:] if len(r) == len(h)]
ci = {n: h.index(n) for n in h}
E, S, SRC = ci["Instructions Executed"], ci["Warp Stall Sampling (All Samples)"], ci["Source"]
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
tot_e = sum(float(r[E] or 0) for r in body)
tot_s = sum(float(r[S] or 0) for r in body)
marks = [i for i, r in enumerate(body) if "USETMAXREG" in r[SRC]]
assert len(marks) == 3, marks
names = ["prologue", "converters (warps 0-7)", "summers (warps 8-23)", "epilogue + sorter + MMA issuer + producer + teardown"]
bounds = [0] + marks + [len(body)]
print(f"{len(body)} SASS instructions, {tot_e / 1e6:.2f} M warp instructions executed, {tot_s:.0f} stall samples")
def show(name, blk):
    e = sum(float(r[E] or 0) for r in blk)
    s = sum(float(r[S] or 0) for r in blk)
    d = {n: sum(float(r[ci[n]] or 0) for r in blk) for n in stalls}
    top = sorted(d.items(), key=lambda kv: -kv[1])[:6]
    poll = sum(float(r[E] or 0) for r in blk if "SYNCS.PHASECHK" in r[SRC] or "NANOSLEEP" in r[SRC])
    print(f"{name:55s} instr {100 * e / tot_e:5.1f}%  samples {100 * s / tot_s:5.1f}%  (polling {100 * poll / tot_e:4.1f}%)   "
          + " ".join(f"{k[6:]}={100 * v / max(s, 1):.0f}%" for k, v in top))
for name, a, b in zip(names, bounds[:-1], bounds[1:]):
    show(name, body[a:b])
last = body[marks[2]:]
epi = [r for r in last if abs(float(r[E] or 0) - 4 * tiles) <= 0.03 * 4 * tiles]
show("  of which per-tile epilogue lines (x 4 warps x tiles)", epi)
ops = {}
for r in body:
    s = r[SRC].split()
    if not s:
        continue
    op = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
    ops[op] = ops.get(op, 0) + float(r[E] or 0)
print("top opcodes by executed count:", ", ".join(f"{k} {100 * v / tot_e:.1f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:14]))
