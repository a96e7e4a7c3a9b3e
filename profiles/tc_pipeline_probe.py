"""Reads the tcgen05 kernel's per-warp wait counters (onda_debug_set_buffer) for the bench workload.
Run on the GPU box:  python profiles/tc_pipeline_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from onda_b200 import prototype_handler, _native as nat

dev = torch.device("cuda:0")
if len(sys.argv) > 1:
    bench.B_PER_GPU = int(sys.argv[1])          # images (default: the bench workload's 32)
protos, sq, cnt, sets = bench.gpu_inputs(torch, dev, 256, 1, 1234)
h = prototype_handler(impl="tcgen05", **bench.PARAMS)
h.prototypes, h.squared_mean, h.counter = protos, sq, cnt
lib = nat.load()
feat, prior, out = sets[0]
for _ in range(3):
    h.pseudo_labels_fused(feat, prior, out); h.ma(feat, out)
buf = torch.zeros(148 * 32 * 8, dtype=torch.int64, device=dev)
lib.onda_debug_set_buffer(nat.ptr(buf))
h.pseudo_labels_fused(feat, prior, out)
torch.cuda.synchronize()
lib.onda_debug_set_buffer(None)
d = buf.view(148, 32, 8).double().cpu()
tot = d[:, :, 7]
names = {"converter": (0, 8, ["wait ring_full", "wait acc_empty", "wait empty_a(mma)", "convert + tcgen05.st + wait::st"]),
         "summer": (8, 24, ["wait ring_full", "wait sort_ready", "-", "class-sum loop"]),
         "epilogue": (24, 28, ["wait acc_full"]), "sorter": (28, 30, ["wait sort_free"]),
         "mma": (30, 31, ["wait acc_empty", "wait full_a"]), "producer": (31, 32, ["wait ring_empty"])}
print("mean total cycles per warp:", tot[:, :32].mean().item(), " max over CTAs:", tot[:, :32].mean(dim=1).max().item())
print("prologue cycles (mean):", d[:, :, 4].mean().item(), " wait-for-slowest-role:", d[:, :, 5].mean().item(), " teardown:", d[:, :, 6].mean().item())
for role, (a, b, labels) in names.items():
    t = tot[:, a:b].mean().item()
    print(f"{role:9s} total {t:10.0f} cyc   (prologue {d[:, a:b, 4].mean().item():.0f}, waiting for the slowest role {d[:, a:b, 5].mean().item():.0f}, teardown mean {d[:, a:b, 6].mean().item():.0f} max {d[:, a:b, 6].max().item():.0f})")
    for i, lab in enumerate(labels):
        v = d[:, a:b, i].mean().item()
        print(f"     {lab:24s} {v:10.0f} cyc  {100 * v / t:5.1f}%")
# per-CTA spread: CTAs 0..(tiles % grid - 1) carry one tile more than the rest
per_cta = tot[:, 24:28].mean(dim=1)          # epilogue warps: the last role to finish
import math
tiles = bench.B_PER_GPU * math.ceil(65 * 129 / 128)
extra = tiles % 148
if tiles < 148:
    print(f"{tiles} tiles on {tiles} CTAs: epilogue role cycles mean {per_cta[:tiles].mean().item():.0f} max {per_cta[:tiles].max().item():.0f}")
    sys.exit(0)
a, b = per_cta[:extra], per_cta[extra:]
print(f"CTAs with {tiles // 148 + 1} tiles: n={len(a)} mean {a.mean().item():.0f} min {a.min().item():.0f} max {a.max().item():.0f}")
print(f"CTAs with {tiles // 148} tiles: n={len(b)} mean {b.mean().item():.0f} min {b.min().item():.0f} max {b.max().item():.0f}")
print("per-CTA cycles / tile, sorted deciles:", [round(x) for x in torch.quantile(torch.cat([a / (tiles // 148 + 1), b / (tiles // 148)]), torch.linspace(0, 1, 11).double()).tolist()])
