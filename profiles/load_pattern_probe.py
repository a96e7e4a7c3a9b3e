"""Bandwidth of the fused pass's feature-load pattern alone (onda_debug_load_probe), for several L1 sizes
(shared-memory carve-outs) and worker-warp counts.  Run on the GPU box:  python profiles/load_pattern_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onda_b200 import _native as nat

dev = torch.device("cuda:0")
lib = nat.load()
B, D, H, W = 32, 256, 65, 129
HW = H * W
feat = torch.randn(B, D, H, W, device=dev)
out = torch.zeros(148 * 32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
nbytes = feat.numel() * 4
stream = torch.cuda.current_stream().cuda_stream
for workers in (16, 24, 32):
    for smem_kb in (0, 99, 131, 163, 195, 226):
        ts = []
        for it in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            nat.check(lib.onda_debug_load_probe(nat.ptr(feat), B, D, HW, workers, smem_kb * 1024, nat.ptr(out), stream))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts[2:])[len(ts[2:]) // 2]
        print(f"workers {workers:2d}  smem {smem_kb:3d} KB  {t * 1e3:7.1f} us  {nbytes / t / 1e6:7.0f} GB/s")
