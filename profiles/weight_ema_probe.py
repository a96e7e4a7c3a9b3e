"""Model-weight EMA: one launch of onda_weight_ema_update vs the reference's per-parameter loop, on a
DeepLabV2-ResNet50-sized parameter list (about 59 M float32 parameters in ~320 tensors).
Run on the GPU box:  python profiles/weight_ema_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from onda_b200 import WeightEma

dev = torch.device("cuda:0")


class Bag(torch.nn.Module):
    def __init__(self):
        super().__init__()
        sizes = [64 * 3 * 49, 64]
        for cin, cmid, n in ((64, 64, 3), (256, 128, 4), (512, 256, 6), (1024, 512, 3)):   # bottleneck stacks
            for i in range(n):
                c0 = cin if i == 0 else cmid * 4
                sizes += [c0 * cmid, cmid, cmid * cmid * 9, cmid, cmid * cmid * 4, cmid * 4]
                if i == 0:
                    sizes += [c0 * cmid * 4, cmid * 4]
        sizes += [2048 * 256 * 9] * 4 + [256] * 4 + [256 * 19] * 4 + [19] * 4                 # ASPP-like head
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(s)) for s in sizes])
        for i in range(106):                                                               # BatchNorm buffers
            self.register_buffer(f"rm{i}", torch.randn(256))
            self.register_buffer(f"rv{i}", torch.rand(256))
            self.register_buffer(f"nb{i}", torch.tensor(7))


q, k = Bag().to(dev), Bag().to(dev)
plan = WeightEma(q, k)
nbytes = 3 * plan.param_bytes + 2 * plan.buffer_bytes
print(f"tensors {len(list(q.parameters()))} params + {len(list(q.buffers()))} buffers, {plan.param_bytes / 4e6:.1f} M parameters, {plan.n_chunks} chunks")


def reference_loop(a):
    for pq, pk in zip(q.parameters(), k.parameters()):
        pk.data = pk.data.clone() * a + pq.data.clone() * (1.0 - a)
    for bq, bk in zip(q.buffers(), k.buffers()):
        bk.data = bq.data.clone()


def timed(fn, n):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

t_one = timed(lambda: plan.update(0.999), 20)
print(f"onda_weight_ema_update: {t_one * 1e3:8.1f} us per update, {nbytes / t_one / 1e6:7.0f} GB/s (read 2x + write 1x parameters = {nbytes / 1e6:.0f} MB per update, larger than L2)")
t_ref = timed(lambda: reference_loop(0.999), 5)
print(f"reference loop (torch, same GPU): {t_ref * 1e3:8.1f} us per update  ->  {t_ref / t_one:.1f}x")
