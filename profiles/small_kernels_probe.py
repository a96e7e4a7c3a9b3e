"""Times the small per-step kernels and the host-side launch cost (run on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from onda_b200 import prototype_handler, _native as nat

dev = torch.device("cuda:0")
protos, sq, cnt, sets = bench.gpu_inputs(torch, dev, 256, 2, 1234)
h = prototype_handler(impl=sys.argv[1] if len(sys.argv) > 1 else "auto", **bench.PARAMS)
h.prototypes, h.squared_mean, h.counter = protos, sq, cnt
lib = nat.load()

def step(i):
    f, p, o = sets[i % 2]
    h.pseudo_labels_fused(f, p, o)
    h.ma(f, o)

for i in range(5):
    step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(50):
    step(i)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host launch time per step {(t1 - t0) / 50 * 1e6:.1f} us ; wall per step {(t2 - t0) / 50 * 1e6:.1f} us")

def timeit(fn, n=200):
    for _ in range(10):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

C, D = 19, 256
table = torch.zeros(lib.onda_table_floats(C, D), device=dev)
sums = torch.rand(lib.onda_sums_floats(C, D), device=dev) * 100
s = nat.C.c_void_p(torch.cuda.current_stream().cuda_stream)
P, S = protos.clone(), sq.clone()
print("table build      %.2f us" % timeit(lambda: lib.onda_build_distance_table(nat.ptr(P), nat.ptr(S), nat.ptr(cnt), C, D, 1, nat.ptr(table), s)))
print("ema + table      %.2f us" % timeit(lambda: lib.onda_ema_update_and_table(nat.ptr(P), nat.ptr(S), nat.ptr(cnt), nat.ptr(sums), C, D, 0.9995, 1, nat.ptr(table), s)))
print("ema only         %.2f us" % timeit(lambda: lib.onda_ema_update(nat.ptr(P), nat.ptr(S), nat.ptr(sums), C, D, 0.9995, s)))
print("empty torch op   %.2f us" % timeit(lambda: cnt.add_(0)))
