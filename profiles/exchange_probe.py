"""Breakdown of the multi-GPU step's exchange (fused into the EMA / table kernel): where the time between the end of
the fused pass and the end of the step goes.  Run under torchrun on N GPUs:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/exchange_probe.py
Time stamps (globaltimer, ns) of CTA 0 of table_kernel: start | handshake done (flags of all ranks seen) | gather done |
end, plus CUDA-event times of the three launches of a step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from onda_b200 import prototype_handler, _native as nat

world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
group = None
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    group = dist.group.WORLD
protos, sq, cnt, sets = bench.gpu_inputs(torch, dev, 256, 2, 1234 + rank)
if world > 1:
    for t in (protos, sq, cnt):
        dist.broadcast(t, 0)
h = prototype_handler(process_group=group, allreduce="oneshot", **bench.PARAMS)
h.prototypes, h.squared_mean, h.counter = protos, sq, cnt
lib = nat.load()
for i in range(6):
    f, p, o = sets[i % 2]
    h.pseudo_labels_fused(f, p, o); h.ma(f, o)
torch.cuda.synchronize()
buf = torch.zeros(64, dtype=torch.int64, device=dev)
lib.onda_debug_set_buffer(nat.ptr(buf))      # note: also selects the profiling build of the fused kernel
rows = []
for i in range(40):
    f, p, o = sets[i % 2]
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e[0].record()
    h.pseudo_labels_fused(f, p, o)
    e[1].record()
    h.ma(f, o)
    e[2].record()
    torch.cuda.synchronize()
    st = buf[:4].tolist()
    rows.append((e[0].elapsed_time(e[1]) * 1e3, e[1].elapsed_time(e[2]) * 1e3, (st[1] - st[0]) / 1e3, (st[2] - st[1]) / 1e3, (st[3] - st[2]) / 1e3))
lib.onda_debug_set_buffer(None)
r = torch.tensor(rows[8:]).median(dim=0)[0].tolist()
line = (f"rank {rank}/{world}: fused pass + combine {r[0]:7.1f} us | ma() launch {r[1]:6.1f} us, of which inside table_kernel: "
        f"handshake {r[2]:5.1f} us, gather {r[3]:5.1f} us, blend + table + done flags {r[4]:5.1f} us")
if world > 1:
    out = [None] * world
    dist.all_gather_object(out, line)
    if rank == 0:
        print("\n".join(out))
    dist.destroy_process_group()
else:
    print(line)
