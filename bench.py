#!/usr/bin/env python
"""Benchmark of the prototype pseudo-labelling hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl onda|reference] [--d 256|2048]
                    [--scaling weak|strong] [--batch B]

One *step* = the fused pass over one batch of synthetic, Cityscapes-shaped input: hard
pseudo-labels + soft predictions + per-class feature sum / sum of squares / count + batch
confidence statistics, followed by the EMA prototype update -- what the reference does with
``pseudo_labels`` x2 + ``ma`` (prototypes_hybrid_switch.py:89-93, prototypes.py:292-294).

Workload (config.workload): BASELINE.json configs[2], the batch-sharded prototype path at
1024x512 (65x129 stride-8 map), 19 classes, mahalanobis, hybrid_switch.yml parameters.
``--scaling weak`` (default): B=32 images PER GPU, every rank keeps 32 images as N grows;
``--scaling strong``: B=32 images IN TOTAL, 32/N per rank (the split BASELINE configs[2] names).
The only collective is the exchange of the 19x(2D+1)+8 class-sum/statistics buffer before the
EMA update.

Printed (rank 0, one JSON line): ``value`` = pixels/s over all ranks with inputs resident in
HBM; ``e2e`` = the same through the public API from pinned HOST buffers (H2D of feat/prior/out
and D2H of labels + soft predictions + statistics inside the timed region); ``roofline`` for the
dominant kernel (algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json); ``cpu_baseline``
= the reference's own prototype_handler (oracle/_ref, staged by ``__graft_entry__.build()``; the
oracle port if that file is absent) timed on this box's host cores on the same batch.  The default single-GPU run
also reports ``other_configs``: the prototype path of BASELINE configs[0], [1] and [3] measured the same way, and
``aux_kernels`` (model-weight EMA, evaluation counters).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

C = 19
H, W = 65, 129            # stride-8 map of a 1024x512 image (H/8+1, W/8+1)
B_PER_GPU = 32
PARAMS = dict(ma_lambda=0.9995, tau=1, thresh=0.3, distance_metric="mahalanobis")  # configs/hybrid_switch.yml
METRIC = "prototype pseudo-label px/s"
UNIT = "px/s"


def algorithmic_bytes_per_pixel(d):
    """feat read once + prior read + EMA-logit read + soft write + int64 label."""
    return 4 * d + 4 * C + 4 * C + 4 * C + 8


def measured_traffic(d):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (None if it does not apply)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_tc_traffic.json")) as f:
            t = json.load(f)
        if d == 256:
            return t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        pass
    return None


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.15 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[1]) for r in rows]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons),
                "samples": len(rows), "power_w_max": max(float(r[3]) for r in rows)}


def gpu_inputs(torch, device, d, n_sets, seed, block=8, margin=4.0, shape=None):
    """Synthetic Cityscapes-shaped inputs generated on the device (SURVEY.md 8d recipe).  ``shape`` = (images, h, w)
    overrides the bench workload's geometry (the other BASELINE configs)."""
    B_PER_GPU, H, W = shape if shape is not None else (globals()["B_PER_GPU"], globals()["H"], globals()["W"])
    g = torch.Generator(device=device).manual_seed(seed)
    protos = torch.randn(C, d, generator=g, device=device) * 2.5
    sq_mean = protos ** 2 + torch.rand(C, d, generator=g, device=device) * 1.5 + 0.5
    counter = torch.floor(torch.rand(C, generator=g, device=device) * 6.9e4 + 1e3)
    sets = []
    for _ in range(n_sets):
        lab = torch.randint(0, C, (B_PER_GPU, (H + block - 1) // block, (W + block - 1) // block),
                            generator=g, device=device)
        lab = lab.repeat_interleave(block, 1).repeat_interleave(block, 2)[:, :H, :W]
        feat = torch.randn(B_PER_GPU, d, H, W, generator=g, device=device) * 2.5
        feat += 0.5 * protos[lab].permute(0, 3, 1, 2)
        keep = (torch.rand(B_PER_GPU, d, 1, 1, generator=g, device=device) >= 0.1).float() / 0.9
        feat *= keep                                   # Dropout2d pattern of the EMA model in train() mode
        hot = torch.nn.functional.one_hot(lab, C).permute(0, 3, 1, 2).float()
        out = torch.randn(B_PER_GPU, C, H, W, generator=g, device=device) * 3 + margin * hot
        prior = (torch.randn(B_PER_GPU, C, H, W, generator=g, device=device) * 3 + 4 * hot).softmax(1)
        sets.append((feat.contiguous(), prior.contiguous(), out.contiguous()))
    return protos, sq_mean, counter, sets


def other_config_numbers(torch, device, lib, steps=40):
    """The other BASELINE configs' prototype path on this GPU, measured like the headline (graph-replayed steps over
    rotating input sets larger than twice the L2, kernel time from CUDA events in a short eager loop): parity of these
    shapes is tests/test_gpu_parity.py (test_full_size_oracle_parity, test_oracle_parity_shapes)."""
    from onda_b200 import prototype_handler
    from onda_b200 import _native as nat
    peak, _ = hbm_peak()
    out = {}
    for name, (b, d, hh, ww) in (("configs[0] B=1 D=2048 65x129", (1, 2048, 65, 129)),
                                 ("configs[1] prototype part B=1 D=256 65x129", (1, 256, 65, 129)),
                                 ("configs[3] B=8 D=256 129x257", (8, 256, 129, 257))):
        n = b * hh * ww
        set_bytes = n * algorithmic_bytes_per_pixel(d)
        n_sets = int(max(2, min(32, -(-2.2 * 126e6 // set_bytes))))
        protos, sq_mean, counter, sets = gpu_inputs(torch, device, d, n_sets, 4321, shape=(b, hh, ww))
        h = prototype_handler(**PARAMS)
        h.prototypes, h.squared_mean, h.counter = protos, sq_mean, counter

        def eager(i):
            feat, prior, logits = sets[i % n_sets]
            h.pseudo_labels_fused(feat, prior, logits)
            h.ma(feat, logits)
        for i in range(3):
            eager(i)
        lib.onda_kernel_timing_enable(1)
        for i in range(8):
            eager(i)
        torch.cuda.synchronize()
        tot_ms, n_timed = nat.C.c_float(0), nat.C.c_int(0)
        nat.check(lib.onda_kernel_timing_read(nat.C.byref(tot_ms), nat.C.byref(n_timed)))
        lib.onda_kernel_timing_enable(0)
        graphs = []
        for k in range(n_sets):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                eager(k)
            graphs.append(g)
        for i in range(5):
            graphs[i % n_sets].replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(steps):
            graphs[i % n_sets].replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        kernel_ms = tot_ms.value / max(n_timed.value, 1)      # D > 256: the tcgen05 kernel alone (its finishing launch is in the step)
        out[name] = {"ms_per_step": ms, "px_per_s": n / (ms * 1e-3), "kernel": h.impl, "kernel_ms": kernel_ms,
                     "roofline_frac_kernel": set_bytes / (kernel_ms * 1e-3) / 1e9 / peak if kernel_ms > 0 else None,
                     "roofline_frac_step": set_bytes / (ms * 1e-3) / 1e9 / peak, "input_sets": n_sets, "steps": steps}
        del graphs, sets, h
        torch.cuda.empty_cache()
    return out


def cpu_reference_rate(torch, d, images, steps, warmup, threads):
    """Times the reference path (pseudo_labels hard + soft + ma) on host cores: the reference's own prototype_handler
    when ``oracle/_ref`` holds it (kind "reference"), else the oracle port (kind "port").  Returns (pixels per step,
    list of step times, kind)."""
    from oracle import proto_oracle as po
    from oracle.build_ref import load_reference_handler
    torch.set_num_threads(threads)
    case = po.synth_case(1234, images, d, H, W)
    ref_cls = load_reference_handler()
    if ref_cls is not None:
        h = ref_cls(ma_lambda=PARAMS["ma_lambda"], tau=PARAMS["tau"], thresh=PARAMS["thresh"],
                    distance_metric=PARAMS["distance_metric"])
        kind = "reference"

        def one_step():
            h.pseudo_labels(case["feat"], case["prior"])
            h.pseudo_labels(case["feat"], case["prior"], soft=True)
            h.ma(case["feat"], case["out"])
    else:
        h = po.OracleHandler(**PARAMS)
        kind = "port"

        def one_step():
            po.fused_step(h, case["feat"], case["prior"], case["out"])
    h.prototypes, h.squared_mean, h.counter = case["protos"].clone(), case["sq_mean"].clone(), case["counter"].clone()
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            one_step()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return images * H * W, times, kind


def per_rank_images(args, world):
    if args.scaling == "strong":
        if args.batch % world:
            raise SystemExit(f"--scaling strong needs --batch ({args.batch}) divisible by the number of GPUs ({world})")
        return args.batch // world
    return args.batch


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of one GPU's share of the step, all host threads."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    images = per_rank_images(args, args.gpus)
    steps = max(1, min(args.steps, 8 if args.d <= 256 else 2))        # a full batch is ~1.4 s of CPU at D = 256
    px, times, kind = cpu_reference_rate(torch, args.d, images, steps, 1, threads)
    dt = sum(times) / len(times)
    value = px / dt
    what = "the reference's prototype_handler (oracle/_ref)" if kind == "reference" else "oracle port of the reference torch path"
    sample = (f"one rank's full batch ({images} images, {px} px) per step, {steps} timed steps after 1 warm-up, {what}, "
              f"{threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.d, args.gpus, args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(d, n_gpus, args):
    """Identical for both arms (the driver compares them)."""
    b = per_rank_images(args, n_gpus)
    split = (f"B={args.batch} images per GPU (weak scaling)" if args.scaling == "weak"
             else f"B={args.batch} images in total, {b} per GPU (strong scaling)")
    return {"workload": (f"BASELINE configs[2]: batch-sharded prototype path, {split} at 1024x512 "
                         f"(65x129 stride-8 map), D={d}, C={C}, mahalanobis, hybrid_switch.yml parameters; step = fused "
                         "hard+soft pseudo-labels + class sum/sumsq/count + statistics + EMA update"),
            "B_per_gpu": b, "D": d, "H": H, "W": W, "classes": C, "parallelism": f"batch-sharded x{n_gpus}",
            "collective": "exchange of 19x(2D+1)+8 floats per step" if n_gpus > 1 else "none",
            "l2_policy": "rotating input sets, together larger than twice the 126 MB L2 (two sets of 338 MB at B=32, D=256)"}


def aux_kernel_numbers(torch, device):
    """The widened rows, measured in the same run: model-weight EMA (one launch over a DeepLabV2-sized parameter set)
    and the evaluation counters (upsample + argmax + confusion matrix of one 2048x1024 image)."""
    from onda_b200 import WeightEma, ConfusionMeter
    out = {}
    try:
        g = torch.Generator(device=device).manual_seed(3)
        shapes = [(64, 3, 7, 7)] + [(256, 256, 3, 3)] * 60 + [(512,)] * 100 + [(2048, 512, 1, 1)] * 4
        net = lambda: torch.nn.ParameterList([torch.nn.Parameter(torch.randn(*sh, generator=g, device=device)) for sh in shapes])
        q, k = net(), net()
        plan = WeightEma(q, k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            plan.update(0.999)
        e0.record()
        for _ in range(10):
            plan.update(0.999)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        n = sum(p.numel() for p in q)
        out["weight_ema"] = {"params": n, "ms": ms, "gbs": 12 * n / (ms * 1e-3) / 1e9, "bytes_per_param": 12}
        meter = ConfusionMeter(C)
        pred = torch.randn(1, C, 129, 257, generator=g, device=device)
        lab = torch.randint(0, C, (1, 1024, 2048), generator=g, device=device)
        for _ in range(3):
            meter.update(pred, lab)
        e0.record()
        for _ in range(10):
            meter.update(pred, lab)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        out["confusion"] = {"full_res_pixels": 1024 * 2048, "ms": ms, "gpx_s": 1024 * 2048 / (ms * 1e-3) / 1e9}
    except Exception as exc:                      # never let the side numbers break the bench line
        out["error"] = repr(exc)
    return out


def run_onda(args):
    import torch
    import torch.distributed as dist
    from onda_b200 import prototype_handler, Monitor
    from onda_b200 import _native as nat

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        group = dist.group.WORLD
    d = args.d
    globals()["B_PER_GPU"] = per_rank_images(args, world)
    N = B_PER_GPU * H * W
    lib = nat.load()

    # rotating input sets: together larger than twice the 126 MB L2, so no step finds its inputs cached
    set_bytes = N * algorithmic_bytes_per_pixel(d)
    n_sets = int(max(2, min(16, -(-2.2 * 126e6 // set_bytes))))
    protos, sq_mean, counter, sets = gpu_inputs(torch, device, d, n_sets, 1234 + rank, block=args.label_block, margin=args.logit_margin)
    if world > 1:   # identical prototypes everywhere
        for t in (protos, sq_mean, counter):
            dist.broadcast(t, 0)
    h = prototype_handler(process_group=group, impl=args.kernel, allreduce=args.allreduce, tile_schedule=args.tile_schedule, **PARAMS)
    h.prototypes, h.squared_mean, h.counter = protos.clone(), sq_mean.clone(), counter.clone()

    def step(i):
        feat, prior, out = sets[i % len(sets)]
        labels, soft = h.pseudo_labels_fused(feat, prior, out)
        h.ma(feat, out)
        return labels, soft

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ---------------------------------------------------
    for i in range(args.warmup):
        step(i)
    barrier()
    # Kernel time first, in a short eager loop: CUDA events recorded by the library around every 4th launch of the
    # dominant kernel (not possible inside a captured graph).
    lib.onda_kernel_timing_enable(4)
    for i in range(max(8, min(args.steps, 24))):
        step(i)
    torch.cuda.synchronize()
    tot_ms, n_timed = nat.C.c_float(0), nat.C.c_int(0)
    nat.check(lib.onda_kernel_timing_read(nat.C.byref(tot_ms), nat.C.byref(n_timed)))
    lib.onda_kernel_timing_enable(0)
    # The timed steps replay CUDA graphs (one per rotating input set; SURVEY 8d: "graph-replayed steps") on one GPU:
    # the step is three launches, and at small batch the host cannot enqueue them as fast as the GPU runs them.
    # Multi-GPU: the exchange fused into ma() keeps its epoch in device memory, so those steps replay too (the two
    # graphs alternate strictly, like the two peer-visible slots); the NCCL variant stays eager.
    eager_step, graphs, launches_per_step = step, None, None
    if args.graph and (world == 1 or args.allreduce == "oneshot"):
        try:
            l0 = lib.onda_launch_count()
            graphs = []
            for k in range(len(sets)):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    eager_step(k)
                graphs.append(g)
            launches_per_step = (lib.onda_launch_count() - l0) // len(sets)

            def step(i):
                graphs[i % len(graphs)].replay()
            for i in range(args.warmup):
                step(i)
        except Exception as exc:           # capture unavailable: measure the eager loop and say so
            print(f"[bench] CUDA graph capture failed ({exc!r}); timing the eager loop", file=sys.stderr)
            graphs, step = None, eager_step
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = lib.onda_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record()
    t_host0 = time.perf_counter()
    for i in range(args.steps):
        step(i)
    host_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps      # host time to enqueue one step (no sync inside)
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = lib.onda_launch_count() - launches0 if graphs is None else launches_per_step * args.steps
    step = eager_step
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    t = torch.tensor([ms], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * N * args.steps / (ms * 1e-3)

    # ---- end to end from pinned host buffers ------------------------------------------------
    # ---- multi-GPU correctness, visible to the driver: prototypes bit-identical on every rank, and the sharded result
    # equal to one GPU running the concatenated shards (fp32 summation order apart)
    xrank = None
    if world > 1:
        def digest(t):
            b = t.detach().contiguous().reshape(-1).view(torch.int32).to(torch.int64)
            return (b * (torch.arange(b.numel(), device=b.device) % 1000003 + 1)).sum().reshape(1)
        mine = torch.cat([digest(h.prototypes), digest(h.squared_mean)])
        allh = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allh, mine)
        identical = all(torch.equal(a, allh[0]) for a in allh)
        # two more steps from a common state, sharded, vs rank 0 alone on the gathered batch
        start = [t.clone() for t in (h.prototypes, h.squared_mean, h.counter)]
        for i in range(2):
            eager_step(i)
        torch.cuda.synchronize()
        gathered = []
        for i in range(2):
            parts = []
            for x in sets[i]:
                buf = [torch.empty_like(x) for _ in range(world)]
                dist.all_gather(buf, x)
                parts.append(torch.cat(buf))
            gathered.append(parts)
        relerr = None
        if rank == 0:
            h1 = prototype_handler(impl=args.kernel, tile_schedule=args.tile_schedule, **PARAMS)
            h1.prototypes, h1.squared_mean, h1.counter = (t.clone() for t in start)
            for i in range(2):
                feat, prior, out = gathered[i]
                h1.pseudo_labels_fused(feat, prior, out)
                h1.ma(feat, out)
            relerr = max(float((h1.prototypes - h.prototypes).abs().max() / h.prototypes.abs().max()),
                         float((h1.squared_mean - h.squared_mean).abs().max() / h.squared_mean.abs().max()))
        del gathered
        xrank = {"xrank_identical": bool(identical), "vs_single_gpu_relerr": relerr}
        barrier()

    host = [tuple(x.cpu().pin_memory() for x in s) for s in sets[:2]]
    dev_in = [tuple(torch.empty_like(x) for x in s) for s in sets[:2]]
    lab_host = torch.empty((N, 1), dtype=torch.int64).pin_memory()
    soft_host = torch.empty((N, C), dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(device)
    mon = Monitor(200, 0.003, "hamming")
    h2d_bytes = sum(x.numel() * x.element_size() for x in host[0])
    d2h_bytes = lab_host.numel() * 8 + soft_host.numel() * 4 + nat.NUM_STATS * 4

    def upload(i, after=None):
        with torch.cuda.stream(copy_stream):
            if after is not None:
                copy_stream.wait_event(after)          # the buffer set's previous consumer has finished
            for dst, src in zip(dev_in[i % 2], host[i % 2]):
                dst.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    def e2e_loop(k):
        ev = upload(0)
        prev_done = None
        for i in range(k):
            torch.cuda.current_stream().wait_event(ev)
            feat, prior, out = dev_in[i % 2]
            if i + 1 < k:
                ev = upload(i + 1, prev_done)          # overlaps with this step's compute (other buffer set)
            labels, soft = h.pseudo_labels_fused(feat, prior, out, confidence_monitor=mon)   # reads the stats (D2H)
            h.ma(feat, out)
            lab_host.copy_(labels, non_blocking=True)
            soft_host.copy_(soft, non_blocking=True)
            prev_done = torch.cuda.Event()
            prev_done.record()
        torch.cuda.synchronize()

    e2e_steps = max(3, min(args.steps, 10))
    e2e_loop(2)
    barrier()
    t0 = time.perf_counter()
    e2e_loop(e2e_steps)
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * N * e2e_steps / float(t.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = hbm_peak()
    kernel_ms = tot_ms.value / max(n_timed.value, 1)
    alg_bytes = N * algorithmic_bytes_per_pixel(d)
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "host_enqueue_ms_per_step": host_ms, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(d, world, args),
        "launch_mode": "CUDA graph replay (one graph per input set)" if graphs is not None else "eager launches",
        "tile_schedule": args.tile_schedule,
        "programmatic_dependent_launch": os.environ.get("ONDA_PDL", "1") != "0",
        "input_sets": len(sets),
        "exchange": (None if world == 1 else
                     {"requested": args.allreduce, "effective": h.allreduce,
                      "peer_memory": bool(h._symm is not None), "fused_into_ema_kernel": bool(h._symm is not None and h.allreduce == "oneshot")}),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "steps": e2e_steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": measured_traffic(d) if args.kernel in ("auto", "tcgen05") else None,
                     "traffic_source": "profiles/r2_tc_traffic.json (ncu --set full capture of this command, dram__bytes_read.sum + dram__bytes_write.sum per launch)",
                     "kernel": f"fused pseudo-label pass ({h.impl})", "kernel_ms": kernel_ms,
                     "launches_timed": int(n_timed.value), "bytes_per_px": algorithmic_bytes_per_pixel(d),
                     "peak_source": peak_src, "step_frac": alg_bytes / (ms / args.steps * 1e-3) / 1e9 / peak},
    }
    if xrank is not None:
        line.update(xrank)
    if world == 1:
        threads = os.cpu_count() or 1
        reps = 3 if d <= 256 else 1
        px, times, kind = cpu_reference_rate(torch, d, B_PER_GPU, reps, 1, threads)
        best = min(times)
        what = "the reference's prototype_handler (oracle/_ref)" if kind == "reference" else "oracle port of the reference torch path"
        line["cpu_baseline"] = {"value": px / best, "unit": UNIT, "cores": threads, "kind": kind,
                                "sample": f"the full batch ({B_PER_GPU} images, {px} px), best of {reps} after 1 warm-up, {what}"}
        line["aux_kernels"] = aux_kernel_numbers(torch, device)
        if d == 256 and B_PER_GPU == 32 and args.kernel == "auto":      # the default run also reports the other configs
            line["other_configs"] = other_config_numbers(torch, device, lib)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="onda", choices=["onda", "reference"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--allreduce", default="oneshot", choices=["nccl", "oneshot"],
                    help="exchange of the class-sum buffer at N>1: NCCL all_reduce or the library's one-shot NVLink kernel")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="time eager launches instead of CUDA-graph replays (single-GPU runs replay graphs by default)")
    ap.add_argument("--tile-schedule", default="dynamic", choices=["dynamic", "fixed"],
                    help="tcgen05 kernel: tiles drawn from a device counter (default) or the fixed, bit-reproducible round-robin")
    ap.add_argument("--batch", type=int, default=32, help="images per GPU (weak) or in total (strong); default 32 = the bench workload")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch images on every GPU; strong: --batch images in total, split over the GPUs")
    ap.add_argument("--label-block", type=int, default=8, help="side of the constant-label blocks of the synthetic maps")
    ap.add_argument("--logit-margin", type=float, default=4.0, help="logit bonus of the block's label (coherence of the argmax)")
    ap.add_argument("--d", "--dim", dest="d", type=int, default=256, help="feature width (256 = real ProDA head, 2048 = stress size; --dim is the spelling to use under torchrun, whose own parser claims --d)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    globals()["B_PER_GPU"] = args.batch
    if args.impl == "reference":
        run_reference(args)
    else:
        run_onda(args)


if __name__ == "__main__":
    main()
