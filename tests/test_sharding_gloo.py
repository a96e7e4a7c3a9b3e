"""World-size-2 gloo test (CPU) of the multi-GPU host logic: batch sharding, the single all-reduce of the
`sums` buffer, identical prototypes on every rank and identical switch decisions from the replicated Monitor.
The per-rank numbers come from the oracle (the kernels need a GPU); what is under test is onda_b200.sharding
and onda_b200.switching."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import proto_oracle as po


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_sums(case, C, D, rank, world):
    """This rank's `sums` buffer, built with the oracle on its shard of the batch."""
    from onda_b200 import sharding
    feat, prior, out = sharding.shard_batch([case["feat"], case["prior"], case["out"]], world, rank)
    buf = torch.zeros(sharding.sums_numel(C, D))
    s1v, s2v, cntv, tail = sharding.split_sums(buf, C, D)
    if feat.shape[0]:
        s1, cnt = po.class_sums(feat, out)
        s2, _ = po.class_sums(feat ** 2, out)
        s1v.copy_(s1); s2v.copy_(s2); cntv.copy_(cnt)
        sigma = po.pooled_std(case["protos"], case["sq_mean"], case["counter"])
        shifted = po.shift_by_row_min(po.raw_distance(feat, case["protos"], sigma))
        q, r = po.rectify(shifted, po.to_rows(prior), 1.0)
        n = shifted.shape[0]
        tail[sharding.STAT_PROTO_CONF] = q.max(1)[0].sum()
        tail[sharding.STAT_PRIOR_CONF] = po.to_rows(prior).max(1)[0].sum()
        tail[sharding.STAT_PL_CONF] = r.max(1)[0].sum()
        tail[sharding.STAT_PL_PIXELS] = float((po.hard_labels(r, 0.3) != 255).sum())
        tail[sharding.STAT_PIXELS] = n
    return buf


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from onda_b200 import sharding, Monitor, HybridSelect
        C, D = 19, 24
        mon = Monitor(6, 0.003, "hamming")
        sel = HybridSelect(0, (0.05, 0.2), 1e-4)
        protos = sq = None
        trace = []
        for step in range(10):
            case = po.synth_case(900 + step, 5, D, 7, 9)        # 5 images over 2 ranks: 3 + 2
            if protos is None:
                protos, sq, counter = case["protos"].clone(), case["sq_mean"].clone(), case["counter"].clone()
            case["protos"], case["sq_mean"], case["counter"] = protos, sq, counter
            buf = _rank_sums(case, C, D, rank, world)
            sharding.allreduce_sums(buf)                          # the one collective of the step
            s1, s2, cnt, tail = sharding.split_sums(buf, C, D)
            stats = sharding.stats_from_tail(tail.tolist())
            mon.add({"prior static": stats["prior"]})
            sel.evaluate(mon.avg("prior static"), mon.dev_avg("prior static"))
            trace.append((sel.current, round(stats["prototypes"], 9), stats["pixels"]))
            # ma() on the all-reduced sums (prototype_handler.py:88-99)
            rho = 0.9 ** (cnt > 0).float()
            safe = torch.where(cnt > 0, cnt, torch.ones_like(cnt))
            protos = (protos.T * rho).T + ((1 - rho) * (s1.T / safe)).T
            sq = (sq.T * rho).T + ((1 - rho) * (s2.T / safe)).T
        ret[rank] = (protos, sq, trace)
    finally:
        dist.destroy_process_group()


def test_two_ranks_agree_with_single_process():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    (p0, s0, t0), (p1, s1, t1) = ret[0], ret[1]
    assert torch.equal(p0, p1) and torch.equal(s0, s1)            # bit-identical prototypes on every rank
    assert t0 == t1                                               # identical statistics and switch decisions
    assert all(px == 5 * 7 * 9 for _, _, px in t0)
    # single-process oracle over the whole batch gives the same prototypes up to fp32 summation order
    protos = sq = None
    for step in range(10):
        case = po.synth_case(900 + step, 5, 24, 7, 9)
        if protos is None:
            protos, sq = case["protos"].clone(), case["sq_mean"].clone()
        protos, sq = po.ema_update(protos, sq, case["feat"], case["out"], 0.9)
    assert float((p0 - protos).abs().max()) <= 1e-5 * float(protos.abs().max())
    assert float((s0 - sq).abs().max()) <= 1e-5 * float(sq.abs().max())


def test_shard_bounds_cover_the_batch():
    from onda_b200 import sharding
    for n in (0, 1, 5, 32, 33):
        for world in (1, 2, 4, 8):
            spans = [sharding.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(4, 2, 2)
