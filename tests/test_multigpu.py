"""Multi-GPU tests of the batch-sharded path (need >= 2 CUDA devices; skipped otherwise): R-way sharded result
vs the single-GPU result, prototypes bit-identical across ranks, NCCL and one-shot all-reduce agree."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, mode, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from onda_b200 import prototype_handler, sharding
        from oracle import proto_oracle as po
        case = po.synth_case(77, 6, 128, 33, 41)
        h = prototype_handler(ma_lambda=0.9, tau=1, thresh=0.3, distance_metric="mahalanobis",
                              process_group=dist.group.WORLD, allreduce=mode)
        h.prototypes, h.squared_mean, h.counter = (case[k].to(dev) for k in ("protos", "sq_mean", "counter"))
        labels = []
        for step in range(4):
            c = po.synth_case(100 + step, 6, 128, 33, 41, protos=case["protos"], counter=case["counter"])
            feat, prior, out = (t.to(dev) for t in sharding.shard_batch([c["feat"], c["prior"], c["out"]], world, rank))
            lab, soft = h.pseudo_labels_fused(feat, prior, out)
            h.ma(feat, out)
            labels.append(lab.cpu())
        torch.cuda.synchronize()
        ret[rank] = (h.prototypes.cpu(), h.squared_mean.cpu(), torch.cat(labels), h.last_stats)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["nccl", "oneshot"])
def test_sharded_matches_single_gpu(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), mode, ret), nprocs=world, join=True)
    (p0, s0, l0, st0), (p1, s1, l1, st1) = ret[0], ret[1]
    assert torch.equal(p0, p1) and torch.equal(s0, s1)          # identical prototypes on every rank
    assert st0["pixels"] == st1["pixels"] == 6 * 33 * 41        # statistics are global after ma()
    # single GPU over the whole batch
    from onda_b200 import prototype_handler
    from oracle import proto_oracle as po
    dev = torch.device("cuda:0")
    case = po.synth_case(77, 6, 128, 33, 41)
    h = prototype_handler(ma_lambda=0.9, tau=1, thresh=0.3, distance_metric="mahalanobis")
    h.prototypes, h.squared_mean, h.counter = (case[k].to(dev) for k in ("protos", "sq_mean", "counter"))
    labs = []
    for step in range(4):
        c = po.synth_case(100 + step, 6, 128, 33, 41, protos=case["protos"], counter=case["counter"])
        feat, prior, out = (c[k].to(dev) for k in ("feat", "prior", "out"))
        lab, _ = h.pseudo_labels_fused(feat, prior, out)
        h.ma(feat, out)
        labs.append(lab.cpu().view(6, -1))
    assert float((h.prototypes.cpu() - p0).abs().max()) <= 1e-5 * float(p0.abs().max())
    assert float((h.squared_mean.cpu() - s0).abs().max()) <= 1e-5 * float(s0.abs().max())
    # labels of the two shards, step by step, equal the single-GPU labels except where prototypes differ by rounding
    single = torch.cat([torch.cat([l[:3].reshape(-1), l[3:].reshape(-1)]) for l in labs])
    # ranks hold images [0:3] and [3:6] of each step
    per_rank = [l0.view(4, -1), l1.view(4, -1)]
    sharded = torch.cat([torch.cat([per_rank[0][s], per_rank[1][s]]) for s in range(4)])
    assert float((single != sharded).float().mean()) < 1e-3


def _graph_worker(rank, world, port, n_graphs, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from onda_b200 import prototype_handler, sharding
        from oracle import proto_oracle as po
        case = po.synth_case(77, 6, 128, 33, 41)
        sets = []
        for k in range(2):
            c = po.synth_case(200 + k, 6, 128, 33, 41, protos=case["protos"], counter=case["counter"])
            sets.append(tuple(t.to(dev) for t in sharding.shard_batch([c["feat"], c["prior"], c["out"]], world, rank)))
        results = []
        for use_graph in (False, True):
            h = prototype_handler(ma_lambda=0.9, tau=1, thresh=0.3, distance_metric="mahalanobis",
                                  process_group=dist.group.WORLD, allreduce="oneshot", tile_schedule="fixed")   # bit-for-bit comparison below
            h.prototypes, h.squared_mean, h.counter = (case[k].clone().to(dev) for k in ("protos", "sq_mean", "counter"))

            def step(k):
                feat, prior, out = sets[k]
                h.pseudo_labels_fused(feat, prior, out)
                h.ma(feat, out)
            step(0), step(1)                                  # 2 eager steps (buffers, symmetric memory)
            if use_graph:
                graphs = []
                for k in range(n_graphs):                     # capture also runs nothing: 2 + 6 replayed steps
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        step(k)
                    graphs.append(g)
                dist.barrier()
                for i in range(6):
                    graphs[i % n_graphs].replay()
                    if rank == 0 and i == 2:
                        torch.cuda.synchronize()              # rank skew: rank 1 runs ahead as far as the protocol lets it
                        import time
                        time.sleep(0.2)
            else:
                for i in range(6):
                    step(i % n_graphs)
            torch.cuda.synchronize()
            dist.barrier()
            results.append((h.prototypes.cpu().clone(), h.squared_mean.cpu().clone()))
        ret[rank] = results
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_graphs", [2, 1])
def test_graph_replayed_steps_equal_eager_steps(n_graphs):
    """The exchange fused into ma() keeps its epoch in device memory and the writer of a peer-visible slot waits for the
    peers' "done" words, so a step can be replayed from a CUDA graph -- two graphs alternating over the two slots, or ONE
    graph that reuses a single slot every step (with one rank deliberately delayed).  Eight steps (two eager + six
    replayed) give bit-identical prototypes to eight eager steps, on every rank."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_graph_worker, args=(world, _free_port(), n_graphs, ret), nprocs=world, join=True)
    for r in range(world):
        (pe, se), (pg, sg) = ret[r]
        assert torch.equal(pe, pg) and torch.equal(se, sg)
    assert torch.equal(ret[0][1][0], ret[1][1][0])
