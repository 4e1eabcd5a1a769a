"""N > 1 host logic on CPU: two gloo ranks each take their shard of the regions (tools/workloads.shard_regions, the
same function bench.py uses under torchrun), compute it independently (here with the CPU oracle, since this suite has
no GPU), rank 0 reassembles the shards and the result must equal the unsharded call.  No collective touches the data
path: gather happens only to check."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, bam, preset, gs, q):
    import oracle_api as O
    import workloads as WL
    from bamsignals_b200 import api
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gr_all, kw, fn = WL.regions(preset, gs)
    gr, idx = WL.shard_regions(gr_all, rank, world)
    flat = WL.as_flat(getattr(O, fn)(bam, gr, **kw)).astype(np.int32)
    recs = torch.tensor([O.stats()["records"]], dtype=torch.int64)
    dist.all_reduce(recs, op=dist.ReduceOp.SUM)                      # the only "collective": bookkeeping
    sizes = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([len(idx), len(flat)], dtype=torch.int64))
    if rank == 0:
        parts = [(idx, flat)]
        for r in range(1, world):
            bi = torch.zeros(int(sizes[r][0]), dtype=torch.int64)
            bf = torch.zeros(int(sizes[r][1]), dtype=torch.int32)
            dist.recv(bi, src=r)
            dist.recv(bf, src=r)
            parts.append((bi.numpy(), bf.numpy()))
        ca = api.core_args(fn, **kw)
        off = api.output_layout(gr_all.width, ca.get("binsize", 1), ca.get("ss", False))
        merged = WL.merge_shards(parts, off)
        full = WL.as_flat(getattr(O, fn)(bam, gr_all, **kw))
        q.put((bool(np.array_equal(merged, full)), int(merged.sum()), int(recs.item())))
    else:
        dist.send(torch.from_numpy(idx.astype(np.int64)), dst=0)
        dist.send(torch.from_numpy(flat), dst=0)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("preset,gs", [("c2", 0.002), ("c5", 0.001), ("c3", 0.004)])
def test_two_rank_sharding_matches_unsharded(tmp_path_factory, preset, gs):
    import workloads as WL
    bam, _ = WL.make_bam(preset, gs, str(tmp_path_factory.mktemp("shard")))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, bam, preset, gs, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok, total, recs = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and total > 0 and recs > 0


def test_shards_partition_the_regions():
    import workloads as WL
    gr, _, _ = WL.regions("c5", 0.001)
    for world in (1, 2, 3, 8):
        seen = np.concatenate([WL.shard_regions(gr, r, world)[1] for r in range(world)])
        assert np.array_equal(np.sort(seen), np.arange(len(gr)))


def test_split_wide_regions_concatenates(tmp_path):
    """bench.py shards profile / coverage configs by bin-aligned PIECES of wide regions (tools/workloads.py
    split_wide_regions; the library does the same inside a multi-device call): the pieces' results, concatenated in
    5' -> 3' order, must be the regions' results for '+', '-' and '*' regions."""
    import bamsignals_b200 as B
    import oracle_api as O
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import workloads as WL
    bam, _ = WL.make_bam("c4", 0.002, str(tmp_path))
    lens = WL.contig_lens("c4", 0.002)
    gr = B.GRanges(WL.NAMES[:4], [1, 101, 7, 1], [lens[0], lens[1] - 200, lens[2] - 10, lens[3]], ["+", "-", "*", "-"])
    for bs, kw in ((200, dict(paired_end="midpoint", tlenFilter=(70, 200))), (7, dict(ss=True)), (1, dict(ss=True, shift=30))):
        pieces = WL.split_wide_regions(gr, 8, bs)
        assert len(pieces) > 8 * len(gr)
        assert np.array_equal(WL.as_flat(O.bamProfile(bam, pieces, binsize=bs, **kw)), WL.as_flat(O.bamProfile(bam, gr, binsize=bs, **kw)))
    assert np.array_equal(WL.as_flat(O.bamCoverage(bam, WL.split_wide_regions(gr, 8, 1), paired_end="extend")),
                          WL.as_flat(O.bamCoverage(bam, gr, paired_end="extend")))
    parts = [WL.shard_regions(gr, r, 4, split_align=200)[0] for r in range(4)]
    assert sum(len(p) for p in parts) == len(WL.split_wide_regions(gr, 4, 200))
