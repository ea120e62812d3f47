"""N>1 host logic on CPU: world_size-2 gloo run of the shard assignment and the
timing reduction bench.py uses (MAX of elapsed time, SUM of audio), no GPU."""
import importlib
import os

import numpy as np
import pytest

import __graft_entry__ as ge


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ge.package()
    sh = importlib.import_module("pockettts_b200.sharding")
    wl = importlib.import_module("pockettts_b200.workload")
    lo, hi = sh.shard_range(rank, world, 6)
    chunks = wl.make_workload(hi - lo, 5, 60, seed=7, first_utt=lo)
    frames = float(sum(c.expected_frames for c in chunks))
    utts = sorted({c.utt for c in chunks})
    ms, fr, sm, wall = sh.reduce_timing(dist, torch, "cpu", 100.0 * (rank + 1), frames, frames * 1920, 50.0 + rank)
    q.put((rank, utts, frames, ms, fr, sm, wall))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_reduction():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29640 + os.getpid() % 200
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=240) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, u0, f0, ms0, fr0, sm0, w0), (r1, u1, f1, ms1, fr1, sm1, w1) = res
    assert u0 == list(range(0, 6)) and u1 == list(range(6, 12))     # disjoint, contiguous shards
    assert ms0 == ms1 == 200.0 and w0 == w1 == 51.0                  # MAX over ranks
    assert fr0 == fr1 == f0 + f1 and sm0 == sm1 == (f0 + f1) * 1920  # SUM over ranks
    ge.package()
    sh = importlib.import_module("pockettts_b200.sharding")
    assert sh.job_throughput(fr0, ms0) == pytest.approx((f0 + f1) * 0.08 / 0.2)


def test_single_rank_and_bad_rank():
    import torch
    ge.package()
    sh = importlib.import_module("pockettts_b200.sharding")
    assert sh.shard_range(3, 8, 512) == (1536, 2048)
    with pytest.raises(ValueError):
        sh.shard_range(8, 8, 512)
    assert sh.reduce_timing(None, torch, "cpu", 5.0, 10.0, 20.0, 6.0) == (5.0, 10.0, 20.0, 6.0)
