"""world_size-2 gloo test (CPU) of the multi-GPU host logic: contiguous world shards, per-world
parameters keyed by the global world id, results independent of the number of ranks.  The step
path itself has no collective; the gather below is only the test's way to compare."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mars_ode_physics_b200.sharding import owner_of, shard_range  # noqa: E402

TOTAL, STEPS = 6, 25


def _run_shard(begin, end):
    import orc
    import bench
    oc = orc.OracleBatch(end - begin)
    bench.build_rover_worlds(oc, end - begin, world_offset=begin)
    for _ in range(STEPS):
        oc.collide()
        oc.step(0.001)
    return oc.state_download(5)


def _worker(rank, world_size, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    b, e = shard_range(TOTAL, rank, world_size)
    st = torch.from_numpy(_run_shard(b, e))
    gathered = [torch.empty_like(st) for _ in range(world_size)]
    dist.all_gather(gathered, st)
    if rank == 0:
        np.save(out, torch.cat(gathered, dim=2).numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partition():
    for total in (1, 7, 64, 65536, 524288):
        for ws in (1, 2, 3, 4, 8):
            seen = []
            for r in range(ws):
                b, e = shard_range(total, r, ws)
                seen += list(range(b, min(e, b + 3))) + ([e - 1] if e > b else [])
                for w in (b, e - 1):
                    if e > b:
                        assert owner_of(w, total, ws) == r
            sizes = [shard_range(total, r, ws)[1] - shard_range(total, r, ws)[0] for r in range(ws)]
            assert sum(sizes) == total and max(sizes) - min(sizes) <= 1
            assert shard_range(total, 0, ws)[0] == 0 and shard_range(total, ws - 1, ws)[1] == total


def test_two_rank_run_equals_single_process(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    sharded = np.load(out)
    single = _run_shard(0, TOTAL)
    assert sharded.shape == single.shape
    assert np.array_equal(sharded, single)  # bitwise: same per-world seeds regardless of rank count
