"""World-size-2 checks of the N>1 host logic on CPU (gloo): contiguous slices tile the batch exactly (same rule as the
C ABI's per-device split), per-rank seeds differ, and the max-over-ranks timing reduction bench.py uses."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n, q):
    import torch
    import torch.distributed as dist
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = bench.shard_bounds(n, world, rank)
    t = torch.tensor([lo, hi], dtype=torch.int64)
    allb = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allb, t)
    ms = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.barrier()
    if rank == 0:
        q.put(([tuple(int(x) for x in b) for b in allb], float(ms.item())))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [0, 1, 5, 1 << 20, (1 << 20) + 3])
def test_slices_tile_the_batch_world2(n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    bounds, ms = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ms == 11.0
    assert bounds[0][0] == 0 and bounds[-1][1] == n
    assert all(bounds[i][1] == bounds[i + 1][0] for i in range(len(bounds) - 1))
    per = (n + 1) // 2
    assert all(hi - lo <= per for lo, hi in bounds)


def test_shard_bounds_matches_c_abi_rule():
    import bench
    for n in (0, 1, 7, 8, 9, 1000, 1 << 24):
        for world in (1, 2, 4, 8):
            per = (n + world - 1) // world
            cover = []
            for r in range(world):
                lo, hi = bench.shard_bounds(n, world, r)
                assert lo == min(n, r * per) and hi == min(n, lo + per)
                cover += list(range(lo, hi)) if n <= 1000 else []
            if n <= 1000:
                assert cover == list(range(n))
