"""World-size-2 gloo tests (CPU) of the data-parallel plumbing: rank seeds, unit partition, max-over-ranks timing."""
import os
import socket
import sys

import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from graphlearninglayer_b200 import ranks

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert ranks.rank_info() == (rank, world, rank)
        mine = list(ranks.units_for_rank(7, rank, world))
        # rank 1 is "slower": the job time is the max, the job work is the sum
        units, ms, thr = ranks.aggregate_throughput(len(mine), 10.0 * (rank + 1))
        # the independent graphs of different ranks must differ: exchange the seeds
        seeds = [None] * world
        dist.all_gather_object(seeds, ranks.rank_seed(1000, rank))
        dist.barrier()
        out.put((rank, mine, units, ms, thr, seeds))
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, u0, units0, ms0, thr0, seeds0), (r1, u1, units1, ms1, thr1, seeds1) = res
    assert u0 == [0, 1, 2] and u1 == [3, 4, 5, 6]               # contiguous, disjoint, complete
    assert units0 == units1 == 7 and ms0 == ms1 == 20.0           # sum of work, max of time, identical on all ranks
    assert abs(thr0 - 7 / 0.020) < 1e-9
    assert seeds0 == seeds1 and len(set(seeds0)) == 2


def test_single_process_degenerates():
    sys.path.insert(0, ROOT)
    from graphlearninglayer_b200 import ranks

    assert ranks.aggregate_throughput(3, 6.0) == (3, 6.0, 500.0)
    assert list(ranks.units_for_rank(5, 0, 1)) == [0, 1, 2, 3, 4]
    assert ranks.rank_seed(1000, 0) == 1000 and ranks.rank_seed(1000, 3) != ranks.rank_seed(1001, 2)


def test_row_and_column_blocks_partition():
    """Slicing rules of the sharded (one graph on N GPUs) path: blocks are disjoint, complete, tile-aligned."""
    sys.path.insert(0, ROOT)
    from graphlearninglayer_b200.sharded import col_block, row_block

    for n, world in [(3000, 3), (1 << 20, 8), (130, 4), (4608, 2)]:
        rows = []
        for r in range(world):
            lo, hi, per = row_block(n, r, world)
            assert per % 128 == 0 and lo % 128 == 0 or lo == n
            assert hi - lo <= per
            rows += list(range(lo, hi)) if n <= 5000 else [lo, hi]
        if n <= 5000:
            assert rows == list(range(n))
        else:
            assert rows[0] == 0 and rows[-1] == n and all(rows[2 * i + 1] == rows[2 * i + 2] for i in range(world - 1))
    for l, world in [(100, 8), (10, 4), (13, 3), (3, 8)]:
        cols = []
        for r in range(world):
            lo, hi, per = col_block(l, r, world)
            assert hi - lo <= per
            cols += list(range(lo, hi))
        assert cols == list(range(l))
