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


def _rows_worker(rank, world, port, out):
    """The two collectives of the row-partitioned CG (sharded._Comm) on real ranks: in-place block all-gather of the
    iterate and all-reduce of the fp64 dot products -- a Jacobi-CG written with torch ops on each rank's row block plays
    the kernels' role so that the exchange pattern of sharded._solve_rows is what is under test."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import numpy as np

    from graphlearninglayer_b200.sharded import _Comm, m_block

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = _Comm()
        assert comm.real and comm.world == world and comm.ranks == [rank]
        rng = np.random.default_rng(0)                      # same system on every rank
        m, l = 203, 5
        M = rng.random((m, m)) * (rng.random((m, m)) < 0.05)
        M = np.triu(M, 1)
        M = M + M.T
        A = torch.as_tensor(np.diag(M.sum(1) + 0.3) - M)
        B = torch.as_tensor(rng.standard_normal((m, l)))
        lo, hi, per = m_block(m, rank, world)
        dinv = 1.0 / torch.diagonal(A)[lo:hi, None]
        u_full = torch.zeros((world * per, l), dtype=torch.float64)
        x_full = torch.zeros_like(u_full)
        r = B[lo:hi].clone()
        p = torch.zeros_like(r)
        sv = torch.zeros_like(r)
        u_full[lo:hi] = r * dinv
        g_old = a_old = None
        for it in range(300):
            comm.all_gather_blocks_(u_full, per)
            u = u_full[lo:hi]
            w = A[lo:hi] @ u_full[:m]
            sums = torch.cat([(r * u).sum(0), (w * u).sum(0), (r * r).sum(0)])
            comm.all_reduce_sum_({rank: sums})
            g, d, rr = sums[:l], sums[l:2 * l], sums[2 * l:]
            if rr.max().sqrt() < 1e-11:
                break
            beta = torch.zeros(l, dtype=torch.float64) if it == 0 else g / g_old
            alpha = g / (d - beta * g / a_old) if it else g / d
            p = u + beta * p
            sv = w + beta * sv
            x_full[lo:hi] += alpha * p
            r = r - alpha * sv
            u_full[lo:hi] = r * dinv
            g_old, a_old = g, alpha
        comm.all_gather_blocks_(x_full, per)
        err = float((A @ x_full[:m] - B).abs().max())
        out.put((rank, it, err))
    finally:
        dist.destroy_process_group()


def test_row_partitioned_cg_exchange_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rows_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, it0, e0), (_, it1, e1) = res
    assert it0 == it1 and 0 < it0 < 300          # every rank takes the same decisions from the same reduced sums
    assert e0 < 1e-9 and e1 < 1e-9               # and ends with the full solution


def test_m_block_partition():
    sys.path.insert(0, ROOT)
    from graphlearninglayer_b200.sharded import m_block

    for m, world in [(2300, 5), (983040, 8), (31, 4), (512, 1)]:
        edges = []
        for r in range(world):
            lo, hi, per = m_block(m, r, world)
            assert per % 32 == 0 and lo == min(m, r * per) and hi - lo <= per
            edges.append((lo, hi))
        assert edges[0][0] == 0 and edges[-1][1] == m and all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
