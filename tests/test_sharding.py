"""CPU tests of the multi-GPU host logic: world_size-2 gloo run of the sharded query with a
stub index (the oracle standing in for the device) must equal the single-call result."""

import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from sknnr_b200._sharding import shard_bounds


def test_shard_bounds_cover_rows_exactly():
    for n in (0, 1, 7, 8, 1000, 1001):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            assert all(0 <= lo <= hi <= n for lo, hi in blocks)


class _OracleIndex:
    """Test double with KNNIndex.query's signature, answered by the CPU oracle."""

    def __init__(self, fit_Z, y):
        from oracle import sknnr_oracle as orc

        self.orc = orc
        self.st = orc.FittedState("euclidean", fit_Z=fit_Z, y=y)

    def query(self, X, k, row_offset=0, **kw):
        d, i = self.orc.kneighbors(self.st, X, k=k, row_offset=row_offset, transformed=True)
        return d, i, self.orc.weighted_average(self.st.y, i)


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from sknnr_b200._sharding import sharded_query

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # duplicated reference rows make the |idx - row| tie-break matter: a shard that forgot
        # its row offset would order them differently
        R = np.repeat(np.arange(6, dtype=float), 2)[:, None] * np.ones((1, 3))
        y = np.arange(12, dtype=float)[:, None]
        X = np.repeat(np.arange(6, dtype=float), 2)[:, None] * np.ones((1, 3))[:, :3]
        X = np.vstack([X, X[:3]])
        ix = _OracleIndex(R, y)
        d, i, p = sharded_query(ix, X, 3, dst=0)
        if rank == 0:
            d1, i1, p1 = ix.query(X, 3)
            q.put((np.array_equal(i, i1), np.array_equal(d, d1), np.allclose(p, p1), i.shape))
        else:
            assert d is None and i is None and p is None
    finally:
        dist.destroy_process_group()


def test_sharded_query_equals_single_call_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] and res[1] and res[2] and res[3] == (15, 3)
