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


def test_devices_from_env_and_multi_device_fan_out(monkeypatch):
    """Host logic of the in-process multi-device index with stub parts (no GPU): contiguous blocks,
    global row offsets, disjoint slices of shared result arrays."""
    from sknnr_b200 import _sharding as S

    monkeypatch.delenv("SKNNR_B200_DEVICES", raising=False)
    assert S.devices_from_env() is None
    monkeypatch.setenv("SKNNR_B200_DEVICES", "3")
    assert S.devices_from_env() is None
    monkeypatch.setenv("SKNNR_B200_DEVICES", "0, 2,5")
    assert S.devices_from_env() == [0, 2, 5]

    calls = []

    class Part:
        n_ref, n_out, d_in, d_out = 10, 2, 3, 3

        def __init__(self, dev):
            self.dev = dev

        def query(self, X, k, row_offset=0, out=None, **kw):
            calls.append((self.dev, row_offset, X.shape[0]))
            if out is None:
                out = (np.empty((X.shape[0], k)), np.empty((X.shape[0], k), dtype=np.int64), np.empty((X.shape[0], 2)))
            d, i, p = out
            rows = np.arange(row_offset, row_offset + X.shape[0])
            d[:] = rows[:, None] + 0.5
            i[:] = rows[:, None] * 10 + np.arange(k)
            if p is not None:
                p[:] = X[:, :2] * (self.dev + 1)
            return d, i, p

        def close(self):
            pass

    monkeypatch.setattr(S.MultiDeviceIndex, "min_rows_per_device", 4)
    m = S.MultiDeviceIndex(Part, [0, 1, 2])
    X = np.arange(33, dtype=float)[:, None] * np.ones((1, 3))
    d, i, p = m.query(X, 2, row_offset=100, weights="uniform", with_pred=True)
    assert sorted(calls) == [(0, 100, 11), (1, 111, 11), (2, 122, 11)]
    assert np.array_equal(d[:, 0], np.arange(100, 133) + 0.5) and np.array_equal(i[:, 1], np.arange(100, 133) * 10 + 1)
    assert np.array_equal(p[:11], X[:11, :2]) and np.array_equal(p[22:], X[22:, :2] * 3)
    calls.clear()
    d, i, p = m.query(X[:5], 2)            # too few rows for a second device: one call, own outputs
    assert calls == [(0, 0, 5)] and d.shape == (5, 2)
