"""Multi-GPU surface: the in-process multi-device index behind the estimators
and the one-process-per-GPU ``sharded_query``.  With a single GPU in the box the multi-device index
is exercised with two handles on the same device (the sharding logic - contiguous blocks, global
``row_offset``, disjoint result slices - is identical); the NCCL test needs two GPUs."""

import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _data(seed=0, n_ref=3000, d=12, n_q=5000, n_out=4):
    rng = np.random.default_rng(seed)
    R = rng.standard_normal((n_ref, d))
    R[100:140] = R[:40]                               # exact ties: the |idx - row| key decides
    y = rng.standard_normal((n_ref, n_out))
    Q = np.vstack([R[rng.integers(0, n_ref, n_q // 2)], rng.standard_normal((n_q - n_q // 2, d))])
    return R, y, Q


def test_multi_device_index_equals_single_call(monkeypatch):
    from sknnr_b200._engine import HammingIndex, KNNIndex
    from sknnr_b200._sharding import MultiDeviceIndex

    monkeypatch.setattr(MultiDeviceIndex, "min_rows_per_device", 64)
    R, y, Q = _data()
    single = KNNIndex(R, None, None, None, y, device=0)
    multi = MultiDeviceIndex(lambda dev: KNNIndex(R, None, None, None, y, device=dev), [0, 0, 0])
    for off in (0, 777):
        a = single.query(Q, 5, transformed=True, weights="distance", with_pred=True, row_offset=off)
        b = multi.query(Q, 5, transformed=True, weights="distance", with_pred=True, row_offset=off)
        for x, z in zip(a, b):
            assert np.array_equal(x, z)
    a = single.query(None, 5, exclude_self=True)
    b = multi.query(None, 5, exclude_self=True)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])
    # Hamming twin
    rng = np.random.default_rng(1)
    Rc = rng.integers(0, 9, size=(2000, 60)).astype(np.uint16)
    Qc = Rc[rng.integers(0, 2000, 900)].copy()
    Qc[rng.random(Qc.shape) < 0.2] = 11
    w = np.full(60, 1.0 / 60)
    hs = HammingIndex(Rc, w, y[:2000], device=0)
    hm = MultiDeviceIndex(lambda dev: HammingIndex(Rc, w, y[:2000], device=dev), [0, 0])
    a = hs.query(Qc, 6, weights="uniform", with_pred=True)
    b = hm.query(Qc, 6, weights="uniform", with_pred=True)
    for x, z in zip(a, b):
        assert np.array_equal(x, z)


def test_estimators_use_every_listed_device(monkeypatch):
    """SKNNR_B200_DEVICES behind est.predict / est.kneighbors: same numbers as one device."""
    from sknnr_b200 import EuclideanKNNRegressor, RFNNRegressor
    from sknnr_b200._sharding import MultiDeviceIndex

    R, y, Q = _data(n_ref=1500, n_q=3000)
    one = EuclideanKNNRegressor(n_neighbors=4, weights="distance").fit(R, y)
    p1, (d1, i1) = one.predict(Q), one.kneighbors(Q)
    monkeypatch.setenv("SKNNR_B200_DEVICES", "0,0")
    monkeypatch.setattr(MultiDeviceIndex, "min_rows_per_device", 64)
    two = EuclideanKNNRegressor(n_neighbors=4, weights="distance").fit(R, y)
    assert isinstance(two.regressor_._get_index(), MultiDeviceIndex)
    p2, (d2, i2) = two.predict(Q), two.kneighbors(Q)
    assert np.array_equal(p1, p2) and np.array_equal(i1, i2) and np.array_equal(d1, d2)
    assert two.independent_score_ == one.independent_score_
    rf2 = RFNNRegressor(n_estimators=5, random_state=0, n_neighbors=3).fit(R[:600], y[:600])
    monkeypatch.delenv("SKNNR_B200_DEVICES")
    rf1 = RFNNRegressor(n_estimators=5, random_state=0, n_neighbors=3).fit(R[:600], y[:600])
    assert np.array_equal(rf1.predict(Q), rf2.predict(Q))
    a, b = rf1.kneighbors(Q), rf2.kneighbors(Q)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


_WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from sknnr_b200._engine import KNNIndex
from sknnr_b200._sharding import sharded_query
rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
rng = np.random.default_rng(0)
R = rng.standard_normal((3000, 12)); R[100:140] = R[:40]
y = rng.standard_normal((3000, 4))
Q = np.vstack([R[rng.integers(0, 3000, 2500)], rng.standard_normal((2501, 12))])
ix = KNNIndex(R, None, None, None, y, device=lr)
d, i, p = sharded_query(ix, Q, 5, dst=0, transformed=True, weights="distance", with_pred=True)
if rank == 0:
    d1, i1, p1 = ix.query(Q, 5, transformed=True, weights="distance", with_pred=True)
    assert np.array_equal(i, i1) and np.array_equal(d, d1) and np.array_equal(p, p1), "gathered != single call"
    print("NCCL_SHARDED_OK", i.shape)
dist.destroy_process_group()
"""


def test_sharded_query_nccl_two_ranks_equals_single_call(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip(f"needs 2 GPUs for the NCCL gather, this box has {torch.cuda.device_count()}")
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "NCCL_SHARDED_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
