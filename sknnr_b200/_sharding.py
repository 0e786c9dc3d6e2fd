"""Query sharding over the GPUs of one box.

The path shards embarrassingly: every GPU holds the whole fitted state, takes a contiguous
block of query rows and is told the block's global start as ``row_offset``, so that sknnr's
``|idx - query_row|`` ordering key (ref:src/sknnr/_base.py:171) equals a single call's.  Two ways
to drive it:

* **in one process** (:class:`MultiDeviceIndex`, what the estimators use when
  ``SKNNR_B200_DEVICES`` names several GPUs): one index handle per GPU, one host thread per
  handle, every GPU copies its block of the caller's rows in and DMAs its results straight into
  its slice of the caller's result arrays.  No collective is needed at all - the reference's own
  knob for this is ``n_jobs`` (ref:src/sknnr/_base.py:206,263), which stays accepted and unused;
* **one process per GPU** (:func:`sharded_query`, ``torch.distributed``): every rank answers
  its block through the device-pointer call and the blocks are gathered on ``dst`` with one
  collective per result array (NCCL over NVLink; gloo with host arrays in the CPU tests).
  ``bench.py`` goes one step further and lets the finishing kernels store into the root's arrays
  over NVLink (``sknnr_ipc_*``), which removes the collective from the data path.
"""

from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np


def shard_bounds(n_rows: int, world: int, rank: int) -> tuple[int, int]:
    """Rows ``[start, stop)`` of rank ``rank``: blocks of ``ceil(n_rows / world)``."""
    per = -(-n_rows // world) if world > 0 else n_rows
    start = min(rank * per, n_rows)
    return start, min(start + per, n_rows)


def devices_from_env() -> list[int] | None:
    """``SKNNR_B200_DEVICES``: comma-separated device ordinals, or ``all``.  None / one entry:
    single-device operation (the default)."""
    v = os.environ.get("SKNNR_B200_DEVICES", "").strip()
    if not v:
        return None
    if v.lower() == "all":
        from . import _lib as L

        devs = list(range(L.device_count()))
    else:
        devs = [int(t) for t in v.split(",") if t.strip() != ""]
    return devs if len(devs) > 1 else None


class MultiDeviceIndex:
    """The same fitted state on several GPUs behind the interface of one index.

    ``make(device)`` builds the per-device handle (a ``KNNIndex`` or ``HammingIndex``).  Queries
    with rows are split into contiguous blocks, one per device, answered concurrently (the C
    calls release the GIL) and written into disjoint slices of shared result arrays; ``X=None``
    self-queries and the small helper calls run on the first device.  Blocks below
    ``min_rows_per_device`` are not worth a second GPU and stay on the first one.
    """

    min_rows_per_device = 65536

    def __init__(self, make, devices):
        self.devices = list(devices)
        self.parts = [make(d) for d in self.devices]
        self._pool = ThreadPoolExecutor(max_workers=len(self.parts), thread_name_prefix="sknnr-b200-dev")
        first = self.parts[0]
        for name in ("n_ref", "n_out", "d_in", "d_out", "n_trees"):
            if hasattr(first, name):
                setattr(self, name, getattr(first, name))

    # -- helpers that do not shard ---------------------------------------------------------
    def __getattr__(self, name):   # transform, weighted_average, stats, ... : first device
        if name in ("parts", "devices", "_pool"):
            raise AttributeError(name)
        return getattr(self.parts[0], name)

    def close(self):
        for p in self.parts:
            p.close()
        self._pool.shutdown(wait=False)

    def _split(self, n_rows):
        world = max(1, min(len(self.parts), n_rows // self.min_rows_per_device))
        return [shard_bounds(n_rows, world, r) for r in range(world)]

    def _fan_out(self, method, X, k, args, kw):
        from ._engine import _result_empty, _weights_mode
        from . import _lib as L

        X = np.asarray(X)
        n = X.shape[0]
        blocks = self._split(n)
        if len(blocks) == 1:
            a_0 = tuple(a[0] if isinstance(a, _PerDevice) else a for a in args)
            return getattr(self.parts[0], method)(*a_0, X, k, **kw)
        base = int(kw.pop("row_offset", 0))
        mode = _weights_mode(kw.get("weights"), kw.get("with_pred", False))
        dist = _result_empty((n, k), np.float64) if kw.get("return_distance", True) else None
        idx = _result_empty((n, k), np.int64) if kw.get("return_index", True) else None
        pred = _result_empty((n, self.n_out), np.float64) if mode != L.W_NONE else None

        def run(r):
            lo, hi = blocks[r]
            out = tuple(None if a is None else a[lo:hi] for a in (dist, idx, pred))
            # (forest handles are per device as well: args hold one entry per part)
            a_r = tuple(a[r] if isinstance(a, _PerDevice) else a for a in args)
            getattr(self.parts[r], method)(*a_r, X[lo:hi], k, row_offset=base + lo, out=out, **kw)

        for f in [self._pool.submit(run, r) for r in range(len(blocks))]:
            f.result()
        return dist, idx, pred

    def query(self, X, k, **kw):
        if X is None or kw.get("exclude_self"):
            return self.parts[0].query(X, k, **kw)
        return self._fan_out("query", X, k, (), kw)

    def query_forest(self, forest, X, k, **kw):
        return self._fan_out("query_forest", X, k, (forest,), kw)


class _PerDevice(list):
    """One helper object (e.g. a ForestIndex) per device of a MultiDeviceIndex, in device order."""


def sharded_query(index, X, k, *, dst=0, group=None, device=None, **query_kw):
    """One process per GPU: answer this rank's block of ``X`` and gather the blocks on ``dst``.

    ``X`` is the full query matrix (or this rank's view of it: only the rank's own rows are
    read).  Returns ``(dist, idx, pred)`` on ``dst`` (arrays may be None exactly as
    ``index.query`` returns them) and ``(None, None, None)`` elsewhere.

    NCCL: the block goes to the device once, is answered through the device-pointer call
    (``index.query_device``) into device tensors, and those are what the collective gathers -
    nothing is staged through host memory between the search and the gather.  Other backends
    (gloo in the CPU tests): ``index.query`` on host arrays.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = X.shape[0]
    lo, hi = shard_bounds(n, world, rank)
    per = -(-n // world)
    rows = hi - lo
    want_pred = bool(query_kw.get("with_pred", False))
    on_device = dist.get_backend(group) == "nccl" and hasattr(index, "query_device")

    if on_device:
        dev = torch.device("cuda", index.device if device is None else device)
        torch.cuda.set_device(dev)
        blk = np.ascontiguousarray(X[lo:hi], dtype=np.float64)
        x_d = torch.from_numpy(blk).to(dev)
        n_out = getattr(index, "n_out", 0)
        outs = [torch.zeros((per, k), dtype=torch.float64, device=dev),
                torch.zeros((per, k), dtype=torch.int64, device=dev),
                torch.zeros((per, n_out), dtype=torch.float64, device=dev) if want_pred else None]
        if rows:
            st = torch.cuda.current_stream(dev)
            index.query_device(x_d.data_ptr(), False, rows, blk.shape[1], k, dist_ptr=outs[0].data_ptr(),
                               idx_ptr=outs[1].data_ptr(), pred_ptr=outs[2].data_ptr() if want_pred else 0,
                               weights=query_kw.get("weights"), deterministic=query_kw.get("deterministic", True),
                               decimals=query_kw.get("decimals", 10), row_offset=lo + int(query_kw.get("row_offset", 0)),
                               transformed=query_kw.get("transformed", False), stream=st.cuda_stream)
    else:
        d_blk, i_blk, p_blk = index.query(X[lo:hi], k, row_offset=lo, **query_kw)
        outs = []
        for blk in (d_blk, i_blk, p_blk):
            if blk is None:
                outs.append(None)
                continue
            pad = np.zeros((per,) + blk.shape[1:], dtype=blk.dtype)
            pad[:rows] = blk
            outs.append(torch.from_numpy(pad))

    result = []
    for t in outs:
        if t is None:
            result.append(None)
            continue
        parts = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
        dist.gather(t, parts, dst=dst, group=group)
        result.append(torch.cat(parts, dim=0)[:n].cpu().numpy() if rank == dst else None)
    return tuple(result)
