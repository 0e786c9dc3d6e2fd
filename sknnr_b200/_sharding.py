"""Query sharding over the GPUs of one box.

The path shards embarrassingly: every rank holds the whole fitted state, takes a contiguous
block of query rows and passes the block's global start as ``row_offset`` so that sknnr's
``|idx - query_row|`` ordering key (ref:src/sknnr/_base.py:171) equals a single call's.  The only
collective is the final gather of ``(dist, idx, pred)``; ``torch.distributed`` provides it
(NCCL over NVLink on GPUs, gloo in the CPU tests).
"""

from __future__ import annotations

import numpy as np


def shard_bounds(n_rows: int, world: int, rank: int) -> tuple[int, int]:
    """Rows ``[start, stop)`` of rank ``rank``: blocks of ``ceil(n_rows / world)``."""
    per = -(-n_rows // world) if world > 0 else n_rows
    start = min(rank * per, n_rows)
    return start, min(start + per, n_rows)


def sharded_query(index, X, k, *, dst=0, group=None, device=None, **query_kw):
    """Run ``index.query`` on this rank's block of ``X`` and gather the blocks on ``dst``.

    ``X`` is the full query matrix (or this rank's view of it: only the rank's own rows are
    read).  Returns ``(dist, idx, pred)`` on ``dst`` (arrays may be None exactly as
    ``index.query`` returns them) and ``(None, None, None)`` elsewhere.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n = X.shape[0]
    lo, hi = shard_bounds(n, world, rank)
    d_blk, i_blk, p_blk = index.query(X[lo:hi], k, row_offset=lo, **query_kw)

    backend = dist.get_backend(group)
    dev = torch.device("cpu")
    if backend == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    per = -(-n // world)
    out = []
    for blk in (d_blk, i_blk, p_blk):
        if blk is None:
            out.append(None)
            continue
        pad = np.zeros((per,) + blk.shape[1:], dtype=blk.dtype)
        pad[: hi - lo] = blk
        t = torch.from_numpy(pad).to(dev)
        parts = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
        dist.gather(t, parts, dst=dst, group=group)
        if rank == dst:
            full = torch.cat(parts, dim=0)[:n].cpu().numpy()
            out.append(full)
        else:
            out.append(None)
    return tuple(out)
