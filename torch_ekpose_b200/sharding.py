"""Multi-GPU sharding of the post-processing path: one process per GPU, images partitioned.

Images are independent (the reference processes them one at a time: lib/evaluate/estimator.py:80,
run_video.py:57-64, eval.py:144-167), so the batch is cut into contiguous slices, every rank
runs stages 1-5 on its own slice on its own GPU, and the ONLY exchange is a final gather of the
fixed-size result tables (``torch.distributed.all_gather`` -- NCCL over NVLink on the GPU box,
gloo in the CPU tests).  There is no data-path collective.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def shard_bounds(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced partition: the first ``n_items % world`` ranks get one extra image."""
    if world < 1 or not (0 <= rank < world) or n_items < 0:
        raise ValueError(f"bad shard request: n_items={n_items} world={world} rank={rank}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_tables(num_humans: np.ndarray, subset: np.ndarray, n_total: int, group=None, device=None):
    """all_gather the per-rank result tables and return them in global image order.

    num_humans [n_local] int32, subset [n_local, max_humans, 20] float32 (rows beyond
    num_humans[i] are ignored).  Every rank must pass the same max_humans.  Returns
    (num_humans [n_total], subset [n_total, max_humans, 20]) on every rank.
    """
    if dist is None or not dist.is_initialized():
        return num_humans.copy(), subset.copy()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    cap = max(shard_bounds(n_total, world, r)[1] - shard_bounds(n_total, world, r)[0] for r in range(world))
    mh = subset.shape[1]
    dev = device if device is not None else torch.device("cpu")
    num_pad = torch.zeros(cap, dtype=torch.int32, device=dev)
    sub_pad = torch.zeros((cap, mh, 20), dtype=torch.float32, device=dev)
    n_local = len(num_humans)
    if n_local:
        num_pad[:n_local] = torch.from_numpy(np.ascontiguousarray(num_humans, np.int32)).to(dev)
        sub_pad[:n_local] = torch.from_numpy(np.ascontiguousarray(subset, np.float32)).to(dev)
    nums = [torch.zeros_like(num_pad) for _ in range(world)]
    subs = [torch.zeros_like(sub_pad) for _ in range(world)]
    dist.all_gather(nums, num_pad, group=group)
    dist.all_gather(subs, sub_pad, group=group)
    out_num = np.zeros(n_total, np.int32)
    out_sub = np.zeros((n_total, mh, 20), np.float32)
    for r in range(world):
        lo, hi = shard_bounds(n_total, world, r)
        out_num[lo:hi] = nums[r][:hi - lo].cpu().numpy()
        out_sub[lo:hi] = subs[r][:hi - lo].cpu().numpy()
    return out_num, out_sub


def postprocess_sharded(heat, paf, *, layout: str = "nchw", frontend: str = "dense", thr: float = 0.15,
                        materialize: bool = False, max_humans: int = 64, max_peaks: int = 1024, group=None,
                        compute: Optional[Callable] = None):
    """Every rank passes the SAME full batch (host arrays or CPU tensors); rank r post-processes
    images shard_bounds(n, world, r) on its own GPU and all ranks get the gathered tables
    (num_humans [n], subset [n, max_humans, 20]).

    ``compute(heat_slice, paf_slice) -> (num_humans, subset)`` replaces the GPU call in the CPU
    tests of this host logic; by default it is a PostProcessor on cuda:LOCAL_RANK.
    """
    import os
    n = heat.shape[0]
    world = dist.get_world_size(group) if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    lo, hi = shard_bounds(n, world, rank)
    gather_dev = None
    if compute is None:
        from .paf_to_pose import PostProcessor
        local = int(os.environ.get("LOCAL_RANK", rank))
        shp = heat.shape
        h, w = (shp[2], shp[3]) if layout == "nchw" else (shp[1], shp[2])
        pp = PostProcessor(device=local, max_batch=max(hi - lo, 1), max_h=h, max_w=w, max_peaks=max_peaks, max_humans=max_humans)

        def compute(hs, ps):
            pp.run(hs, ps, layout=layout, frontend=frontend, thr=thr, materialize=materialize)
            r = pp.results()
            return r["num_humans"], r["subset"]
        if torch is not None and torch.cuda.is_available():
            gather_dev = torch.device("cuda", local)
    if hi > lo:
        num, sub = compute(heat[lo:hi], paf[lo:hi])
    else:
        num, sub = np.zeros(0, np.int32), np.zeros((0, max_humans, 20), np.float32)
    return gather_tables(np.asarray(num), np.asarray(sub), n, group=group, device=gather_dev)
