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


def _gather_device(group=None):
    """NCCL gathers CUDA tensors, gloo (the CPU tests) host tensors."""
    backend = str(dist.get_backend(group)).lower()
    if "nccl" in backend:
        import os
        return torch.device("cuda", int(os.environ.get("LOCAL_RANK", torch.cuda.current_device())))
    return torch.device("cpu")


def gather_tables(num_humans: np.ndarray, subset: np.ndarray, n_total: int, group=None, device=None, overflow=None, status: int = 0):
    """all_gather the per-rank result tables and return them in global image order.

    num_humans [n_local] int32, subset [n_local, max_humans, 20] float32 (rows beyond
    num_humans[i] are ignored), optional overflow [n_local] (EKP_OVF_* bits) and a per-rank ``status``
    (0 = fine).  Every rank must pass the same max_humans.  Returns
    (num_humans [n_total], subset [n_total, max_humans, 20], overflow [n_total], status [world]) on every
    rank -- a rank that failed locally still takes part in the collective, so nobody is left waiting.
    """
    n_local = len(num_humans)
    ovf_local = np.zeros(n_local, np.uint32) if overflow is None else np.asarray(overflow, np.uint32)
    if dist is None or not dist.is_initialized():
        return num_humans.copy(), subset.copy(), ovf_local.copy(), np.array([status], np.int32)
    world = dist.get_world_size(group)
    cap = max(shard_bounds(n_total, world, r)[1] - shard_bounds(n_total, world, r)[0] for r in range(world))
    mh = subset.shape[1]
    dev = device if device is not None else _gather_device(group)
    # one int32 table [cap + 1, 2]: row 0 = (status, n_local), then (num_humans, overflow) per image; one float table
    meta = torch.zeros((cap + 1, 2), dtype=torch.int32)
    sub_pad = torch.zeros((cap, mh, 20), dtype=torch.float32)
    meta[0, 0], meta[0, 1] = int(status), n_local
    if n_local:
        meta[1:n_local + 1, 0] = torch.from_numpy(np.ascontiguousarray(num_humans, np.int32))
        meta[1:n_local + 1, 1] = torch.from_numpy(ovf_local.astype(np.int64)).to(torch.int32)
        sub_pad[:n_local] = torch.from_numpy(np.ascontiguousarray(subset, np.float32))
    meta, sub_pad = meta.to(dev), sub_pad.to(dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    subs = [torch.zeros_like(sub_pad) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    dist.all_gather(subs, sub_pad, group=group)
    out_num = np.zeros(n_total, np.int32)
    out_sub = np.zeros((n_total, mh, 20), np.float32)
    out_ovf = np.zeros(n_total, np.uint32)
    out_status = np.zeros(world, np.int32)
    for r in range(world):
        lo, hi = shard_bounds(n_total, world, r)
        m = metas[r].cpu().numpy()
        out_status[r] = m[0, 0]
        out_num[lo:hi] = m[1:hi - lo + 1, 0]
        out_ovf[lo:hi] = m[1:hi - lo + 1, 1].astype(np.uint32)
        out_sub[lo:hi] = subs[r][:hi - lo].cpu().numpy()
    return out_num, out_sub, out_ovf, out_status


class ShardError(RuntimeError):
    """Raised on EVERY rank after the gather when any rank failed or overflowed."""


def postprocess_sharded(heat, paf, *, layout: str = "nchw", frontend: str = "reference", thr: float = 0.15,
                        materialize: bool = False, max_humans: int = 64, max_peaks: int = 1024, max_part: int = 0,
                        max_cand: int = 0, group=None, compute: Optional[Callable] = None):
    """Every rank passes the SAME full batch (host arrays or CPU tensors); rank r post-processes
    images shard_bounds(n, world, r) on its own GPU and all ranks get the gathered tables
    (num_humans [n], subset [n, max_humans, 20]).

    ``compute(heat_slice, paf_slice) -> (num_humans, subset[, overflow])`` replaces the GPU call in the CPU
    tests of this host logic; by default it is a PostProcessor on cuda:LOCAL_RANK.  Nothing raises before the
    collective: a rank whose shard failed (bad device, CUDA error) or overflowed a capacity reports that through the
    gather, and then EVERY rank raises ShardError naming the ranks / images concerned.
    """
    import os
    n = heat.shape[0]
    world = dist.get_world_size(group) if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    lo, hi = shard_bounds(n, world, rank)
    num, sub, ovf = np.zeros(0, np.int32), np.zeros((0, max_humans, 20), np.float32), np.zeros(0, np.uint32)
    status, err = 0, None
    pp = None
    try:
        if compute is None:
            from .paf_to_pose import PostProcessor
            local = int(os.environ.get("LOCAL_RANK", rank))
            shp = heat.shape
            h, w = (shp[2], shp[3]) if layout == "nchw" else (shp[1], shp[2])
            pp = PostProcessor(device=local, max_batch=max(hi - lo, 1), max_h=h, max_w=w, max_peaks=max_peaks, max_humans=max_humans,
                               max_part=max_part, max_cand=max_cand)

            def compute(hs, ps):
                pp.run(hs, ps, layout=layout, frontend=frontend, thr=thr, materialize=materialize)
                r = pp.results(raise_on_overflow=False)
                return r["num_humans"], r["subset"], r["overflow"]
        if hi > lo:
            out = compute(heat[lo:hi], paf[lo:hi])
            num, sub = np.asarray(out[0]), np.asarray(out[1])
            ovf = np.asarray(out[2], np.uint32) if len(out) > 2 else np.zeros(hi - lo, np.uint32)
            if sub.shape[1:] != (max_humans, 20) or len(num) != hi - lo:
                raise ValueError(f"compute returned tables of shape {num.shape} / {sub.shape} for {hi - lo} images")
    except Exception as e:   # reported through the gather, raised on every rank afterwards
        status, err = 1, e
        num, sub, ovf = np.zeros(hi - lo, np.int32), np.zeros((hi - lo, max_humans, 20), np.float32), np.zeros(hi - lo, np.uint32)
    finally:
        if pp is not None:
            pp.close()
    g_num, g_sub, g_ovf, g_status = gather_tables(num, sub, n, group=group, overflow=ovf, status=status)
    if g_status.any() or g_ovf.any():
        bad_ranks = [int(r) for r in np.nonzero(g_status)[0]]
        bad_imgs = [int(i) for i in np.nonzero(g_ovf)[0]]
        msg = f"sharded post-processing failed: ranks with errors {bad_ranks}, images with capacity overflow {bad_imgs[:16]}" \
              f"{' ...' if len(bad_imgs) > 16 else ''} (overflow bits 0x{int(np.bitwise_or.reduce(g_ovf)) if len(g_ovf) else 0:x})"
        if err is not None:
            raise ShardError(f"{msg}; this rank: {type(err).__name__}: {err}") from err
        raise ShardError(msg)
    return g_num, g_sub
