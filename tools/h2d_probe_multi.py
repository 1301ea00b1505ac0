"""Concurrent pinned host -> device bandwidth of the box, all ranks at once (the ceiling of bench.py's e2e number).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 tools/h2d_probe_multi.py

Every rank copies the bench's per-batch input (one 36.2 MB block = heat|paf of 64 images at 368x432, or the same bytes
as the 12.1 + 24.2 MB pair round 1 used) from pinned host memory to its own GPU on 4 streams, all ranks between the
same two barriers; variants: host buffers from torch's pin_memory(), from cudaHostAlloc (default / write-combined),
with and without pinning the rank to its own slice of the host cores before allocating.  Rank 0 prints one JSON line
per variant with per-GPU and aggregate GB/s (device-timed, max over ranks)."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    sys.stdout.flush()
    fd = os.dup(1); os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier(); torch.cuda.synchronize(dev)
    sys.stdout.flush(); os.dup2(fd, 1); os.close(fd)

from torch_ekpose_b200 import _lib

HEAT, PAF = 64 * 46 * 54 * 19 * 4, 64 * 46 * 54 * 38 * 4
STREAMS, REPS = 4, 16
all_cores = sorted(os.sched_getaffinity(0))


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def host_buffer(kind, nbytes):
    if kind == "torch_pin":
        return torch.empty(nbytes, dtype=torch.uint8).pin_memory(), None
    p = ctypes.c_void_p()
    _lib.check(_lib.lib.ekp_host_alloc(ctypes.byref(p), nbytes, 1 if kind == "cuda_wc" else 0))
    arr = np.frombuffer((ctypes.c_uint8 * nbytes).from_address(p.value), dtype=np.uint8)
    arr[:] = 1
    return torch.from_numpy(arr), p


def measure(kind, sizes, pin_cores):
    if pin_cores and world > 1:
        per = max(1, len(all_cores) // world)
        os.sched_setaffinity(0, all_cores[local * per:(local + 1) * per] or all_cores)
    else:
        os.sched_setaffinity(0, all_cores)
    hs, keep = [], []
    for _ in range(STREAMS):
        row = []
        for s in sizes:
            t, p = host_buffer(kind, s)
            row.append(t); keep.append(p)
        hs.append(row)
    ds = [[torch.empty(s, dtype=torch.uint8, device=dev) for s in sizes] for _ in range(STREAMS)]
    st = [torch.cuda.Stream(dev) for _ in range(STREAMS)]
    main = torch.cuda.current_stream(dev)

    def go(reps):
        for _ in range(reps):
            for k in range(STREAMS):
                with torch.cuda.stream(st[k]):
                    for h, d in zip(hs[k], ds[k]):
                        d.copy_(h, non_blocking=True)
    go(2)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(main)
    for s in st:
        s.wait_stream(main)
    go(REPS)
    for s in st:
        main.wait_stream(s)
    b.record(main)
    barrier()
    gbs = sum(sizes) * REPS * STREAMS / (a.elapsed_time(b) / 1e3) / 1e9
    t = torch.tensor([gbs], dtype=torch.float64, device=dev)
    allv = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(allv, t)
    else:
        allv = [t]
    del hs, ds
    for p in keep:
        if p is not None:
            _lib.lib.ekp_host_free(p)
    vals = [float(x[0]) for x in allv]
    if rank == 0:
        print(json.dumps({"n_gpus": world, "host_memory": kind, "copies_per_batch": len(sizes), "bytes_per_batch": sum(sizes),
                          "rank_pinned_to_own_cores": bool(pin_cores and world > 1), "host_cores": len(all_cores),
                          "gbs_per_gpu": [round(v, 2) for v in vals], "gbs_aggregate": round(sum(vals), 2),
                          "gbs_min_rank_x_n": round(min(vals) * world, 2)}), flush=True)


for kind in ("torch_pin", "cuda_default", "cuda_wc"):
    for sizes in ([HEAT + PAF], [HEAT, PAF]):
        for pin in ((False, True) if world > 1 else (False,)):
            measure(kind, sizes, pin)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
