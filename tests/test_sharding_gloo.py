"""CPU: the N>1 path (partition + final result gather) with world_size 2 over gloo.

The GPU compute is replaced by the oracle so that only the host-side sharding logic is under test:
the gathered tables must equal the single-process result image for image."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _oracle_compute(heat, paf, max_humans=32):
    import oracle
    fe, port = oracle.Frontend(), oracle.PortPaf()
    n = heat.shape[0]
    num = np.zeros(n, np.int32)
    sub = np.zeros((n, max_humans, 20), np.float32)
    for i in range(n):
        hw = np.ascontiguousarray(heat[i].transpose(1, 2, 0))
        pw = np.ascontiguousarray(paf[i].transpose(1, 2, 0))
        peaks = fe.ref_nms(hw)
        s, _ = oracle.subset_of(port, peaks, hw.shape[0] * 8, hw.shape[1] * 8, fe.upsample_nearest(pw))
        num[i] = len(s)
        sub[i, :len(s)] = s
    return num, sub


def _worker(rank, world, port, n_images, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from torch_ekpose_b200 import synthetic
    from torch_ekpose_b200.sharding import postprocess_sharded, shard_bounds
    heat, paf = synthetic.make_batch(n_images, 46, 54, (1, 3), seed=5)
    calls = []

    def compute(hs, ps):
        calls.append(hs.shape[0])
        return _oracle_compute(hs, ps)

    num, sub = postprocess_sharded(heat, paf, max_humans=32, compute=compute)
    lo, hi = shard_bounds(n_images, world, rank)
    assert calls == ([hi - lo] if hi > lo else [])
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), num=num, sub=sub)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n_images", [5, 1])
def test_sharded_equals_single_process(tmp_path, n_images):
    from torch_ekpose_b200 import synthetic
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_images, str(tmp_path)), nprocs=world, join=True)
    heat, paf = synthetic.make_batch(n_images, 46, 54, (1, 3), seed=5)
    want_num, want_sub = _oracle_compute(heat, paf)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(z["num"], want_num)
        assert np.array_equal(z["sub"].view(np.uint32), want_sub.view(np.uint32))
    assert want_num.sum() >= n_images


def _failing_worker(rank, world, port, mode, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from torch_ekpose_b200 import synthetic
    from torch_ekpose_b200.sharding import ShardError, postprocess_sharded
    heat, paf = synthetic.make_batch(4, 46, 54, (1, 2), seed=6)

    def compute(hs, ps):
        num, sub = _oracle_compute(hs, ps)
        if mode == "raise" and rank == 1:
            raise RuntimeError("device lost on this rank")
        ovf = np.zeros(len(num), np.uint32)
        if mode == "overflow" and rank == 0:
            ovf[1] = 4   # EKP_OVF_CANDIDATES on global image 1
        return num, sub, ovf

    try:
        postprocess_sharded(heat, paf, max_humans=32, compute=compute)
        outcome = "returned"
    except ShardError as e:
        outcome = str(e)
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write(outcome)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["raise", "overflow"])
def test_a_failing_rank_fails_every_rank_after_the_gather(tmp_path, mode):
    """A rank whose shard raised (or overflowed a capacity) must not leave the others waiting in the collective:
    the status and overflow bits travel with the tables and EVERY rank raises afterwards."""
    world = 2
    mp.spawn(_failing_worker, args=(world, _free_port(), mode, str(tmp_path)), nprocs=world, join=True)
    texts = [open(tmp_path / f"rank{r}.txt").read() for r in range(world)]
    for t in texts:
        assert t != "returned" and "sharded post-processing failed" in t
    if mode == "raise":
        assert all("ranks with errors [1]" in t for t in texts) and "device lost" in texts[1] and "device lost" not in texts[0]
    else:
        assert all("capacity overflow [1]" in t and "0x4" in t for t in texts)
