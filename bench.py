#!/usr/bin/env python
"""bench.py -- the BASELINE.json metric: PAF post-processing images/sec at 368x432.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the whole hot path (stages 1-5: bilinear x8 upsample materialising the
operator-surface tensors, Gaussian smoothing, 3x3 NMS peak extraction, PAF line-integral scoring,
greedy limb assignment + person assembly) over one batch of 64 synthetic 368x432 images per GPU
(BASELINE.json configs[1]).  Images shard across GPUs with no data-path collective (weak scaling:
every rank processes its own 64-image batch per step); the only exchange is one final result
gather.  One JSON line is printed by rank 0.

  value      images/s, whole job, inputs resident in HBM, timed with CUDA events (max over ranks)
  e2e        images/s through the host-buffer C-ABI entry (pinned host -> H2D -> kernels -> D2H of
             the result tables -> host arrays), copies inside the timed region
  roofline   the fused stage 1-3 kernel: algorithmic bytes (SURVEY.md 8d materialising contract)
             / its mean launch duration from CUDA events recorded by the library on its stream
  cpu_baseline  the reference's CPU path on ONE host core on a bounded sample (rank 0, N=1 only)

--impl reference times the reference's own CPU implementation of the same path (stride-8 NMS +
bicubic refinement, nearest x8 upsample of PAF and heat, process_paf) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "postprocess images/sec at 368x432"
UNIT = "images/s"
BATCH = 64            # images per GPU per step (configs[1])
H_LO, W_LO = 46, 54   # stride-8 map of a 368x432 image
PEOPLE = (1, 6)
INPUT_SETS = 4        # distinct input batches rotated between steps
N_CTX = int(os.environ.get("EKP_BENCH_CONTEXTS", "4"))   # contexts / CUDA streams the steps rotate over
ALGO_BYTES_PER_IMAGE = 4 * H_LO * W_LO * 57 + 4 * (8 * H_LO) * (8 * W_LO) * 57   # 36,812,880
CONFIG = {
    "workload": "configs[1]: batch 64 synthetic heat(19ch)/PAF(38ch) at 46x54 stride-8 (368x432), "
                "dense front-end materialising heat_mat/paf_mat + PAF scoring + assembly",
    "batch_per_gpu": BATCH, "shape": "368x432", "people_per_image": "1-6", "frontend": "dense",
    "materialize": True,
    "pipelining": f"{N_CTX} contexts on {N_CTX} CUDA streams take the steps in turn (stages 4-5 of one batch overlap the "
                  "front-end kernel of the next)",
    "l2": "every step writes 2.3 GB of operator-surface tensors (>> 126 MB L2); inputs rotate over "
          f"{INPUT_SETS} distinct batches",
}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic():
    """Per-launch DRAM bytes of the fused kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("dense_frontend_kernel_dram_bytes_per_launch")
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------
# clocks: NVML sampled in a thread DURING the timed regions (the same counters nvidia-smi prints)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._active = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            if self._active.is_set():
                try:
                    mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                    try:
                        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    self.samples.append(mhz)
                    for bit, name in self.REASONS.items():
                        if r & bit and name != "gpu_idle":
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.004)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def region(self, on: bool):
        (self._active.set if on else self._active.clear)()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=1)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------
# the reference's CPU path (the checker code under oracle/, timed -- never shipped)
# ---------------------------------------------------------------------------------------------
def _cpu_one_image(args):
    """paf_to_pose_cpp for one image on the CPU: NMS + nearest x8 of PAF and heat + process_paf +
    getter loop (paf_to_pose.py:346-378)."""
    heat_hwc, paf_hwc = args
    import oracle
    g = _cpu_one_image.__dict__
    if "fe" not in g:
        g["fe"] = oracle.Frontend()
        g["paf"] = oracle.RefPaf() if oracle.have_ref() else oracle.PortPaf()
    fe, impl = g["fe"], g["paf"]
    peaks = fe.ref_nms(heat_hwc, np.float32(0.15))
    if len(peaks) == 0:
        return 0
    paf_up = fe.upsample_nearest(paf_hwc)
    heat_up = fe.upsample_nearest(heat_hwc)
    impl.process_paf(peaks[None], heat_up, paf_up)
    n = impl.get_num_humans()
    for hid in range(n):
        for part in range(18):
            cid = impl.get_part_cid(hid, part)
            if cid >= 0:
                impl.get_part_x(cid), impl.get_part_y(cid), impl.get_part_score(cid)
        impl.get_score(hid)
    return n


def cpu_path_kind():
    import oracle
    return "reference" if oracle.have_ref() else "port"


def cpu_baseline_sample(images_hwc, budget_s=12.0):
    """One host core, bounded: loops over the sample images until ~budget_s of CPU work is done."""
    t0 = time.perf_counter()
    done = 0
    while True:
        for hw in images_hwc:
            _cpu_one_image(hw)
            done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done / dt, done, dt


def hwc_images(heat_nchw, paf_nchw, count):
    return [(np.ascontiguousarray(heat_nchw[i].transpose(1, 2, 0)), np.ascontiguousarray(paf_nchw[i].transpose(1, 2, 0)))
            for i in range(count)]


def run_reference_arm(args, rank, world):
    """The reference's own CPU implementation on all host cores (rank 0 only)."""
    if rank != 0:
        return
    import multiprocessing as mp

    import oracle
    from torch_ekpose_b200 import synthetic
    oracle.build()
    heat, paf = synthetic.make_batch(BATCH, H_LO, W_LO, PEOPLE, seed=100)
    imgs = hwc_images(heat, paf, BATCH)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        # calibrate, then bound the per-step sample so that warmup + steps end within ~2 minutes
        pool.map(_cpu_one_image, imgs[:cores], chunksize=1)
        t0 = time.perf_counter()
        pool.map(_cpu_one_image, imgs, chunksize=max(1, BATCH // (cores * 2)))
        rate = BATCH / (time.perf_counter() - t0)
        per_step = int(max(cores, min(BATCH, 100.0 * rate / max(args.steps + args.warmup, 1))))
        sample = imgs[:per_step]
        chunk = max(1, per_step // (cores * 2))
        for _ in range(args.warmup):
            pool.map(_cpu_one_image, sample, chunksize=chunk)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_one_image, sample, chunksize=chunk)
        dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    kind = cpu_path_kind()
    sample = (f"{args.steps} steps x {per_step} images of configs[1] on {cores} processes; front-end = C restatement of the "
              f"reference's Python NMS (oracle/frontend_oracle.c; the Python original cannot leave the authoring "
              f"container), nearest x8 upsample of PAF+heat, process_paf = "
              f"{'unmodified reference C++ (oracle/_ref)' if kind == 'reference' else 'C port (oracle/paf_oracle.c)'}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": CONFIG,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    import torch_ekpose_b200 as ek
    from torch_ekpose_b200 import synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL may print its version banner on stdout when the communicator is created; stdout must
        # carry exactly one JSON line, so fd 1 points at stderr while the communicator comes up.
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    # synthetic inputs: INPUT_SETS distinct batches per rank, resident in HBM and in pinned host memory
    sets_dev, sets_pin = [], []
    for s in range(INPUT_SETS):
        heat, paf = synthetic.make_batch(BATCH, H_LO, W_LO, PEOPLE, seed=100 + 17 * rank + s)
        ht, pt = torch.from_numpy(heat), torch.from_numpy(paf)
        sets_pin.append((ht.pin_memory(), pt.pin_memory()))
        sets_dev.append((ht.to(dev), pt.to(dev)))
        if s == 0:
            sample_imgs = hwc_images(heat, paf, 8)
    # N_CTX contexts on N_CTX streams take the steps in turn, so the small latency-bound kernels of
    # stages 4-5 of batch i run under the HBM-bound front-end kernel of batch i+1 (and with three, the
    # front-end of batch i+2 does not queue behind stages 4-5 of batch i on the same stream).
    mk = lambda: ek.PostProcessor(device=local_rank, max_batch=BATCH, max_h=H_LO, max_w=W_LO, max_peaks=1024, max_humans=32)
    pps = [mk() for _ in range(N_CTX)]
    pp = pps[0]
    stream = torch.cuda.current_stream(dev)
    streams = [torch.cuda.Stream(dev) for _ in range(N_CTX)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_device(i):
        hd, pd = sets_dev[i % INPUT_SETS]
        pps[i % N_CTX].run(hd, pd, layout="nchw", frontend="dense", materialize=True, stream=streams[i % N_CTX])

    clocks = ClockSampler(local_rank)
    clocks.start()

    # ---- value: device-resident inputs ------------------------------------------------------------
    for i in range(max(args.warmup, 3) + 1):
        step_device(i)
    for p_ in pps:
        p_.results()
        p_.set_timing(True)
    launches0 = sum(p_.kernel_launches() for p_ in pps)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    clocks.region(True)
    ev0.record(stream)
    for s_ in streams:
        s_.wait_stream(stream)
    for i in range(args.steps):
        step_device(i)
    for s_ in streams:
        stream.wait_stream(s_)
    ev1.record(stream)
    barrier()
    clocks.region(False)
    ms = ev0.elapsed_time(ev1)
    res = pps[(args.steps - 1) % N_CTX].results()
    launches = sum(p_.kernel_launches() for p_ in pps) - launches0
    per_ctx = [p_.stage_times() for p_ in pps[:max(1, min(N_CTX, args.steps))]]
    stage_runs = sum(r for _, r in per_ctx)
    stage_ms = {k: sum(st[k] * r for st, r in per_ctx) / max(stage_runs, 1) for k in per_ctx[0][0]}
    for p_ in pps:
        p_.set_timing(False)
    total_humans = int(res["num_humans"].sum())

    # the same steps on ONE context / stream (nothing overlaps): per-stage device times in isolation
    torch.cuda.synchronize(dev)
    pp.set_timing(True)
    for i in range(32):
        hd, pd = sets_dev[i % INPUT_SETS]
        pp.run(hd, pd, layout="nchw", frontend="dense", materialize=True, stream=streams[0])
    pp.results()
    iso_ms, _ = pp.stage_times()
    pp.set_timing(False)

    # ---- context numbers (not the headline): the same batch without materialising the operator-surface
    #      tensors, and through the reference's own front-end (stride-8 NMS + bicubic refinement)
    def variant(frontend, materialize, steps=100):
        for i in range(2 * N_CTX):
            pps[i % N_CTX].run(*sets_dev[i % INPUT_SETS], layout="nchw", frontend=frontend, materialize=materialize, stream=streams[i % N_CTX])
        torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for s_ in streams:
            s_.wait_stream(stream)
        for i in range(steps):
            pps[i % N_CTX].run(*sets_dev[i % INPUT_SETS], layout="nchw", frontend=frontend, materialize=materialize, stream=streams[i % N_CTX])
        for s_ in streams:
            stream.wait_stream(s_)
        b.record(stream)
        torch.cuda.synchronize(dev)
        return BATCH * steps / (a.elapsed_time(b) / 1000.0)

    variants = {"dense_frontend_no_materialise_images_per_s": variant("dense", False),
                "reference_frontend_no_materialise_images_per_s": variant("reference", False)}

    # ---- e2e: host buffers through the C ABI, copies in the timed region ---------------------------
    # A stream of batches the way a caller would drive it: N_CTX contexts on N_CTX streams, so the H2D
    # copy of batch i+1 overlaps the kernels of batch i; EVERY step's inputs come from pinned host
    # memory and EVERY step's result tables are read back on the host inside the timed region.
    def e2e_submit(i):
        hp, ppin = sets_pin[i % INPUT_SETS]
        pps[i % N_CTX].run(hp, ppin, layout="nchw", frontend="dense", materialize=True, stream=streams[i % N_CTX])

    def e2e_loop(steps):
        got = None
        for i in range(steps):
            if i >= N_CTX:
                got = pps[i % N_CTX].human_tables()      # results of step i-N_CTX (waits for it)
            e2e_submit(i)
        for i in range(max(steps - N_CTX, 0), steps):
            got = pps[i % N_CTX].human_tables()
        return got

    e2e_loop(2 * N_CTX)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    clocks.region(True)
    t_wall = time.perf_counter()
    e0.record(stream)
    for s_ in streams:
        s_.wait_stream(stream)
    num, parts, scores = e2e_loop(args.steps)
    for s_ in streams:
        stream.wait_stream(s_)
    e1.record(stream)
    barrier()
    clocks.region(False)
    e2e_wall_ms = 1000.0 * (time.perf_counter() - t_wall)
    e2e_ms = max(e0.elapsed_time(e1), 0.0)
    clk = clocks.stop()
    h2d = BATCH * H_LO * W_LO * 57 * 4
    # what every run copies back: one packed record per image (header + subset rows + per-human part table + scores)
    d2h = int(BATCH * (((16 + (80 + 16 * 18 + 4) * pp.max_humans) + 15) // 16) * 16)

    # ---- max over ranks, final result gather -------------------------------------------------------
    t = torch.tensor([ms, e2e_ms, e2e_wall_ms, float(total_humans)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros(BATCH, dtype=torch.int32, device=dev) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(res["num_humans"]).to(dev))   # the final result gather
        total_humans = int(sum(int(g.sum()) for g in gathered))
        ms, e2e_ms, e2e_wall_ms = float(tmax[0]), float(tmax[1]), float(tmax[2])

    if rank == 0:
        peak, peak_src = measured_peak()
        images = BATCH * world * args.steps
        value = images / (ms / 1000.0)
        # Roofline of the dominant kernel (fused stages 1-3).  The steps are pipelined over N_CTX streams,
        # so per-kernel event intervals overlap each other; the honest per-launch duration over the
        # timed region is (timed region) / (launches) = ms_per_step, which also charges the kernel for
        # everything else in the step (a lower bound on its bandwidth).  kernel_ms_isolated is the same
        # kernel timed alone (events recorded by the library around it, no other stream active).
        per_launch_ms = ms / args.steps
        achieved = (ALGO_BYTES_PER_IMAGE * BATCH) / (per_launch_ms / 1000.0) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": CONFIG,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(), "kernel": "dense_frontend_kernel (stages 1-3 fused)",
                         "kernel_ms": per_launch_ms, "kernel_ms_isolated": iso_ms["frontend"],
                         "frac_isolated": (ALGO_BYTES_PER_IMAGE * BATCH) / (iso_ms["frontend"] / 1000.0) / 1e9 / peak if iso_ms["frontend"] > 0 else None,
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_IMAGE * BATCH,
                         "peak_source": peak_src, "launches_averaged": args.steps},
            "stage_ms_isolated": iso_ms, "stage_ms_pipelined": stage_ms,
            "e2e": {"value": images / (max(e2e_ms, e2e_wall_ms) / 1000.0), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "device_ms_per_step": e2e_ms / args.steps,
                    "wall_ms_per_step": e2e_wall_ms / args.steps,
                    "api": f"ekp_postprocess_host + ekp_results_humans (pinned host buffers; {N_CTX} contexts / {N_CTX} streams "
                           "so the H2D of one batch overlaps the kernels of the previous one)"},
            "gpu_launches": int(launches), "clocks": clk, "humans_found_last_step": total_humans,
            "variants_per_gpu": variants,
        }
        if world == 1 and not args.no_cpu_baseline:
            import oracle
            oracle.build()
            v, done, dt = cpu_baseline_sample(sample_imgs, budget_s=args.cpu_seconds)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": 1, "kind": cpu_path_kind(),
                "sample": f"{done} images ({dt:.1f} s) of the same synthetic 368x432 workload through the reference's CPU path: "
                          "stride-8 NMS + bicubic refinement (C restatement), nearest x8 upsample of PAF+heat, "
                          "reference process_paf + getter loop; 1 core"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    for p_ in pps:
        p_.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # single-process launch asked for N GPUs: re-launch under torchrun as the driver would
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__), "--gpus", str(args.gpus),
               "--steps", str(args.steps), "--warmup", str(args.warmup)]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
