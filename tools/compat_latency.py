"""Latency of the reference operator surface (host pointers): process_paf + the getter loop, per call, with the
sparse upload (default) and with the whole paf_mat uploaded, next to the compiled reference on this box's CPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
import torch_ekpose_b200 as ek
from torch_ekpose_b200 import synthetic

fe = oracle.Frontend()
ref = oracle.RefPaf() if oracle.have_ref() else oracle.PortPaf()

def getters(mod):
    n = mod.get_num_humans()
    for h in range(n):
        for p in range(18):
            c = mod.get_part_cid(h, p)
            if c >= 0:
                mod.get_part_x(c); mod.get_part_y(c); mod.get_part_score(c)
        mod.get_score(h)
    return n

def run(label, h, w, people, seed):
    heat, paf = synthetic.make_batch(1, h, w, people, seed=seed)
    hw = np.ascontiguousarray(heat[0].transpose(1, 2, 0)); pw = np.ascontiguousarray(paf[0].transpose(1, 2, 0))
    peaks = fe.ref_nms(hw)[None]
    paf_up = fe.upsample_nearest(pw)
    heat_up = np.zeros((8 * h, 8 * w, 19), np.float32)
    out = []
    for mode in ("listed", "sparse", "dense"):
        os.environ["EKP_PROCESS_PAF_UPLOAD"] = mode
        for _ in range(3): ek.pafprocess.process_paf(peaks, heat_up, paf_up)
        t = time.perf_counter(); reps = 20
        for _ in range(reps):
            ek.pafprocess.process_paf(peaks, heat_up, paf_up); n = getters(ek.pafprocess)
        out.append((mode, (time.perf_counter() - t) / reps * 1e3, n))
    t = time.perf_counter(); reps = 5
    for _ in range(reps):
        sub, _ = oracle.subset_of(ref, peaks[0], 8 * h, 8 * w, paf_up)
    tref = (time.perf_counter() - t) / reps * 1e3
    # the library calls alone (raw C entry points, no Python wrapper, no getter loop) on both sides
    import ctypes as C
    from torch_ekpose_b200 import _lib
    os.environ["EKP_PROCESS_PAF_UPLOAD"] = "listed"
    pk32 = np.ascontiguousarray(peaks, np.float32)
    args_ours = (pk32.shape[0], pk32.shape[1], pk32.shape[2], pk32.ctypes.data, 8 * h, 8 * w, 19, None, paf_up.shape[0], paf_up.shape[1], paf_up.shape[2], paf_up.ctypes.data)
    for _ in range(5): _lib.lib.process_paf(*args_ours)
    t = time.perf_counter(); reps = 50
    for _ in range(reps): _lib.lib.process_paf(*args_ours)
    raw_ours = (time.perf_counter() - t) / reps * 1e3
    t = time.perf_counter()
    for _ in range(reps): ref._process(pk32.shape[0], pk32.shape[1], pk32.shape[2], pk32, 8 * h, 8 * w, 19, None, paf_up.shape[0], paf_up.shape[1], paf_up.shape[2], paf_up)
    raw_ref = (time.perf_counter() - t) / reps * 1e3
    t = time.perf_counter()
    for _ in range(reps): getters(ek.pafprocess)
    t_get = (time.perf_counter() - t) / reps * 1e3
    print(f"{label}: raw C call process_paf: ours {raw_ours:.3f} ms | compiled reference {raw_ref:.3f} ms | Python getter loop (either module) {t_get:.3f} ms")
    # route 2: the whole of paf_to_pose_cpp (numpy HWC in, list[Human] out) on the GPU, one image per call
    for _ in range(3): ek.paf_to_pose_cpp(hw, pw, ek.cfg)
    t = time.perf_counter(); reps = 20
    for _ in range(reps): humans = ek.paf_to_pose_cpp(hw, pw, ek.cfg)
    t2 = (time.perf_counter() - t) / reps * 1e3
    t = time.perf_counter(); reps = 5
    for _ in range(reps):
        pk = fe.ref_nms(hw); up = fe.upsample_nearest(pw); fe.upsample_nearest(hw); oracle.subset_of(ref, pk, 8 * h, 8 * w, up)
    t2ref = (time.perf_counter() - t) / reps * 1e3
    print(f"{label}: paf_to_pose_cpp one image: ours {t2:.3f} ms ({len(humans)} humans) | CPU path (C restatement of NMS + nearest x8 + "
          f"reference process_paf) {t2ref:.3f} ms")
    print(f"{label}: {peaks.shape[1]} peaks, {out[0][2]} humans | ours listed on the host (default) {out[0][1]:.3f} ms, listed by a kernel {out[1][1]:.3f} ms, whole tensor {out[2][1]:.3f} ms "
          f"| reference C++ on this CPU {tref:.3f} ms (process_paf only)")

run("368x432, 3 people", 46, 54, (3, 3), 1)
run("656x368, 8 people", 46, 82, (8, 8), 2)
run("1312x736, 35 people", 92, 164, (35, 35), 3)
