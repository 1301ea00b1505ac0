"""Stage times for the BASELINE.json configurations with the capacities bench.py uses (context for DESIGN.md).
usage: python tools/time_configs.py [default_caps]   (default_caps: the library's default max_part / max_cand instead)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_ekpose_b200 as ek
from torch_ekpose_b200 import synthetic

DEFAULT_CAPS = len(sys.argv) > 1 and sys.argv[1] == "default_caps"

def run(label, n, h, w, people, frontend, materialize, max_peaks=1024, max_humans=32, max_part=64, max_cand=512):
    heat, paf = synthetic.make_batch(n, h, w, people, seed=7)
    hd, pd = torch.from_numpy(heat).cuda(), torch.from_numpy(paf).cuda()
    if DEFAULT_CAPS:
        max_part = max_cand = 0
    pp = ek.PostProcessor(device=0, max_batch=n, max_h=h, max_w=w, max_peaks=max_peaks, max_humans=max_humans, max_part=max_part, max_cand=max_cand)
    for _ in range(3): pp.run(hd, pd, frontend=frontend, materialize=materialize)
    res = pp.results(); pp.set_timing(True)
    for _ in range(20): pp.run(hd, pd, frontend=frontend, materialize=materialize)
    pp.results(); st, _ = pp.stage_times(); pp.set_timing(False)
    # the same batches as CUDA graph replays (what a stream of frames through the same buffers gets), whole batch latency
    for _ in range(3): pp.run(hd, pd, frontend=frontend, materialize=materialize)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): pp.run(hd, pd, frontend=frontend, materialize=materialize)
    b.record(); torch.cuda.synchronize()
    graph_us = a.elapsed_time(b) / 20 * 1e3
    tot = sum(st.values())
    algo = 4 * h * w * 57 * 65 * n if materialize else 4 * h * w * 57 * n
    print(f"{label:40s} n={n:3d} humans/img={res['num_humans'].mean():5.1f} peaks/img={res['n_peaks'].mean():6.1f} | frontend {st['frontend']*1e3:7.1f} sort {st['peak_sort']*1e3:6.1f} "
          f"connect {st['connect']*1e3:6.1f} assemble {st['assemble']*1e3:6.1f} us (4-5: {(tot - st['frontend'])*1e3:6.1f}) | one stream eager {n/tot*1e3:8.0f} img/s, graph {graph_us:7.1f} us/batch = "
          f"{n/graph_us*1e6:8.0f} img/s | front-end {algo/st['frontend']/1e6:6.0f} GB/s")
    pp.close()

run("C2 368x432 dense+mat", 64, 46, 54, (1, 6), "dense", True)
run("C2 368x432 dense lean", 64, 46, 54, (1, 6), "dense", False)
run("C2 368x432 reference lean", 64, 46, 54, (1, 6), "reference", False)
run("C3 656x368 dense+mat", 256, 46, 82, (2, 8), "dense", True)
run("C3 656x368 dense lean", 256, 46, 82, (2, 8), "dense", False)
run("C3 656x368 reference lean", 256, 46, 82, (2, 8), "reference", False)
run("C4 1312x736 crowded dense+mat", 16, 92, 164, (30, 40), "dense", True, 2048, 128, 128, 1024)
run("C4 1312x736 crowded dense lean", 16, 92, 164, (30, 40), "dense", False, 2048, 128, 128, 1024)
run("C4 1312x736 crowded reference lean", 16, 92, 164, (30, 40), "reference", False, 2048, 128, 128, 1024)
