"""Stage times for the other BASELINE.json configurations (not bench lines; context for DESIGN.md)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_ekpose_b200 as ek
from torch_ekpose_b200 import synthetic

def run(label, n, h, w, people, frontend, materialize, max_peaks=2048, max_humans=128):
    heat, paf = synthetic.make_batch(n, h, w, people, seed=7)
    hd, pd = torch.from_numpy(heat).cuda(), torch.from_numpy(paf).cuda()
    pp = ek.PostProcessor(device=0, max_batch=n, max_h=h, max_w=w, max_peaks=max_peaks, max_humans=max_humans)
    for _ in range(3): pp.run(hd, pd, frontend=frontend, materialize=materialize)
    res = pp.results(); pp.set_timing(True)
    for _ in range(20): pp.run(hd, pd, frontend=frontend, materialize=materialize)
    pp.results(); st, _ = pp.stage_times(); pp.set_timing(False)
    tot = sum(st.values())
    algo = 4 * h * w * 57 * 65 * n if materialize else 4 * h * w * 57 * n
    print(f"{label:44s} n={n:3d} humans/img={res['num_humans'].mean():5.1f} peaks/img={res['n_peaks'].mean():6.1f} | frontend {st['frontend']*1e3:8.1f} sort {st['peak_sort']*1e3:7.1f} connect {st['connect']*1e3:8.1f} assemble {st['assemble']*1e3:7.1f} us | {n/tot*1e3:9.0f} img/s | front-end {algo/st['frontend']/1e6:7.0f} GB/s")
    pp.close()

run("C2 368x432 dense+mat", 64, 46, 54, (1, 6), "dense", True)
run("C3 656x368 dense+mat", 256, 46, 82, (2, 8), "dense", True)
run("C3 656x368 dense lean", 256, 46, 82, (2, 8), "dense", False)
run("C3 656x368 reference lean", 256, 46, 82, (2, 8), "reference", False)
run("C4 1312x736 crowded dense+mat", 16, 92, 164, (30, 40), "dense", True)
run("C4 1312x736 crowded dense lean", 16, 92, 164, (30, 40), "dense", False)
run("C4 1312x736 crowded reference lean", 16, 92, 164, (30, 40), "reference", False)
