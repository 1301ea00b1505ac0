"""Time the front-end kernel alone under different settings (which phase costs what)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_ekpose_b200 as ek
from torch_ekpose_b200 import synthetic
heat, paf = synthetic.make_batch(64, 46, 54, (1, 6), seed=100)
hd, pd = torch.from_numpy(heat).cuda(), torch.from_numpy(paf).cuda()
pp = ek.PostProcessor(device=0, max_batch=64, max_h=46, max_w=54, max_peaks=1024, max_humans=32)
def t(label, **kw):
    for _ in range(5): pp.run(hd, pd, **kw)
    torch.cuda.synchronize(); pp.set_timing(True)
    for _ in range(40): pp.run(hd, pd, **kw)
    torch.cuda.synchronize(); st, n = pp.stage_times(); pp.set_timing(False)
    print(f"{label:48s} frontend {st['frontend']*1000:8.1f} us   sort {st['peak_sort']*1000:6.1f} connect {st['connect']*1000:6.1f} assemble {st['assemble']*1000:6.1f}")
t("dense + materialise (bench config)", frontend="dense", materialize=True)
t("dense + materialise, thr=1e9 (all NMS culled)", frontend="dense", materialize=True, thr=1e9)
t("dense lean (no materialise)", frontend="dense", materialize=False)
t("dense lean, thr=1e9", frontend="dense", materialize=False, thr=1e9)
t("dense + materialise, thr=0.08 (more strips pass the cull)", frontend="dense", materialize=True, thr=0.08)
t("reference front-end lean", frontend="reference", materialize=False)
t("reference front-end + nearest materialise", frontend="reference", materialize=True)
