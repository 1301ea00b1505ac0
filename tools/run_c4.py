"""Minimal launch loop of the crowded configuration (1312x736, 30-40 people, 16 images) for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_ekpose_b200 as ek
from torch_ekpose_b200 import synthetic
heat, paf = synthetic.make_batch(16, 92, 164, (30, 40), seed=7)
hd, pd = torch.from_numpy(heat).cuda(), torch.from_numpy(paf).cuda()
pp = ek.PostProcessor(device=0, max_batch=16, max_h=92, max_w=164, max_peaks=2048, max_humans=128)
lean = len(sys.argv) > 1 and sys.argv[1] == "lean"
for _ in range(4):
    pp.run(hd, pd, frontend="dense", materialize=not lean)
res = pp.results()
print("ok", res["num_humans"].mean())
