"""Minimal launch loop of the bench configuration (64 x 368x432, 1-6 people) for ncu."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_ekpose_b200 as ek
from torch_ekpose_b200 import synthetic
heat, paf = synthetic.make_batch(64, 46, 54, (1, 6), seed=100)
hd, pd = torch.from_numpy(heat).cuda(), torch.from_numpy(paf).cuda()
pp = ek.PostProcessor(device=0, max_batch=64, max_h=46, max_w=54, max_peaks=1024, max_humans=32)
lean = len(sys.argv) > 1 and sys.argv[1] == "lean"
for _ in range(4):
    pp.run(hd, pd, frontend="dense", materialize=not lean)
res = pp.results()
print("ok", res["num_humans"].mean())
