"""Minimal launch loop of one BASELINE.json configuration for ncu (bench.py's capacities).
usage: python tools/run_cfg.py c2|c3|c4 lean|mat [dense|reference]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_ekpose_b200 as ek
from torch_ekpose_b200 import synthetic
cfg, mode = sys.argv[1], sys.argv[2]
frontend = sys.argv[3] if len(sys.argv) > 3 else "dense"
n, h, w, people, caps = {"c2": (64, 46, 54, (1, 6), (1024, 32, 64, 512)), "c3": (256, 46, 82, (2, 8), (1024, 32, 64, 512)),
                         "c4": (16, 92, 164, (30, 40), (2048, 128, 128, 1024))}[cfg]
heat, paf = synthetic.make_batch(n, h, w, people, seed={"c2": 100, "c3": 300, "c4": 400}[cfg])
hd, pd = torch.from_numpy(heat).cuda(), torch.from_numpy(paf).cuda()
pp = ek.PostProcessor(device=0, max_batch=n, max_h=h, max_w=w, max_peaks=caps[0], max_humans=caps[1], max_part=caps[2], max_cand=caps[3])
pp.set_timing(True)   # eager launches (no CUDA graph), so that ncu sees plain kernel launches
for _ in range(4):
    pp.run(hd, pd, frontend=frontend, materialize=mode == "mat")
res = pp.results()
print("ok", cfg, mode, frontend, float(res["num_humans"].mean()))
