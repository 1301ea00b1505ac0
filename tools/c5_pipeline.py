"""Context for BASELINE.json configs[4]: a cuDNN backbone of the reference's shape feeding the CUDA post-processing on
the same GPU (nothing leaves the device between the two).  The network here is a random-init stand-in with the public
OpenPose-2016 layout the reference's vgg2016 follows (VGG-19 front to conv4_2, two CPM convs, six two-branch stages with
38 PAF + 19 heat channels at stride 8); its weights and outputs are meaningless, so the post-processing is fed the
synthetic maps of the bench while the network's forward pass is timed next to it.
usage: python tools/c5_pipeline.py [batch] [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn as nn
import torch_ekpose_b200 as ek
from torch_ekpose_b200 import synthetic

def conv(i, o, k):
    return [nn.Conv2d(i, o, k, padding=k // 2), nn.ReLU(inplace=True)]

class Stage(nn.Module):
    def __init__(self, cin, k, n, mid, out):
        super().__init__()
        layers, c = [], cin
        for _ in range(n):
            layers += conv(c, 128, k); c = 128
        layers += conv(c, mid, 1) + [nn.Conv2d(mid, out, 1)]
        self.net = nn.Sequential(*layers)
    def forward(self, x):
        return self.net(x)

class Pose2016(nn.Module):
    def __init__(self):
        super().__init__()
        cfg = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512]
        layers, c = [], 3
        for v in cfg:
            if v == "M":
                layers.append(nn.MaxPool2d(2, 2))
            else:
                layers += conv(c, v, 3); c = v
        layers += conv(512, 256, 3) + conv(256, 128, 3)
        self.front = nn.Sequential(*layers)
        self.paf = nn.ModuleList([Stage(128, 3, 3, 512, 38)] + [Stage(185, 7, 5, 128, 38) for _ in range(5)])
        self.heat = nn.ModuleList([Stage(128, 3, 3, 512, 19)] + [Stage(185, 7, 5, 128, 19) for _ in range(5)])
    def forward(self, x):
        f = self.front(x)
        p, h = self.paf[0](f), self.heat[0](f)
        for s in range(1, 6):
            z = torch.cat([p, h, f], 1)
            p, h = self.paf[s](z), self.heat[s](z)
        return p, h

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
net = Pose2016().to(dev).eval().to(memory_format=torch.channels_last)
frames = torch.randint(0, 255, (batch, 368, 432, 3), dtype=torch.uint8, device=dev)
heat, paf = synthetic.make_batch(batch, 46, 54, (1, 6), seed=5)
hd, pd = torch.from_numpy(heat).to(dev), torch.from_numpy(paf).to(dev)
pp = ek.PostProcessor(device=0, max_batch=batch, max_h=46, max_w=54, max_peaks=1024, max_humans=32)

def step(dtype):
    x, _ = pp.preprocess(frames, mode="vgg", dest_size=432)   # padding + normalisation on the GPU (row f4); 368x432 network input
    with torch.no_grad(), torch.autocast("cuda", dtype=dtype, enabled=dtype is not None):
        p, h = net(x.contiguous(memory_format=torch.channels_last))
    assert p.shape[1:] == (38, 46, 54) and h.shape[1:] == (19, 46, 54)
    pp.run(hd, pd, frontend="reference")              # synthetic maps: the stand-in's outputs contain no people
    return pp

for name, dtype in (("fp32 (TF32 off)", None), ("bf16 autocast", torch.bfloat16)):
    for _ in range(3): step(dtype)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps): step(dtype)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    pp.set_timing(True); pp.run(hd, pd, frontend="reference"); pp.results(); st, _ = pp.stage_times(); pp.set_timing(False)
    post = sum(st.values())
    print(f"{name}: batch {batch}: {ms:.2f} ms per step = {batch / ms * 1e3:.0f} frames/s on one B200; post-processing (reference front-end, "
          f"stages 1-5) {post * 1e3:.0f} us of it = {100 * post / ms:.2f} %; humans in the last batch {int(pp.results()['num_humans'].sum())}")
