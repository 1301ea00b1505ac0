"""Per-phase times inside paf_connect_kernel (needs a library built with -DEKP_CONN_PROFILE:
VARIANT_SRC=paf_connect.cu tools/build_variants.sh prof "-DEKP_CONN_PROFILE"; EKPOSE_B200_SO=build/variants/prof.so)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_ekpose_b200 as ek
from torch_ekpose_b200 import synthetic, _lib
lib = ctypes.CDLL(_lib.SO_PATH)
names = ["stage peaks", "pass 1", "pass 2", "rank sort", "std::sort replay", "greedy"]
def run(label, n, h, w, people, frontend, materialize):
    heat, paf = synthetic.make_batch(n, h, w, people, seed=7)
    hd, pd = torch.from_numpy(heat).cuda(), torch.from_numpy(paf).cuda()
    pp = ek.PostProcessor(device=0, max_batch=n, max_h=h, max_w=w, max_peaks=2048, max_humans=128)
    for _ in range(3): pp.run(hd, pd, frontend=frontend, materialize=materialize)
    pp.results()
    buf = (ctypes.c_ulonglong * 16)()
    lib.ekp_debug_conn_profile(buf, 1)
    pp.run(hd, pd, frontend=frontend, materialize=materialize); pp.results()
    lib.ekp_debug_conn_profile(buf, 1)
    nb = n * 19
    print(f"{label}: blocks {nb}, candidates/limb mean {buf[6]/nb:.1f} max {buf[14]}, replays {buf[7]}")
    for k, nm in enumerate(names):
        print(f"   {nm:18s} mean {buf[k]/nb/1e3:8.2f} us   slowest block {buf[8+k]/1e3:8.2f} us")
    pp.close()
run("C4 crowded dense lean", 16, 92, 164, (30, 40), "dense", False)
run("C4 crowded dense mat", 16, 92, 164, (30, 40), "dense", True)
run("C4 crowded reference lean", 16, 92, 164, (30, 40), "reference", False)
run("C3 dense lean", 256, 46, 82, (2, 8), "dense", False)
run("C2 dense mat", 64, 46, 54, (1, 6), "dense", True)
