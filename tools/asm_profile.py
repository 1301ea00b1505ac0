"""Per-phase times inside assemble_kernel (needs a library built with -DEKP_ASM_PROFILE:
VARIANT_SRC=assemble.cu tools/build_variants.sh prof "-DEKP_ASM_PROFILE"; EKPOSE_B200_SO=build/variants/prof.so)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_ekpose_b200 as ek
from torch_ekpose_b200 import synthetic, _lib
lib = ctypes.CDLL(_lib.SO_PATH)
def run(label, n, h, w, people, frontend, materialize):
    heat, paf = synthetic.make_batch(n, h, w, people, seed=7)
    hd, pd = torch.from_numpy(heat).cuda(), torch.from_numpy(paf).cuda()
    pp = ek.PostProcessor(device=0, max_batch=n, max_h=h, max_w=w, max_peaks=2048, max_humans=128)
    for _ in range(3): pp.run(hd, pd, frontend=frontend, materialize=materialize)
    pp.results()
    buf = (ctypes.c_ulonglong * 8)()
    lib.ekp_debug_asm_profile(buf, 1)
    pp.run(hd, pd, frontend=frontend, materialize=materialize); pp.results()
    lib.ekp_debug_asm_profile(buf, 1)
    print(f"{label}: per image: staging {buf[0]/n/1e3:.2f} us, limbs tried in parallel {buf[1]/n/1e3:.2f} us, sequential limbs {buf[2]/n/1e3:.2f} us, "
          f"prune+record {buf[3]/n/1e3:.2f} us; limbs parallel {buf[4]/n:.1f} sequential {buf[5]/n:.1f}; "
          f"slowest image {buf[6]/1e3:.1f} us, most sequential limbs in one image {buf[7]}")
    pp.close()
run("C4 crowded dense lean", 16, 92, 164, (30, 40), "dense", False)
run("C4 crowded reference lean", 16, 92, 164, (30, 40), "reference", False)
run("C3 dense lean", 256, 46, 82, (2, 8), "dense", False)
run("C2 dense mat", 64, 46, 54, (1, 6), "dense", True)
