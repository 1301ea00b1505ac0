"""Per-phase times inside assemble_kernel (needs a library built with -DEKP_ASM_PROFILE:
VARIANT_SRC=assemble.cu tools/build_variants.sh asmprof "-DEKP_ASM_PROFILE"; EKPOSE_B200_SO=build/variants/asmprof.so)."""
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
    u = lambda k: buf[k] / n / 1e3
    print(f"{label}: per image (us): staging {u(0):.2f} | limbs total {u(1):.2f} = lookup+simple {u(3):.2f} + complex walk {u(4):.2f} + new rows {u(5):.2f} "
          f"| prune+record {u(2):.2f}; complex connections per image {buf[6]/n:.1f}, map rebuilds per image {buf[7]/n:.1f}")
    pp.close()
run("C4 crowded dense lean", 16, 92, 164, (30, 40), "dense", False)
run("C4 crowded reference lean", 16, 92, 164, (30, 40), "reference", False)
run("C3 dense lean", 256, 46, 82, (2, 8), "dense", False)
run("C2 dense mat", 64, 46, 54, (1, 6), "dense", True)
