"""Stage times on heavier crowds than BASELINE.json's configs[3]: 16 x 1312x736 with P people per image, reference front-end,
nothing materialised.  (Round 2 ran it with EKP_CONN_PLANES=0 / 1 on a build that could also stage the limb's PAF planes in
shared memory: profiles/r2_crowd_planes_vs_gathers.txt; that variant lost and is gone.)
usage: python tools/time_crowd.py"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    import torch_ekpose_b200 as ek
    from torch_ekpose_b200 import synthetic
    people = int(sys.argv[1])
    heat, paf = synthetic.make_batch(16, 92, 164, (people, people), seed=70 + people)
    hd, pd = torch.from_numpy(heat).cuda(), torch.from_numpy(paf).cuda()
    pp = ek.PostProcessor(device=0, max_batch=16, max_h=92, max_w=164, max_peaks=4096, max_humans=256, max_part=256, max_cand=4096)
    for _ in range(3): pp.run(hd, pd, frontend="reference")
    res = pp.results(); pp.set_timing(True)
    for _ in range(10): pp.run(hd, pd, frontend="reference")
    pp.results(); st, _ = pp.stage_times()
    print(f"{people:3d} people/image ({res['n_peaks'].mean():6.0f} peaks, {res['num_humans'].mean():5.1f} found): "
          f"front-end {st['frontend']*1e3:6.1f} sort {st['peak_sort']*1e3:5.1f} connect {st['connect']*1e3:7.1f} assemble {st['assemble']*1e3:6.1f} us")
else:
    for people in (20, 35, 50, 80):
        subprocess.run([sys.executable, __file__, str(people)])
