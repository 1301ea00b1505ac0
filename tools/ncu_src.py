"""Per-source-line summary of a kernel in an .ncu-rep (needs `--import-source on` at capture time and -lineinfo):
reads ncu's own CUDA/SASS correlation (`--page source --print-source cuda,sass`), so template instantiations
and inlined device functions are attributed correctly.
usage: python tools/ncu_src.py report.ncu-rep|export.csv.gz kernel_regex [top_n] [launch_index]"""
import collections
import csv
import os
import subprocess
import sys

rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
if rep.endswith(".csv.gz") or rep.endswith(".csv"):   # an export made on the GPU box: ncu -i X --page source --print-source cuda,sass --csv
    import gzip
    import re
    out = (gzip.open(rep, "rt") if rep.endswith(".gz") else open(rep)).read()
else:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "-k", f"regex:{kname}"],
                         capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# the output is a sequence of (File Path, Function Name, header, rows...) blocks: one per source file and launch
launches = collections.OrderedDict()
cur_file, cur_fn, hdr = None, None, None
seen = collections.Counter()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = os.path.basename(r[1]); continue
    if r[0] == "Function Name":
        cur_fn = r[1]
        hdr = None
        continue
    if r[0] == "Line No":
        hdr = r
        seen[(cur_fn, cur_file)] += 1
        key = (cur_fn, seen[(cur_fn, cur_file)] - 1)
        launches.setdefault(key, [])
        continue
    if hdr is None or not r[0].strip().isdigit():
        continue   # SASS rows and "..." rows
    launches[(cur_fn, seen[(cur_fn, cur_file)] - 1)].append((cur_file, hdr, r))
fns = sorted({k[0] for k in launches if __import__("re").search(kname, k[0])})
for fn in fns:
    items = launches.get((fn, which)) or []
    if not items:
        continue
    agg = []
    for f, h, r in items:
        def g(name):   # columns are addressed from the right: an unescaped quote in the source text may split column 1
            if name not in h:
                return 0
            v = r[h.index(name) - len(h)]
            try:
                return int(float(v))
            except ValueError:
                return 0
        stalls = {n: g(n) for n in h if n.startswith("stall_") and "Not Issued" not in n}
        agg.append((f, int(r[0]), r[1].strip()[:90], g("Warp Stall Sampling (All Samples)"), g("Instructions Executed"),
                    g("L1 Tag Requests Global"), g("L1 Wavefronts Shared"), g("L1 Wavefronts Shared Excessive"), stalls))
    ts = sum(a[3] for a in agg) or 1
    te = sum(a[4] for a in agg) or 1
    tot_stall = collections.Counter()
    for a in agg:
        tot_stall.update(a[8])
    print(f"== {fn[:110]}\n   warp instructions {te}, stall samples {ts}, L1 global tag requests {sum(a[5] for a in agg)}, "
          f"shared wavefronts {sum(a[6] for a in agg)} (excess {sum(a[7] for a in agg)})")
    print("   stall reasons: " + ", ".join(f"{k[6:]} {100 * v / max(sum(tot_stall.values()), 1):.0f}%" for k, v in tot_stall.most_common(7)))
    for a in sorted(agg, key=lambda a: -a[3])[:top]:
        main = max(a[8].items(), key=lambda kv: kv[1])[0][6:] if a[3] else ""
        print(f"   {a[0]:20s}:{a[1]:4d} inst={a[4]:9d} ({100 * a[4] / te:4.1f}%) samp={a[3]:6d} ({100 * a[3] / ts:4.1f}%) tags={a[5]:8d} "
              f"shw={a[6]:8d} {main:10s}| {a[2]}")
