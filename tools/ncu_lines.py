"""Per-source-line breakdown of an ncu capture: joins the SASS page of `ncu --page source --csv` with the
line table of the kernel's cubin (nvdisasm -g), instruction by instruction.
usage: python tools/ncu_lines.py report.ncu-rep kernel_substring [top_n]"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "torch_ekpose_b200", "libekpose_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, capture_output=True)
lines_of = None
for f in sorted(os.listdir(tmp)):
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    if kname not in dis:
        continue
    cur, fn, ln, per_fn = None, "?", 0, collections.defaultdict(list)
    for l in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", l)
        if m:
            cur = m.group(1); continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            fn, ln = os.path.basename(m.group(1)), int(m.group(2)); continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            per_fn[cur].append((fn, ln))
    for k, v in per_fn.items():
        if kname in k:
            lines_of = v
    if lines_of:
        break
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur); continue
    if cur is not None:
        cur["rows"].append(r)
blk = next(b for b in blocks if kname in b["name"])
hdr = blk["rows"][0]
iss, ie = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = [(int(r[iss] or 0), int(r[ie] or 0)) for r in blk["rows"][1:] if len(r) > ie]
if len(data) != len(lines_of):
    print(f"warning: {len(data)} SASS rows vs {len(lines_of)} disassembled instructions", file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0])
for (s, e), (fn, ln) in zip(data, lines_of):
    agg[(fn, ln)][0] += s; agg[(fn, ln)][1] += e
ts, te = sum(v[0] for v in agg.values()) or 1, sum(v[1] for v in agg.values()) or 1
print(f"{blk['name'][:60]}: total warp instr {te} samples {ts}")
src_cache = {}
def src(fn, ln):
    for d in ("torch_ekpose_b200/csrc", "include"):
        p = os.path.join(root, d, fn)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            return src_cache[p][ln - 1].strip()[:100] if ln - 1 < len(src_cache[p]) else ""
    return ""
for (fn, ln), (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{fn:20s}:{ln:4d} inst={e:10d} ({100*e/te:4.1f}%) samp={s:6d} ({100*s/ts:4.1f}%) | {src(fn, ln)}")
