"""profiles/<prefix>_sass_summary.txt + profiles/<prefix>_dense_frontend_mat.sass from the built library (cuobjdump -sass; no GPU needed).
usage: python tools/sass_listing.py r2"""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prefix = sys.argv[1] if len(sys.argv) > 1 else "r2"
txt = subprocess.run(["cuobjdump", "-sass", os.path.join(root, "torch_ekpose_b200", "libekpose_b200.so")], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
KEYS = ("UBLKCP", "UBLKPF", "UTMACMDFLUSH", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "FENCE.VIEW.ASYNC", "STG.E.128", "LDG.E.128", "LDS", "STS",
        "ATOMS", "ATOMG", "REDG", "BAR.SYNC", "BAR.ARV", "SHFL", "VOTE", "FFMA", "DADD", "DMUL", "HMMA", "UTCHMMA", "UTCQMMA")
rows, keep = [], None
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    ops = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, re.M)
    cnt = collections.Counter()
    for op in ops:
        for k in KEYS:
            if op.startswith(k):
                cnt[k] += 1
    demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    rows.append((demangled, len(ops), cnt))
    if "dense_frontend_kernelILb1ELb0ELi2E" in name:
        keep = f
with open(os.path.join(root, "profiles", f"{prefix}_sass_summary.txt"), "w") as o:
    o.write("SASS of torch_ekpose_b200/libekpose_b200.so (cuobjdump -sass, sm_100a only), per kernel: instruction count and the mnemonics that show\n"
            "which engines a kernel uses.  UBLKCP = cp.async.bulk on the TMA engine (.S.G global->shared load, .G.S shared->global store), UBLKPF =\n"
            "cp.async.bulk.prefetch.L2, SYNCS = mbarrier, LDGSTS = cp.async (4-byte), UTMALDG / UTMASTG = tensor-map TMA (unused: these tensors' rows\n"
            "and pixels are not 16-byte strided, dense_frontend.cu), HMMA / UTC*MMA = tensor cores (none: the path has no contraction).\n\n")
    for name, n, c in sorted(rows, key=lambda t: -t[1]):
        o.write(f"{name[:118]:120s} {n:6d} instr  " + ", ".join(f"{k} {v}" for k, v in sorted(c.items())) + "\n")
if keep:
    body = re.sub(r"\s*/\* 0x[0-9a-f]{16} \*/", "", keep)          # drop the encodings, keep address + instruction
    body = "\n".join(l for l in body.splitlines() if l.strip())
    with open(os.path.join(root, "profiles", f"{prefix}_dense_frontend_mat.sass"), "w") as o:
        o.write("Function : " + body + "\n")
print(open(os.path.join(root, "profiles", f"{prefix}_sass_summary.txt")).read())
