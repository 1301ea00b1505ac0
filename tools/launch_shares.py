"""Per-kernel launch counts, mean duration and share of the summed kernel time from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file <csv> <command>`).
usage: python tools/launch_shares.py <launches.csv> ["header text"]"""
import csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
acc = {}
for r in rows:
    name = re.sub(r"^void ", "", r[ki]).split("(")[0].replace("ekp::", "")
    t = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    a = acc.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(a[1] for a in acc.values())
if len(sys.argv) > 2:
    print(sys.argv[2] + "\n")
for name, (n, t) in acc.items():
    print(f"{name:60s} launches={n:4d} mean={t / n:9.1f} us share={t / tot:.3f}")
