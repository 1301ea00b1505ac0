"""profiles/kernels.json from the ncu exports of tools/run_cfg.py captures (gpurun_out/<prefix>_<cfg>_<mode>[_reference]_raw.csv):
per kernel and configuration the per-launch figures bench.py quotes (DRAM traffic, warp instructions, ncu's own time and
issue utilisation), each with the hash of the kernel's source file AS IT WAS ON THE GPU BOX DURING THE CAPTURE (<prefix>_source_hashes.json,
tools/hash_sources.py) so that bench.py can refuse a stale entry.
usage: python tools/make_kernels_json.py <prefix, e.g. gpurun_out/r2_final> <capture label>"""
import csv, glob, hashlib, json, os, re, sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prefix, label = sys.argv[1], sys.argv[2]
SRC = {"dense_plane_kernel": "dense_frontend.cu", "dense_frontend_kernel": "dense_frontend.cu", "ref_scan_kernel": "ref_frontend.cu",
       "ref_refine_kernel": "ref_frontend.cu", "paf_connect_kernel": "paf_connect.cu", "assemble_kernel": "assemble.cu",
       "peaks_sort_kernel": "peaks_sort.cu"}
SHAPE = {"c2": (64, 46, 54), "c3": (256, 46, 82), "c4": (16, 92, 164)}


HASHES = json.load(open(prefix + "_source_hashes.json"))   # written on the GPU box by tools/hash_sources.py during the capture


def sha16(name):
    return HASHES[name]


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


out = {}
for f in sorted(glob.glob(prefix + "_*_raw.csv")):
    tag = os.path.basename(f)[len(os.path.basename(prefix)) + 1:-len("_raw.csv")]   # c2_lean, c4_lean_reference, ...
    cfg, mode = tag.split("_")[0], tag.split("_")[1]
    frontend = "reference" if tag.endswith("_reference") else "dense"
    n, h, w = SHAPE[cfg]
    rows = list(csv.reader(open(f)))
    hdr, units = rows[0], rows[1]
    col = {name: i for i, name in enumerate(hdr)}
    seen = set()
    for r in rows[2:]:
        kname = re.sub(r"^void ", "", r[col["Kernel Name"]])
        base = re.match(r"(?:ekp::)?(\w+)", kname).group(1)
        if base not in SRC or base in seen:
            continue
        seen.add(base)

        def g(metric, scale_unit=None):
            v = num(r[col[metric]]) if metric in col else None
            if v is None:
                return None
            u = units[col[metric]]
            mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "us": 1e-3, "ns": 1e-6, "ms": 1.0, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6}.get(u, 1)
            return v * mult
        variant = "mat" if (base == "dense_frontend_kernel" and mode == "mat") else ""
        key = f"{base}{'<mat>' if variant else ''}|{8 * h}x{8 * w}x{n}|{frontend}_{mode}"
        out[key] = {
            "kernel": kname[:100], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
            "source": SRC[base], "source_sha16": sha16(SRC[base]), "capture": f"{label} ({os.path.basename(f)})",
            "gpu_time_ms_under_ncu": g("gpu__time_duration.sum"),
            "warp_instructions_per_launch": g("smsp__inst_executed.sum"),
            "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "dram_bytes_per_launch": (g("dram__bytes_read.sum") or 0) + (g("dram__bytes_write.sum") or 0),
            "dram_read_bytes": g("dram__bytes_read.sum"), "dram_write_bytes": g("dram__bytes_write.sum"),
            "registers_per_thread": g("launch__registers_per_thread"),
            "l1_global_load_requests": g("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"),
            "l1_global_load_sectors": g("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"),
            "shared_bank_conflicts": g("l1tex__data_bank_conflicts_pipe_lsu.sum"),
        }
json.dump(out, open(os.path.join(root, "profiles", "kernels.json"), "w"), indent=1, sort_keys=True)
print(f"profiles/kernels.json: {len(out)} entries")
for k, v in out.items():
    print(f"  {k:70s} {v['gpu_time_ms_under_ncu'] * 1e3:8.1f} us  {v['warp_instructions_per_launch'] / 1e6:8.2f} M warp-instr  issue {v['issue_active_pct']:5.1f} %  "
          f"DRAM {v['dram_bytes_per_launch'] / 1e6:9.1f} MB")
