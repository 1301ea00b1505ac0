"""CPU: the benchmark's reference arm runs anywhere (no GPU) and prints the contracted JSON line;
the product arm refuses to run without a GPU instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["metric"].startswith("postprocess images/sec at 368x432") and d["scaling"] == "weak" and d["dtype"] == "f32"
    assert d["config"]["workload"].startswith("configs[1]") and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_do_no_work():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            return
    except Exception:
        pass
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                         timeout=300, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def test_committed_ncu_figures_belong_to_the_committed_kernels():
    """profiles/kernels.json (DRAM traffic of the roofline kernel, warp instructions of the issue-bound ones: what the bench
    line quotes) carries the hash of each kernel source at capture time; bench.py refuses a stale entry at run time, and
    this test keeps the committed table and the committed sources together: a kernel edit needs a new capture
    (tools/run_cfg.py under ncu, tools/make_kernels_json.py)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ekp_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    table = bench.ncu_constants()
    assert len(table) >= 20
    stale = sorted(k for k, v in table.items() if v["stale"])
    assert not stale, f"captured from other sources: {stale}"
    mat = table["dense_frontend_kernel<mat>|368x432x64|dense_mat"]
    algo = bench.algo_bytes(46, 54) * 64
    assert 0.9 * algo <= mat["dram_bytes_per_launch"] <= 1.1 * algo   # no wasted traffic, nothing uncounted
