python -m pytest tests -m gpu -x -q -k "gauss or nms or reference_frontend" > gpurun_out/r2c_pytest_gauss.log 2>&1
tail -3 gpurun_out/r2c_pytest_gauss.log
t0=$(date +%s)
python bench.py > gpurun_out/r2c_bench_ctx2.json 2> gpurun_out/r2c_bench_ctx2.err
t1=$(date +%s); echo "bench default took $((t1-t0)) s"
EKP_BENCH_CFG_CONTEXTS=4 python bench.py > gpurun_out/r2c_bench_ctx4.json 2> gpurun_out/r2c_bench_ctx4.err
t2=$(date +%s); echo "bench ctx4 took $((t2-t1)) s"
python - <<'PY'
import json
for f in ("gpurun_out/r2c_bench_ctx2.json","gpurun_out/r2c_bench_ctx4.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value %.0f e2e %.0f" % (d["value"], d["e2e"]["value"]))
        for k,v in d["configs"].items():
            if "images_per_s_per_gpu" in v: print("   %-45s %10.0f img/s  %.1f us/batch" % (k, v["images_per_s_per_gpu"], v["ms_per_batch_pipelined"]*1e3))
    except Exception as e: print(f, "ERR", e)
PY
