python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_k.log 2>&1
tail -3 gpurun_out/r2_pytest_k.log
EKP_TRACE_PROCESS_PAF=1 python tools/compat_latency.py > gpurun_out/r2_compat_latency_k.log 2> gpurun_out/r2_compat_latency_k.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_k.log 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_k.json 2> gpurun_out/r2_bench_k.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref_k.json 2> gpurun_out/r2_bench_ref_k.err
