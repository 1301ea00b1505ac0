set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_a.log 2>&1
tail -5 gpurun_out/r2_pytest_a.log
python tools/time_configs.py > gpurun_out/r2_time_configs_a.log 2>&1
EKP_GRAPHS=0 python tools/time_configs.py > gpurun_out/r2_time_configs_a_nograph.log 2>&1
python tools/compat_latency.py > gpurun_out/r2_compat_latency_a.log 2>&1
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
