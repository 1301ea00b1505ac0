python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_l.log 2>&1
tail -3 gpurun_out/r2_pytest_l.log
python tools/time_configs.py > gpurun_out/r2_time_configs_l.log 2>&1
EKP_FUSE_ASSEMBLE=0 python tools/time_configs.py > gpurun_out/r2_time_configs_l_unfused.log 2>&1
python tools/compat_latency.py > gpurun_out/r2_compat_latency_l.log 2>&1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_l.json 2> gpurun_out/r2_bench_l.err
