python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_g.log 2>&1
tail -3 gpurun_out/r2_pytest_g.log
python tools/time_configs.py > gpurun_out/r2_time_configs_g.log 2>&1
EKP_LEAN_TILED=1 python tools/time_configs.py > gpurun_out/r2_time_configs_g_tiled.log 2>&1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.err
