python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_i.log 2>&1
tail -3 gpurun_out/r2_pytest_i.log
python tools/compat_latency.py > gpurun_out/r2_compat_latency_i.log 2>&1
python tools/time_configs.py > gpurun_out/r2_time_configs_i.log 2>&1
