python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_m.log 2>&1
tail -3 gpurun_out/r2_pytest_m.log
python tools/time_configs.py > gpurun_out/r2_time_configs_m.log 2>&1
