python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_f.log 2>&1
tail -3 gpurun_out/r2_pytest_f.log
python tools/time_configs.py > gpurun_out/r2_time_configs_f.log 2>&1
EKP_CONN_BIG_THREADS=1024 python tools/time_configs.py > gpurun_out/r2_time_configs_f_1024.log 2>&1
EKP_CONN_BIG_THREADS=256 python tools/time_configs.py > gpurun_out/r2_time_configs_f_256.log 2>&1
