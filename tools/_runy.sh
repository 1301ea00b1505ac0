python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_o.log 2>&1
tail -3 gpurun_out/r2_pytest_o.log
python tools/time_configs.py > gpurun_out/r2_time_configs_o.log 2>&1
EKP_CONN_ORDER=0 python tools/time_configs.py > gpurun_out/r2_time_configs_o_noorder.log 2>&1
