python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_e.log 2>&1
tail -3 gpurun_out/r2_pytest_e.log
python tools/time_configs.py > gpurun_out/r2_time_configs_e.log 2>&1
EKPOSE_B200_SO=build/variants/connprof.so python tools/conn_profile.py > gpurun_out/r2_conn_profile_e.log 2>&1
