python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_d.log 2>&1
tail -3 gpurun_out/r2_pytest_d.log
python tools/time_configs.py > gpurun_out/r2_time_configs_d.log 2>&1
python tools/time_configs.py default_caps > gpurun_out/r2_time_configs_d_defcaps.log 2>&1
EKP_STAGE_MIN_PAIRS=0 python tools/time_configs.py > gpurun_out/r2_time_configs_d_alwaysstage.log 2>&1
EKP_BY_SAMPLE_MAX_PAIRS=0 EKP_STAGE_MIN_PAIRS=100000 python tools/time_configs.py > gpurun_out/r2_time_configs_d_old.log 2>&1
EKPOSE_B200_SO=build/variants/asmprof.so python tools/asm_profile.py > gpurun_out/r2_asm_profile_d.log 2>&1
EKPOSE_B200_SO=build/variants/connprof.so python tools/conn_profile.py > gpurun_out/r2_conn_profile_d.log 2>&1
python tools/compat_latency.py > gpurun_out/r2_compat_latency_d.log 2>&1
( time python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err ) 2> gpurun_out/r2_bench_d.time
tail -3 gpurun_out/r2_bench_d.err
