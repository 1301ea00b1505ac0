EKPOSE_B200_SO=build/variants/asmprof.so python tools/asm_profile.py > gpurun_out/r2_asm_profile_a.log 2>&1
EKPOSE_B200_SO=build/variants/connprof.so python tools/conn_profile.py > gpurun_out/r2_conn_profile_a.log 2>&1
