EKP_TRACE_PROCESS_PAF=1 python tools/compat_latency.py > gpurun_out/r2_compat_latency_j.log 2> gpurun_out/r2_compat_latency_j.err
