set -x
python tools/time_configs.py > gpurun_out/r2_time_configs_base.log 2>&1
python tools/compat_latency.py > gpurun_out/r2_compat_latency_base.log 2>&1
python tools/run_c4.py lean > gpurun_out/r2_plain_c4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"paf_connect|assemble|dense_frontend|peaks_sort" -s 8 -c 4 -o gpurun_out/r2_base_c4_lean python tools/run_c4.py lean > gpurun_out/r2_ncu_c4.log 2>&1
python tools/run_c2.py lean > gpurun_out/r2_plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"paf_connect|assemble|dense_frontend|peaks_sort" -s 8 -c 4 -o gpurun_out/r2_base_c2_lean python tools/run_c2.py lean > gpurun_out/r2_ncu_c2.log 2>&1
nvidia-smi topo -m > gpurun_out/r2_topo.log 2>&1; nproc >> gpurun_out/r2_topo.log; lscpu | head -30 >> gpurun_out/r2_topo.log
