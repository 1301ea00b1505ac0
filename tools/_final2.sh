set -x
python tools/hash_sources.py > gpurun_out/r2g_source_hashes.json
python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; tail -2 gpurun_out/r2g_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; tail -1 gpurun_out/r2g_smoke.log
for c in "c2 lean" "c4 lean" "c2 mat" "c3 lean"; do
  tag=$(echo $c | tr ' ' '_')
  ncu --set full --clock-control none --import-source on -k regex:"paf_connect|assemble|dense_|peaks_sort|ref_" -s 12 -c 4 -o /tmp/r2g_$tag python tools/run_cfg.py $c > gpurun_out/r2g_ncu_$tag.log 2>&1
  ncu -i /tmp/r2g_$tag.ncu-rep --page raw --csv > gpurun_out/r2g_${tag}_raw.csv 2>/dev/null
  ncu -i /tmp/r2g_$tag.ncu-rep --page source --print-source cuda,sass --csv 2>/dev/null | gzip > gpurun_out/r2g_${tag}_source.csv.gz
done
for c in "c4 lean reference" "c2 lean reference"; do
  tag=$(echo $c | tr ' ' '_')
  ncu --set full --clock-control none --import-source on -k regex:"paf_connect|assemble|dense_|peaks_sort|ref_" -s 15 -c 5 -o /tmp/r2g_$tag python tools/run_cfg.py $c > gpurun_out/r2g_ncu_$tag.log 2>&1
  ncu -i /tmp/r2g_$tag.ncu-rep --page raw --csv > gpurun_out/r2g_${tag}_raw.csv 2>/dev/null
  ncu -i /tmp/r2g_$tag.ncu-rep --page source --print-source cuda,sass --csv 2>/dev/null | gzip > gpurun_out/r2g_${tag}_source.csv.gz
done
python tools/make_kernels_json.py gpurun_out/r2g "round 2 final kernels" > gpurun_out/r2g_kernels_json.log 2>&1
python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; tail -c 200 gpurun_out/r2g_bench.err
python tools/time_configs.py > gpurun_out/r2g_time_configs.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2g_launches.csv env EKP_BENCH_BATCHES_PER_STEP=4 python bench.py --steps 2 --warmup 3 --headline-only --no-cpu-baseline > gpurun_out/r2g_ncu_launches.log 2>&1
du -sh gpurun_out
