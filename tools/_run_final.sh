python tools/hash_sources.py > gpurun_out/r2_final_source_hashes.json
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_final.log 2>&1
tail -3 gpurun_out/r2_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke_final.log 2>&1
for c in "c2 lean" "c4 lean" "c2 mat" "c3 lean" "c4 lean reference" "c2 lean reference"; do
  tag=$(echo $c | tr ' ' '_')
  python tools/run_cfg.py $c > gpurun_out/r2_plain_$tag.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"paf_connect|assemble|dense_|peaks_sort|ref_" -s 16 -c 5 -o /tmp/r2_final_$tag python tools/run_cfg.py $c > gpurun_out/r2_ncu_$tag.log 2>&1
  ncu -i /tmp/r2_final_$tag.ncu-rep --page raw --csv > gpurun_out/r2_final_${tag}_raw.csv 2>/dev/null
  ncu -i /tmp/r2_final_$tag.ncu-rep --page source --print-source cuda,sass --csv 2>/dev/null | gzip > gpurun_out/r2_final_${tag}_source.csv.gz
done
EKP_BENCH_BATCHES_PER_STEP=4 python bench.py --steps 2 --warmup 3 --headline-only --no-cpu-baseline > gpurun_out/r2_plain_bench_small.log 2>&1 && \
EKP_BENCH_BATCHES_PER_STEP=4 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 3 --headline-only --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
python tools/time_configs.py > gpurun_out/r2_time_configs_final.log 2>&1
python tools/compat_latency.py > gpurun_out/r2_compat_latency_final.log 2>&1
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_ref_final.json 2> gpurun_out/r2_bench_ref_final.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
du -sh gpurun_out
