python tools/hash_sources.py > gpurun_out/r2_final_source_hashes.json
for c in "c2 lean" "c4 lean" "c2 mat" "c3 lean"; do
  tag=$(echo $c | tr ' ' '_')
  python tools/run_cfg.py $c > gpurun_out/r2_plain_$tag.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"paf_connect|assemble|dense_|peaks_sort|ref_" -s 12 -c 4 -o /tmp/r2_final_$tag python tools/run_cfg.py $c > gpurun_out/r2_ncu_$tag.log 2>&1
  ncu -i /tmp/r2_final_$tag.ncu-rep --page raw --csv > gpurun_out/r2_final_${tag}_raw.csv 2>/dev/null
  ncu -i /tmp/r2_final_$tag.ncu-rep --page source --print-source cuda,sass --csv 2>/dev/null | gzip > gpurun_out/r2_final_${tag}_source.csv.gz
done
du -sh gpurun_out
