#!/bin/bash
# Rebuild the library with different tile / CTA shapes and time the bench; run on the GPU box.
# columns: row batch, min blocks per SM, threads per CTA, max tile width (stride-8 columns)
for cfg in "4 3 256 32" "4 2 384 64" "4 1 768 64" "4 2 512 64" "4 3 256 64" "4 4 192 64"; do
  set -- $cfg
  make -C torch_ekpose_b200/csrc -B EXTRA="-DEKP_ROW_BATCH=$1 -DEKP_MIN_BLOCKS=$2 -DEKP_THREADS=$3 -DEKP_MAX_TWL=$4" > /dev/null 2>&1
  python bench.py --no-cpu-baseline --steps 200 > gpurun_out/sweep_$1_$2_$3_$4.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/sweep_$1_$2_$3_$4.json')); print('batch $1 minblocks $2 threads $3 maxtwl $4:', round(d['value']), 'img/s  kernel', round(d['roofline']['kernel_ms_isolated'],4), 'ms  frac', round(d['roofline']['frac'],3))"
done
