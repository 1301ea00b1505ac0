#!/bin/bash
# Rebuild the library with different (row batch, min blocks/SM) and time the bench; run on the GPU box.
for cfg in "1 4" "2 4" "4 3" "4 4" "8 2" "8 3"; do
  set -- $cfg
  make -C torch_ekpose_b200/csrc -B EXTRA="-DEKP_ROW_BATCH=$1 -DEKP_MIN_BLOCKS=$2" > /dev/null 2>&1
  python bench.py --no-cpu-baseline --steps 200 > gpurun_out/sweep_$1_$2.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/sweep_$1_$2.json')); print('batch $1 minblocks $2:', round(d['value']), 'img/s  kernel', round(d['roofline']['kernel_ms_isolated'],4), 'ms  frac', round(d['roofline']['frac'],3))"
done
