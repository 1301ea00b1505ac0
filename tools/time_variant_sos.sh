#!/bin/bash
# On the GPU box: time bench.py with every library variant under build/variants/.
cd "$(dirname "$0")/.."
for so in build/variants/*.so; do
  name=$(basename $so .so)
  EKPOSE_B200_SO=$PWD/$so timeout 150 python bench.py --no-cpu-baseline --steps 200 > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err
  python -c "
import json; d=json.load(open('gpurun_out/var_$name.json')); print('%-28s' % '$name', round(d['value']), 'img/s  kernel isolated', round(d['roofline']['kernel_ms_isolated'],4), 'ms  frac', round(d['roofline']['frac'],3), ' humans', d['humans_found_last_step'])" || tail -3 gpurun_out/var_$name.err
done
