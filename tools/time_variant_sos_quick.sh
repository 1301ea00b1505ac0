#!/bin/bash
# On the GPU box: per-stage times (tools/time_configs.py, tools/time_variants.py) with every library variant.
cd "$(dirname "$0")/.."
for so in build/variants/*.so; do
  echo "== $(basename $so .so)"
  EKPOSE_B200_SO=$PWD/$so python tools/time_configs.py 2>&1
  EKPOSE_B200_SO=$PWD/$so python tools/time_variants.py 2>&1 | grep -v "thr="
done
