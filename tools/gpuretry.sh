#!/bin/bash
# usage: [GPUS=N] gpuretry.sh <outfile> <timeout> <cmd>
out=$1; to=$2; shift 2
extra=""
if [ -n "$GPUS" ]; then extra="--gpus $GPUS"; fi
for i in 1 2 3 4 5 6 7 8 9 10 11 12 13 14 15; do
  /usr/local/graft/bin/gpurun $extra --timeout $to -- "$@" > $out 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $out; then exit $rc; fi
  sleep 120
done
exit 3
