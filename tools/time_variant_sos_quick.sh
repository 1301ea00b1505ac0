for so in build/variants/*.so; do
  name=$(basename $so .so)
  echo "== $name"; EKPOSE_B200_SO=$PWD/$so python tools/time_variants.py 2>&1 | head -2
done
