#!/bin/bash
# Build variants of the library that differ only in ONE source file's compile-time knobs (default
# dense_frontend.cu; VARIANT_SRC=paf_connect.cu selects another), HERE (no GPU needed), into
# build/variants/ (git-ignored; travels to the GPU box).  usage:
#   tools/build_variants.sh name1 "-DEKP_X=1 -DEKP_Y=2" name2 "..." ...
# then on the GPU box:  tools/time_variant_sos.sh
set -e
cd "$(dirname "$0")/.."
make -C torch_ekpose_b200/csrc > /dev/null
mkdir -p build/variants
rm -f build/variants/*.so
ARCH="-gencode arch=compute_100a,code=sm_100a"
SRC=${VARIANT_SRC:-dense_frontend.cu}
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( nvcc $ARCH -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-fno-fast-math,-ffp-contract=off $flags \
      -c torch_ekpose_b200/csrc/$SRC -o build/variants/$name.o &&
    nvcc $ARCH -shared -o build/variants/$name.so build/variants/$name.o \
      $(ls torch_ekpose_b200/csrc/*.o | grep -v ${SRC%.cu}.o) && rm build/variants/$name.o && echo "built $name: $flags" ) &
done
wait
