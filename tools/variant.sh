#!/bin/bash
# Builds an experimental variant of the library: tools/variant.sh NAME -DFOO -DBAR=1  ->  tools/_variants/libdesmo_NAME.so
# (run it with DESMO_B200_LIB=tools/_variants/libdesmo_NAME.so).  Needs an up-to-date desmo_b200/build/*.o.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p tools/_variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
  -c desmo_b200/csrc/fused_tc.cu -o tools/_variants/fused_tc_$name.o
objs=$(ls desmo_b200/build/*.o | grep -v fused_tc.o)
nvcc -shared -o tools/_variants/libdesmo_$name.so $objs tools/_variants/fused_tc_$name.o -lcudart
echo tools/_variants/libdesmo_$name.so
