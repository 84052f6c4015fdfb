#!/bin/sh
# Builds an instrumented / experimental variant of libmcz_b200.so next to the product library:
#   tools/build_variant.sh <tag> [-DMACRO ...]   ->  dbgbuild/libmcz_<tag>.so
# Use it with MCZ_B200_LIB=dbgbuild/libmcz_<tag>.so (pymcz_b200/_lib.py).
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
mkdir -p dbgbuild
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
  -shared -o dbgbuild/libmcz_$tag.so pymcz_b200/csrc/mcz_abi.cu pymcz_b200/csrc/mcz_replay.cu \
  pymcz_b200/csrc/mcz_percentiles.cu pymcz_b200/csrc/mcz_production.cu pymcz_b200/csrc/mcz_histassist.cu -lcudart
echo built dbgbuild/libmcz_$tag.so
