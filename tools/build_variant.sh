#!/bin/bash
# development: build a libimgcorr variant with extra -D flags into gpurun_out-independent path variants/<name>.so
# usage: tools/build_variant.sh name -DK2S_MINB_V=5 ...
set -e
name=$1; shift
mkdir -p variants build/var_$name
cd "$(dirname "$0")/.."
for f in k1_pointwise_median k1_stream k1_stream5 k2_undistort k3_warp k4_ste selftest k5_producers imgcorr_api; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden --expt-relaxed-constexpr "$@" -c imgprocessor_b200/csrc/$f.cu -o build/var_$name/$f.o 2>/dev/null &
done
wait
nvcc -shared -o variants/$name.so build/var_$name/*.o -Xcompiler -fvisibility=hidden 2>/dev/null
echo variants/$name.so
