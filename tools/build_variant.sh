#!/bin/bash
# builds crender_b200/_variants/libv_<name>.so with extra nvcc defines (A/B experiments on the GPU box):
#   tools/build_variant.sh pf1 -DCRB_PREFETCH=1
set -e
cd "$(dirname "$0")/../crender_b200/csrc"
name=$1; shift
out=../_variants; mkdir -p $out/_b_$name
for f in bvh_build scene trace render post multi capi; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off "$@" -c $f.cu -o $out/_b_$name/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libv_$name.so $out/_b_$name/*.o -lcudart -ldl
rm -rf $out/_b_$name
echo built $out/libv_$name.so
