#!/bin/bash
# A/B builds of libshafa_cuda with extra compiler flags:  tools/build_variant.sh NAME "-DFOO=1 ..."
# -> shafa-cd_b200/build/libshafa_NAME.so, selected at run time with SHAFA_LIB=<path> (binding.py).
set -e
name=$1; flags=$2
here=$(cd "$(dirname "$0")/.." && pwd)
cd "$here/shafa-cd_b200"
tmp=$(mktemp -d)
for f in api multi host_pipeline hist sf_build sf_encode sf_decode sf_decode_fused sf_wide rle_encode rle_decode; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function --expt-relaxed-constexpr $flags -c csrc/$f.cu -o $tmp/$f.o >/dev/null 2>&1 &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/libshafa_$name.so $tmp/*.o -lcudart_static -lpthread -ldl -lrt
rm -rf $tmp
echo "built shafa-cd_b200/build/libshafa_$name.so ($flags)"
