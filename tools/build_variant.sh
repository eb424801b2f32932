#!/bin/bash
# development aid: libgptb200_<name>.so = the library with batched4.cu (and api.cu) rebuilt with extra -D flags
# usage: tools/build_variant.sh <name> "<nvcc -D flags>"
set -e
name=$1; shift
flags="$*"
cd "$(dirname "$0")/../gptools_b200/csrc"
for f in api batched4; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $flags -I ../../include -c $f.cu -o /tmp/${f}_${name}.o
done
nvcc -shared -o libgptb200_${name}.so /tmp/api_${name}.o /tmp/batched4_${name}.o assemble.o gemm.o factor.o predict.o -cudart static
echo built libgptb200_${name}.so
