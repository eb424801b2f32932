#!/bin/bash
# development aid: libgptb200_timing.so = the library with per-phase cycle accounting in the batched kernel
set -e
cd "$(dirname "$0")/../gptools_b200/csrc"
for f in api batched4; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DGPT_PHASE_TIMING -I ../../include -c $f.cu -o /tmp/${f}_t.o 2>&1 | grep -E "error" || true
done
nvcc -shared -o libgptb200_timing.so /tmp/api_t.o /tmp/batched4_t.o assemble.o gemm.o factor.o predict.o -cudart static 2>&1 | grep -v warning || true
