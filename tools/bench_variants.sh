#!/bin/bash
# development aid: time the batched kernel for several library variants (tools/build_variant.sh) in one GPU session and
# compare their results with the default library.  usage: tools/bench_variants.sh <B> <name> [<name> ...]
B=$1; shift
python tools/bench_quick.py $B /tmp/v_base.npy | sed 's/^/base: /'
for n in "$@"; do
  GPTB200_LIB=gptools_b200/csrc/libgptb200_$n.so python tools/bench_quick.py $B /tmp/v_$n.npy | sed "s/^/$n: /"
  python -c "import numpy as np; a=np.load('/tmp/v_base.npy'); b=np.load('/tmp/v_$n.npy'); print('$n: max rel diff vs base', float(np.max(np.abs(a-b)/np.maximum(1.0,np.abs(a)))))"
done
