#!/bin/bash
# same-box A/B: _ab/libsmer_b200_A.so (scripts/build_ab.sh <rev>) vs the in-tree build; $1 = prof_kernels filter
mkdir -p gpurun_out
for i in 1 2; do
  echo "== A (old)"; SMER_B200_LIB=$PWD/_ab/libsmer_b200_A.so timeout 300 python scripts/prof_kernels.py $1 2>&1 | tail -n 16
  echo "== B (new)"; timeout 300 python scripts/prof_kernels.py $1 2>&1 | tail -n 16
done
