#!/bin/bash
# builds the csrc/ of git revision $1 (default HEAD) into _ab/libsmer_b200_A.so for same-box A/B timing:
#   SMER_B200_LIB=_ab/libsmer_b200_A.so python scripts/prof_kernels.py attn_
set -e
REV=${1:-HEAD}
rm -rf _ab/src; mkdir -p _ab/src/csrc _ab/include
for f in $(git ls-tree --name-only $REV smer_music_generation_b200/csrc/); do git show $REV:$f > _ab/src/csrc/$(basename $f); done
git show $REV:include/smer_b200.h > _ab/include/smer_b200.h
mkdir -p _ab/src/inc2/include; cp _ab/include/smer_b200.h _ab/src/inc2/include/
objs=""
for f in _ab/src/csrc/*.cu; do
  o=_ab/src/$(basename $f .cu).o
  # sources include "../../include/smer_b200.h" relative to csrc/: mirror that layout
  ( /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -c $f -o $o ) &
  objs="$objs $o"
done
wait
/usr/local/cuda/bin/nvcc -shared -o _ab/libsmer_b200_A.so $objs -lcudart
ls -la _ab/libsmer_b200_A.so
