#!/bin/bash
mkdir -p gpurun_out
for m in 1 0 2; do
  echo "== SMER_GEMM_2SM=$m"
  SMER_GEMM_2SM=$m timeout 300 python scripts/prof_kernels.py gemm 2>&1 | tee gpurun_out/prof_gemm_2sm$m.log
done
