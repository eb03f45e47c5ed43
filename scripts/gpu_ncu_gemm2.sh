#!/bin/bash
# ncu --set full of the QKV GEMM (K=512) with and without CTA pairs (plain runs happened in gpu_gemm3.sh)
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 300 python scripts/prof_kernels.py gemm_qkv > gpurun_out/plain_qkv.log 2>&1 || exit 1
for m in 1 2; do
  SMER_GEMM_2SM=$m timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel --launch-skip 2 -c 1 \
    -o gpurun_out/gemm_qkv_2sm$m python scripts/prof_kernels.py gemm_qkv > gpurun_out/ncu_gemm_$m.log 2>&1
  echo "ncu $m rc=$?"
done
ls -la gpurun_out/*.ncu-rep
