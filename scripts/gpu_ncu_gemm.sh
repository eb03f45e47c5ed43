#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/prof_kernels.py gemm > gpurun_out/prof_gemm.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 1 -c 14 -o gpurun_out/prof_gemm_r1 python scripts/prof_kernels.py gemm > gpurun_out/ncu_gemm.log 2>&1
echo "ncu rc=$?"; tail -n 2 gpurun_out/ncu_gemm.log; ls -la gpurun_out/prof_gemm_r1.ncu-rep
