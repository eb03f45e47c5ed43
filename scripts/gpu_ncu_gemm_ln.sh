#!/bin/bash
# plain run, then ncu --set full of the GEMM / LayerNorm / colsum kernels at the C2 shapes (one launch each after warm-up)
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 300 python scripts/prof_kernels.py > gpurun_out/prof_plain.log 2>&1 || { echo plain failed; exit 1; }
timeout 1200 ncu --set full --clock-control none -k regex:"gemm_tc_kernel|ln_fwd_bf16|ln_bwd_bf16" -c 22 -o gpurun_out/prof_gemm_ln_r1r python scripts/prof_kernels.py > gpurun_out/ncu_gemm_ln.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/*.ncu-rep
