#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 60 -p no:cacheprovider -k "cta_pair" > gpurun_out/t_pair.log 2>&1
echo "pair tests rc=$?"; tail -n 25 gpurun_out/t_pair.log
timeout 240 python -m pytest tests/test_kernels_gpu.py -q --timeout 60 -p no:cacheprovider -k "gemm" > gpurun_out/t_gemm.log 2>&1
echo "gemm tests rc=$?"; tail -n 5 gpurun_out/t_gemm.log
timeout 200 python scripts/prof_kernels.py gemm > gpurun_out/prof_gemm.log 2>&1; echo "prof rc=$?"; cat gpurun_out/prof_gemm.log | tail -9
SMER_GEMM_2SM=0 timeout 200 python scripts/prof_kernels.py gemm > gpurun_out/prof_gemm_1sm.log 2>&1; echo "1sm:"; cat gpurun_out/prof_gemm_1sm.log | tail -8
