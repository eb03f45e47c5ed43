#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q --timeout 120 -p no:cacheprovider -k "gemm or layernorm or embed" > gpurun_out/t_gemm.log 2>&1
echo "gemm tests rc=$?"; tail -n 15 gpurun_out/t_gemm.log
timeout 300 python scripts/prof_kernels.py > gpurun_out/prof_plain.log 2>&1
rc=$?; echo "prof plain rc=$rc"; cat gpurun_out/prof_plain.log | tail -20
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train3.log 2>&1
echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_train3.log
if [ $rc -eq 0 ]; then
  timeout 300 python scripts/prof_kernels.py > /dev/null 2>&1 &&
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|attn_fwd_tc|attn_bwd_d" -c 40 -o gpurun_out/prof_r1b python scripts/prof_kernels.py > gpurun_out/ncu_full.log 2>&1
  echo "ncu rc=$?"; tail -n 3 gpurun_out/ncu_full.log
fi
