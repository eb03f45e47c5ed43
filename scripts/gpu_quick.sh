#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_model_gpu.py -q --timeout 120 -p no:cacheprovider > gpurun_out/t_gemm.log 2>&1
echo "gemm tests rc=$?"; tail -n 15 gpurun_out/t_gemm.log
timeout 300 python scripts/prof_kernels.py > gpurun_out/prof_plain.log 2>&1
echo "prof plain rc=$?"; cat gpurun_out/prof_plain.log | tail -20
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train4.log 2>&1
echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_train4.log
