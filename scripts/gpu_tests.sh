#!/bin/bash
# Runs on the GPU box (gpurun): kernel tests first (tcgen05 GEMM isolated with a short timeout so
# a hang cannot eat the call), then the module tests, then smoke.
mkdir -p gpurun_out
nvidia-smi -L
timeout 240 python -m pytest tests/test_kernels_gpu.py -q -k "gemm_tc" --timeout 60 -p no:cacheprovider > gpurun_out/t_tc.log 2>&1
echo "tc gemm rc=$?"; tail -n 25 gpurun_out/t_tc.log
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "not gemm_tc" --timeout 120 -p no:cacheprovider > gpurun_out/t_kern.log 2>&1
echo "kernels rc=$?"; tail -n 40 gpurun_out/t_kern.log
timeout 1200 python -m pytest tests/test_model_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/t_model.log 2>&1
echo "model rc=$?"; tail -n 60 gpurun_out/t_model.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -n 10 gpurun_out/smoke.log
