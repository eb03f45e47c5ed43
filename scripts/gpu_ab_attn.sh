#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q --timeout 120 -p no:cacheprovider -k "attention" > gpurun_out/t_attn.log 2>&1
echo "attn rc=$?"; tail -n 4 gpurun_out/t_attn.log
bash scripts/gpu_ab.sh attn_
