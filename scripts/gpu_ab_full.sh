#!/bin/bash
# attention + model parity tests, then same-box A/B of the attention kernel timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q --timeout 300 -p no:cacheprovider -k "attention or forward or full_size or c5 or dropout or decode or infill" > gpurun_out/t_attn.log 2>&1
echo "tests rc=$?"; tail -n 6 gpurun_out/t_attn.log
bash scripts/gpu_ab.sh attn_
