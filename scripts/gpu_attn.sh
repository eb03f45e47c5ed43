#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q --timeout 120 -p no:cacheprovider -k "attention" > gpurun_out/t_attn.log 2>&1
echo "attn rc=$?"; tail -n 40 gpurun_out/t_attn.log
timeout 600 python -m pytest tests/test_model_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/t_model.log 2>&1
echo "model rc=$?"; tail -n 30 gpurun_out/t_model.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train2.log 2>&1
echo "bench rc=$?"; tail -c 3500 gpurun_out/bench_train2.log
