#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q --timeout 120 -p no:cacheprovider -k "attention" > gpurun_out/t_attn.log 2>&1
echo "attn rc=$?"; tail -n 12 gpurun_out/t_attn.log
for i in 1 2; do
  echo "== one-buffer"; SMER_ATTN_BWD_PIPE=0 timeout 300 python scripts/prof_kernels.py attn_bwd 2>&1 | tail -n 2
  echo "== pipelined";  SMER_ATTN_BWD_PIPE=1 timeout 300 python scripts/prof_kernels.py attn_bwd 2>&1 | tail -n 2
done
