#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active,utilization.gpu --format=csv -lms 250 > gpurun_out/smi_decode.csv &
SMI=$!
timeout 600 python scripts/gpu_decode_diag2.py 1024 2>&1 | tail -12
kill $SMI
awk -F, 'NR>1{print $2,$3,$4,$6,$7}' gpurun_out/smi_decode.csv | sort | uniq -c | sort -rn | head -20
nproc; uptime
