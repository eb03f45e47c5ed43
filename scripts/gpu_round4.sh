#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_train8.log 2>&1
echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_train8.log
timeout 600 python bench.py --workload decode --pieces 1024 --decode-len 64 --steps 1 --warmup 0 > gpurun_out/plain_dec.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"decode_attn_kernel" -s 40 -c 4 -o gpurun_out/prof_decode_r1 python bench.py --workload decode --pieces 1024 --decode-len 64 --steps 1 --warmup 0 > gpurun_out/ncu_dec.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_decode_r1.ncu-rep
