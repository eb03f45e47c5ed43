#!/bin/bash
# default bench line (with CPU baseline + module API), the reference arm, and the other config shapes on one GPU
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "default rc=$?"; tail -c 600 gpurun_out/bench_default.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "reference rc=$?"; tail -c 700 gpurun_out/bench_reference.log
timeout 600 python bench.py --tgt 256 --steps 10 --warmup 3 --no-cpu-baseline --no-module-api > gpurun_out/bench_T256.log 2>&1; echo "T256 rc=$?"; tail -c 300 gpurun_out/bench_T256.log
timeout 600 python bench.py --batch 64 --seq 2048 --tgt 2048 --steps 5 --warmup 3 --no-cpu-baseline --no-module-api > gpurun_out/bench_c3.log 2>&1; echo "c3 rc=$?"; tail -c 300 gpurun_out/bench_c3.log
timeout 600 python bench.py --batch 2 --seq 4096 --tgt 4096 --d-model 768 --nhead 12 --layers 12 --ff 3072 --steps 5 --warmup 3 --no-cpu-baseline --no-module-api > gpurun_out/bench_c5.log 2>&1; echo "c5 rc=$?"; tail -c 300 gpurun_out/bench_c5.log
