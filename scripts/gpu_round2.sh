#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/t_all.log 2>&1
echo "tests rc=$?"; tail -n 6 gpurun_out/t_all.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train6.log 2>&1
echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_train6.log
timeout 300 python scripts/prof_kernels.py > gpurun_out/prof_plain.log 2>&1; cat gpurun_out/prof_plain.log | tail -15
timeout 600 python bench.py --workload decode --pieces 1024 --steps 2 --warmup 1 > gpurun_out/bench_decode_1024.log 2>&1
echo "decode rc=$?"; tail -c 1500 gpurun_out/bench_decode_1024.log
