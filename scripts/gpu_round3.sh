#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/t_all.log 2>&1
echo "tests rc=$?"; tail -n 5 gpurun_out/t_all.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_train7.log 2>&1
echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_train7.log
timeout 600 python bench.py --steps 5 --warmup 3 --batch 64 --seq 2048 --tgt 2048 --no-cpu-baseline > gpurun_out/bench_train_c3shape.log 2>&1
echo "c3 rc=$?"; tail -c 900 gpurun_out/bench_train_c3shape.log
timeout 600 python bench.py --steps 5 --warmup 3 --batch 2 --seq 4096 --tgt 4096 --d-model 768 --nhead 12 --layers 12 --ff 3072 --no-cpu-baseline > gpurun_out/bench_train_c5shape.log 2>&1
echo "c5 rc=$?"; tail -c 900 gpurun_out/bench_train_c5shape.log
timeout 600 python bench.py --workload decode --pieces 1024 --steps 1 --warmup 1 > gpurun_out/bench_decode_1024.log 2>&1
echo "decode rc=$?"; tail -c 1300 gpurun_out/bench_decode_1024.log
timeout 300 python scripts/prof_kernels.py "layernorm" > gpurun_out/prof_ln.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ln_fwd_kernel|ln_bwd_kernel" -c 4 -o gpurun_out/prof_ln_r1 python scripts/prof_kernels.py "layernorm" > gpurun_out/ncu_ln.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_ln_r1.ncu-rep
