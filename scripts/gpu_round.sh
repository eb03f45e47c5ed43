#!/bin/bash
# full GPU test suite + train/decode bench + a small ncu --set full capture of the attention kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/t_all.log 2>&1
echo "tests rc=$?"; tail -n 12 gpurun_out/t_all.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_train5.log 2>&1
echo "bench rc=$?"; tail -c 2600 gpurun_out/bench_train5.log
timeout 600 python bench.py --workload decode --pieces 1024 --steps 1 --warmup 1 > gpurun_out/bench_decode_1024.log 2>&1
echo "decode1024 rc=$?"; tail -c 1200 gpurun_out/bench_decode_1024.log
timeout 300 python scripts/prof_kernels.py "enc" > gpurun_out/prof_enc.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd_tc|attn_bwd_d" -c 6 -o gpurun_out/prof_attn_r1 python scripts/prof_kernels.py "enc" > gpurun_out/ncu_attn.log 2>&1
echo "ncu rc=$?"; tail -n 2 gpurun_out/ncu_attn.log; ls -la gpurun_out/*.ncu-rep
