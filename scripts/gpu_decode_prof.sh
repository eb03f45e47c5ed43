#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload decode --pieces 1024 --steps 2 --warmup 1 > gpurun_out/bench_decode_1024b.log 2>&1
echo "decode1024 rc=$?"; tail -c 1300 gpurun_out/bench_decode_1024b.log
timeout 600 python bench.py --workload decode --pieces 1024 --decode-len 48 --steps 1 --warmup 0 > gpurun_out/plain_dec.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_decode.csv python bench.py --workload decode --pieces 1024 --decode-len 48 --steps 1 --warmup 0 > gpurun_out/ncu_dec.log 2>&1
echo "ncu rc=$?"
python scripts/summarize_launches.py gpurun_out/launches_decode.csv | head -16
