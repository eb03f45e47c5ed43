#!/bin/bash
# GPU box: tests that failed before, then the train bench and (only if it exited 0) the ncu launch list.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q --timeout 300 -p no:cacheprovider -k "sampler or grads_vs_reference or sampled_decode" > gpurun_out/t_fix.log 2>&1
echo "fix rc=$?"; tail -n 30 gpurun_out/t_fix.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_train.log 2>&1
rc=$?; echo "bench rc=$rc"; tail -c 6000 gpurun_out/bench_train.log
if [ $rc -eq 0 ]; then
  timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
  echo "ncu rc=$?"; tail -n 5 gpurun_out/ncu.log
fi
timeout 900 python bench.py --workload decode --steps 1 --warmup 1 --pieces 256 > gpurun_out/bench_decode.log 2>&1
echo "decode rc=$?"; tail -c 3000 gpurun_out/bench_decode.log
