#!/bin/bash
# kernel + model + engine parity tests, per-kernel timings, short train bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_engine_gpu.py -q --timeout 300 -p no:cacheprovider -x > gpurun_out/t_quick.log 2>&1
echo "tests rc=$?"; tail -n 6 gpurun_out/t_quick.log
timeout 300 python scripts/prof_kernels.py > gpurun_out/prof_plain.log 2>&1
echo "prof rc=$?"; cat gpurun_out/prof_plain.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-module-api > gpurun_out/bench_train6.log 2>&1
echo "bench rc=$?"; grep '^{' gpurun_out/bench_train6.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value']); print({k:v['ms'] for k,v in d['kernel_families'].items()})"
