#!/bin/bash
# attention parity tests + per-kernel timings + short train bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q --timeout 120 -p no:cacheprovider -k "attention" > gpurun_out/t_attn.log 2>&1
echo "attn rc=$?"; tail -n 8 gpurun_out/t_attn.log
timeout 300 python scripts/prof_kernels.py > gpurun_out/prof_plain.log 2>&1
echo "prof rc=$?"; cat gpurun_out/prof_plain.log
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_engine_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/t_model.log 2>&1
echo "model rc=$?"; tail -n 8 gpurun_out/t_model.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train6.log 2>&1
echo "bench rc=$?"; grep '^{' gpurun_out/bench_train6.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value']); print({k:v['ms'] for k,v in d['kernel_families'].items()})"
