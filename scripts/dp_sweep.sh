#!/bin/bash
# N-GPU sweep of the data-parallel knobs, one bench line per setting:
#   dp_sweep.sh N "RESERVED_SMS NCCL_MAX_CTAS HIGH_PRIORITY [BUCKETS]" ...     (0 = NCCL's own CTA count / one bucket per layer)
mkdir -p gpurun_out
n=${1:-8}; shift
port=29520
for cfg in "$@"; do
  set -- $cfg; r=$1; c=$2; hp=$3; nb=${4:-0}
  port=$((port+1))
  if [ "$c" = "0" ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$c; fi
  SMER_DP_BUCKETS=$nb SMER_NCCL_HIGH_PRIORITY=$hp SMER_RESERVED_SMS=$r timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $n --steps 10 --warmup 3 --skip decode+c3+c5_attention+c1+torch_gpu_baseline+padded_layout --no-module-api --no-cpu-baseline \
    > gpurun_out/dp_sweep_r${r}_c${c}_hp${hp}_b${nb}.json 2> gpurun_out/dp_sweep_r${r}_c${c}_hp${hp}_b${nb}.log
  echo "reserved=$r nccl_ctas=$c high_priority=$hp buckets=$nb rc=$?"
done
