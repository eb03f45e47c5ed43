#!/bin/bash
# N-GPU data-parallel train bench (torchrun) + decode bench; N from $1
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_train_dp$N.log 2>&1
echo "dp$N rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/bench_train_dp$N.log | tail -c 1500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload decode --pieces 1024 --steps 1 --warmup 1 > gpurun_out/bench_decode_dp$N.log 2>&1
echo "decode dp$N rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/bench_decode_dp$N.log | tail -c 1300
