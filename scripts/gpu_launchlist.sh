#!/bin/bash
# ncu launch list (gpu__time_duration per launch) of the train bench, after the same command ran clean without ncu
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-module-api --no-graph"
timeout 600 $CMD > gpurun_out/launch_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/launch_plain.log; exit 1; }
tail -c 300 gpurun_out/launch_plain.log
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_train.csv $CMD > gpurun_out/launch_ncu.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/launches_train.csv
