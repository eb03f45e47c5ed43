#!/bin/bash
# plain run first, then one ncu --set full capture of the attention kernels (source counters included)
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
timeout 300 python scripts/prof_kernels.py attn_ > gpurun_out/prof_attn_plain.log 2>&1
rc=$?; echo "plain rc=$rc"; cat gpurun_out/prof_attn_plain.log
if [ $rc -eq 0 ]; then
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"attn_(fwd|bwd)" -c 13 -o gpurun_out/prof_attn_r1m python scripts/prof_kernels.py attn_ > gpurun_out/ncu_attn.log 2>&1
  echo "ncu rc=$?"; tail -n 3 gpurun_out/ncu_attn.log; ls -la gpurun_out/*.ncu-rep
fi
