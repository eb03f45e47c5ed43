#!/bin/bash
# One runner for everything that is executed on the GPU box (through `gpurun -- bash scripts/gpu.sh <steps...>`).
# Every step writes under gpurun_out/ (the only directory that travels back) and is bounded by its own timeout, so a
# hung kernel costs that step, not the box.  Steps (any number, in order):
#   tests[:EXPR]          pytest -m gpu (optionally -k EXPR)                     -> gpurun_out/tests.log
#   smoke                 __graft_entry__.smoke()                                -> gpurun_out/smoke.log
#   bench[:ARGS]          python bench.py ARGS (comma = space)                   -> gpurun_out/bench_<n>.json/.log
#   prof:FILTER           scripts/prof_kernels.py FILTER (CUDA-event kernel times) -> gpurun_out/prof_<FILTER>.log
#   ab:FILTER             same, alternating _ab/libsmer_b200_A.so (scripts/build_ab.sh REV) and the in-tree build
#   launches[:ARGS]       ncu launch list of `bench.py --steps 1 --warmup 1 ARGS` (after a plain run exits 0)
#   ncu:REGEX:FILTER[:N]  one `ncu --set full` capture of kernels matching REGEX from prof_kernels.py FILTER
#   memcheck:EXPR         compute-sanitizer memcheck on pytest -k EXPR
mkdir -p gpurun_out
n=0
for step in "$@"; do
  kind=${step%%:*}; arg=""; [[ "$step" == *:* ]] && arg=${step#*:}
  n=$((n+1))
  echo "=== [$n] $step"
  case $kind in
    tests)
      if [ -n "$arg" ]; then timeout ${TLIMIT:-300} python -m pytest tests -m gpu -q -x --timeout 200 -p no:cacheprovider -k "$arg" > gpurun_out/tests_$n.log 2>&1
      else timeout 1800 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/tests_$n.log 2>&1; fi
      echo "rc=$?"; tail -n 40 gpurun_out/tests_$n.log ;;
    smoke)
      timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -n 12 gpurun_out/smoke.log ;;
    bench)
      timeout 1500 python bench.py ${arg//,/ } > gpurun_out/bench_$n.json 2> gpurun_out/bench_$n.log
      echo "rc=$?"; tail -c 1500 gpurun_out/bench_$n.log; tail -c 6000 gpurun_out/bench_$n.json ;;
    prof)
      timeout 120 python scripts/prof_kernels.py $arg > gpurun_out/prof_${arg:-all}.log 2>&1; echo "rc=$?"; tail -n 40 gpurun_out/prof_${arg:-all}.log ;;
    ab)
      for i in 1 2; do
        echo "== A (old)"; SMER_B200_LIB=$PWD/_ab/libsmer_b200_A.so timeout 300 python scripts/prof_kernels.py $arg 2>&1 | tail -n 24
        echo "== B (new)"; timeout 300 python scripts/prof_kernels.py $arg 2>&1 | tail -n 24
      done | tee gpurun_out/ab_${arg:-all}.log ;;
    launches)
      if timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline ${arg//,/ } > gpurun_out/plain_$n.log 2>&1; then
        timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_$n.csv \
          python bench.py --steps 1 --warmup 1 --no-cpu-baseline ${arg//,/ } > gpurun_out/ncu_$n.log 2>&1
        echo "ncu rc=$?"; tail -n 3 gpurun_out/ncu_$n.log
      else echo "plain run failed"; tail -n 20 gpurun_out/plain_$n.log; fi ;;
    ncu)
      IFS=: read -r regex filter count <<< "$arg"
      if timeout 300 python scripts/prof_kernels.py $filter > gpurun_out/plain_$n.log 2>&1; then
        timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$regex" -c ${count:-8} -o gpurun_out/ncu_$n \
          python scripts/prof_kernels.py $filter > gpurun_out/ncu_$n.log 2>&1
        echo "ncu rc=$?"; tail -n 3 gpurun_out/ncu_$n.log; ls -la gpurun_out/ncu_$n.ncu-rep
      else echo "plain run failed"; tail -n 20 gpurun_out/plain_$n.log; fi ;;
    ncuq)
      # quick counters (one or two replay passes): ncuq:REGEX:FILTER[:N]
      IFS=: read -r regex filter count <<< "$arg"
      timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__cycles_elapsed.max \
        --clock-control none -k regex:"$regex" -c ${count:-8} --csv --log-file gpurun_out/ncuq_$n.csv python scripts/prof_kernels.py $filter > gpurun_out/ncuq_$n.log 2>&1
      echo "ncu rc=$?"; python - <<PYEOF
import csv
rows=[r for r in csv.reader(open("gpurun_out/ncuq_$n.csv")) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
cur=None
for r in rows[1:]:
    key=(r[ix["ID"]], r[ix["Kernel Name"]][:48])
    if key!=cur: print("==", key); cur=key
    print("    %-70s %s" % (r[ix["Metric Name"]], r[ix["Metric Value"]]))
PYEOF
      ;;
    memcheck)
      timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "$arg" > gpurun_out/memcheck_$n.log 2>&1
      echo "rc=$?"; tail -n 30 gpurun_out/memcheck_$n.log ;;
    *) echo "unknown step $step" ;;
  esac
done
