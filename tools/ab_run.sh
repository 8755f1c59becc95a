#!/bin/bash
# A/B harness used during development: op tests, per-op profiles and short bench runs under env switches.
# usage: tools/ab_run.sh OUTDIR "VAR1=.. VAR2=.."  (one variant per extra argument; run on the GPU box through gpurun)
out=${1:-gpurun_out/ab}; shift; mkdir -p $out
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q > $out/tests_ops.log 2>&1; echo "ops rc=$?"; tail -3 $out/tests_ops.log
timeout 600 python -m pytest tests/test_model_gpu.py -x -q > $out/tests_model.log 2>&1; echo "model rc=$?"; tail -3 $out/tests_model.log
i=0
for v in "$@"; do
  i=$((i+1))
  echo "== variant $i: $v"
  env $v timeout 300 python tools/profile_ops.py 4 $out/ops_b4_v$i.json > $out/ops_b4_v$i.txt 2>&1; head -1 $out/ops_b4_v$i.txt
  env $v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_v$i.json 2> $out/bench_v$i.err; cut -c1-200 $out/bench_v$i.json
done
