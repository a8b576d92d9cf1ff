#!/bin/bash
# compute-sanitizer over the small end-to-end cases (tuning/san_small.py); logs -> gpurun_out/r2_sanitizer_<tool>_<case>.log
# usage: gpurun -- bash tuning/sanitize_r2.sh
set -u
mkdir -p gpurun_out
for c in ode mlp adaptive sde r2; do
  timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python tuning/san_small.py $c > gpurun_out/r2_sanitizer_memcheck_$c.log 2>&1
  echo "memcheck $c rc=$? $(grep -c 'ERROR SUMMARY: 0 errors' gpurun_out/r2_sanitizer_memcheck_$c.log)"
done
for c in ode mlp r2; do
  timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python tuning/san_small.py $c > gpurun_out/r2_sanitizer_racecheck_$c.log 2>&1
  echo "racecheck $c rc=$? $(tail -1 gpurun_out/r2_sanitizer_racecheck_$c.log)"
done
for f in gpurun_out/r2_sanitizer_*.log; do tail -c 3000 $f > $f.tail; mv $f.tail $f; done
