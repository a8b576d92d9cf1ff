#!/bin/bash
# GPU pass after: hybrid neural ODE events (MLP), non-linear affect, Relax family
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $O/r2d_gputests.txt; tail -6 $O/r2d_gputests.txt
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python tuning/san_small.py r3 > $O/r2d_sanitizer_memcheck_r3.log 2>&1
echo "memcheck r3 rc=$? $(grep -c 'ERROR SUMMARY: 0 errors' $O/r2d_sanitizer_memcheck_r3.log)"
tail -c 2500 $O/r2d_sanitizer_memcheck_r3.log > $O/x && mv $O/x $O/r2d_sanitizer_memcheck_r3.log
