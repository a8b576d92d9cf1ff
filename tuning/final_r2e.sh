#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/r2e_gputests.txt; tail -8 $O/r2e_gputests.txt
