#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/r2g_gputests.txt; tail -6 $O/r2g_gputests.txt
timeout 200 python bench.py --workload c1 --steps 5 --warmup 3 > $O/r2g_bench_c1.json 2>/dev/null; echo "c1 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2g_bench_c1.json')); print(d['ms_per_step'], d['phases_ms'], d.get('parity'))"
