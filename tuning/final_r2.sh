#!/bin/bash
# final round-2 GPU pass: full -m gpu suite, sanitizer logs, the default bench line, the C1-ensemble line
set -u
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r2_gputests.txt; cat $O/r2_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/r2_smoke.txt
python bench.py --steps 20 --warmup 5 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; echo "bench rc=$?"
python bench.py --workload c1 --steps 5 --warmup 3 > $O/r2_bench_c1.json 2> $O/r2_bench_c1.err; echo "c1 rc=$?"; cat $O/r2_bench_c1.json | cut -c1-1500
python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_reference.json 2>/dev/null; echo "ref rc=$?"
bash tuning/sanitize_r2.sh
du -sh $O
