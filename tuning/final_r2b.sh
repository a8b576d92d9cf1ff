#!/bin/bash
# final round-2 GPU pass (after the TMA refactor of the adaptive Tsit5 path)
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r2_gputests.txt; cat $O/r2_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/r2_smoke.txt
timeout 400 python bench.py --steps 20 --warmup 5 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; echo "bench rc=$?"; tail -2 $O/r2_bench_n1.err
for w in c1 c3; do timeout 200 python bench.py --workload $w --steps 5 --warmup 3 > $O/r2_bench_$w.json 2>/dev/null; echo "$w rc=$?"; done
for c in adaptive r2; do
  timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python tuning/san_small.py $c > $O/r2_sanitizer_memcheck_$c.log 2>&1
  echo "memcheck $c rc=$? $(grep -c 'ERROR SUMMARY: 0 errors' $O/r2_sanitizer_memcheck_$c.log)"
done
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python tuning/san_small.py adaptive > $O/r2_sanitizer_racecheck_adaptive.log 2>&1
echo "racecheck adaptive rc=$? $(tail -1 $O/r2_sanitizer_racecheck_adaptive.log)"
for f in $O/r2_sanitizer_*.log; do tail -c 3000 $f > $f.tail; mv $f.tail $f; done
NCU="ncu --clock-control none"
$NCU --set full --import-source on -k regex:t5a_forward_kernel -s 2 -c 1 -o $O/r2_c1_forward python bench.py --workload c1 --steps 2 --warmup 3 > /dev/null 2>&1
python profiles/summarize.py $O/r2_c1_forward.ncu-rep $O/r2_c1_forward > /dev/null 2>&1; rm -f $O/*.ncu-rep
du -sh $O
