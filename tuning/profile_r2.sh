#!/bin/bash
# Round-2 evidence run (one B200): full GPU test-suite, the bench line, launch lists and ncu --set full captures of the
# dominant kernel of every BASELINE config.  Outputs land in gpurun_out/; profiles/summarize.py turns the reports into
# the tracked summaries.   usage (from the repo root):  gpurun -- bash tuning/profile_r2.sh
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2_gputests.txt; cat gpurun_out/r2_gputests.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
NCU="ncu --clock-control none"
# launch lists (per-launch gpu__time_duration; cold-cache, serialised: the SHARES are what must agree with the bench)
$NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 > /dev/null 2>&1
# full captures of the dominant kernels
$NCU --set full --import-source on -k regex:tsit5_reverse_kernel -s 3 -c 1 -o gpurun_out/r2_reverse python bench.py --steps 2 --warmup 3 --no-secondary > /dev/null 2>&1
$NCU --set full --import-source on -k regex:tsit5_forward_kernel -s 3 -c 1 -o gpurun_out/r2_forward python bench.py --steps 2 --warmup 3 --no-secondary > /dev/null 2>&1
$NCU --set full --import-source on -k regex:mlp_tc_reverse_kernel -s 2 -c 1 -o gpurun_out/r2_c4_tc_reverse python bench.py --workload c4 --steps 2 --warmup 3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:mlp_tcw_reverse_kernel -s 2 -c 1 -o gpurun_out/r2_c4_tcw_reverse python bench.py --workload c4 --members 65536 --steps 2 --warmup 3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:ros23_quadrature_kernel -s 1 -c 1 -o gpurun_out/r2_c3_quadrature python tuning/c3_quick.py > /dev/null 2>&1
$NCU --set full --import-source on -k regex:ros23_reverse_kernel -s 1 -c 1 -o gpurun_out/r2_c3_reverse python tuning/c3_quick.py > /dev/null 2>&1
$NCU --set full --import-source on -k regex:sde_backsolve_kernel -s 2 -c 1 -o gpurun_out/r2_c5_backsolve python bench.py --workload c5 --steps 2 --warmup 3 > /dev/null 2>&1
for w in c2f32 c3 c4 c5; do python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r2_bench_$w.json 2>/dev/null; done
python bench.py --workload c4 --members 65536 --steps 10 --warmup 3 > gpurun_out/r2_bench_c4_n65536.json 2>/dev/null
python tuning/ckpt_time.py > gpurun_out/r2_ckpt_time.txt 2>&1; cat gpurun_out/r2_ckpt_time.txt
ls -la gpurun_out | tail -30
