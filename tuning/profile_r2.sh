#!/bin/bash
# Round-2 evidence run (one B200): the bench line, launch lists and ncu --set full captures of the dominant kernel of every
# BASELINE config.  The .ncu-rep files are summarised ON the box (profiles/summarize.py) and deleted: gpurun_out/ is capped at
# 64 MiB.   usage (from the repo root):  gpurun -- bash tuning/profile_r2.sh [tests]
set -u
mkdir -p gpurun_out
O=gpurun_out
if [ "${1:-}" = "tests" ]; then python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/r2_gputests.txt; cat $O/r2_gputests.txt; fi
python bench.py --steps 20 --warmup 5 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; echo "bench rc=$?"
NCU="ncu --clock-control none"
cap() {   # cap <name> <kernel regex> <skip> <command...>: one --set full capture -> summary files, report removed
    local name=$1 regex=$2 skip=$3; shift 3
    $NCU --set full --import-source on -k regex:$regex -s $skip -c 1 -o $O/$name "$@" > /dev/null 2>&1
    if [ -f $O/$name.ncu-rep ]; then python profiles/summarize.py $O/$name.ncu-rep $O/$name > /dev/null 2>&1; rm -f $O/$name.ncu-rep; else echo "no report for $name"; fi
}
# launch lists (per-launch gpu__time_duration; cold-cache, serialised: the SHARES are what must agree with the bench)
$NCU --metrics gpu__time_duration.sum -c 120 --csv --log-file $O/r2_launches.csv python bench.py --steps 2 --warmup 3 > /dev/null 2>&1
cap r2_reverse tsit5_reverse_kernel 3 python bench.py --steps 2 --warmup 3 --no-secondary
cap r2_forward tsit5_forward_kernel 3 python bench.py --steps 2 --warmup 3 --no-secondary
cap r2_c4_tc_reverse mlp_tc_reverse_kernel 2 python bench.py --workload c4 --steps 2 --warmup 3
cap r2_c4_tcw_reverse mlp_tcw_reverse_kernel 2 python bench.py --workload c4 --members 65536 --steps 2 --warmup 3
cap r2_c3_quadrature ros23_quadrature_kernel 1 python tuning/c3_quick.py
cap r2_c3_reverse ros23_reverse_kernel 1 python tuning/c3_quick.py
cap r2_c5_backsolve sde_backsolve_kernel 2 python bench.py --workload c5 --steps 2 --warmup 3
for w in c2f32 c3 c4 c5; do python bench.py --workload $w --steps 10 --warmup 3 > $O/r2_bench_$w.json 2>/dev/null; done
python bench.py --workload c4 --members 18944 --steps 10 --warmup 3 > $O/r2_bench_c4_n18944.json 2>/dev/null
python bench.py --workload c4 --members 65536 --steps 10 --warmup 3 > $O/r2_bench_c4_n65536.json 2>/dev/null
python tuning/ckpt_time.py > $O/r2_ckpt_time.txt 2>&1; cat $O/r2_ckpt_time.txt
du -sh $O; ls $O | head -60
