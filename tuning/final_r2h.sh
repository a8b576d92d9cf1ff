#!/bin/bash
# final evidence of the re-entry session: default bench line, C1 launch list and one --set full capture of the C1 reverse kernel
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2c_bench_n1.json 2> $O/r2c_bench_n1.err; echo "bench rc=$?"
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 24 --csv --log-file $O/r2c_c1_launches.csv python bench.py --workload c1 --steps 2 --warmup 3 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:t5a_reverse_kernel -s 2 -c 1 -o $O/r2c_c1_reverse_v4 python bench.py --workload c1 --steps 2 --warmup 3 > /dev/null 2>&1
if [ -f $O/r2c_c1_reverse_v4.ncu-rep ]; then python profiles/summarize.py $O/r2c_c1_reverse_v4.ncu-rep $O/r2c_c1_reverse_v4 > /dev/null 2>&1; rm -f $O/r2c_c1_reverse_v4.ncu-rep; fi
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c_bench_n1.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "parity", d["parity"]["ok"], "clocks", d["clocks"])
for k, v in d.get("secondary", {}).items():
    print(k, v.get("error") or (v["ms_per_step"], (v.get("parity") or {}).get("ok"), (v.get("cpu_baseline") or {}).get("value")))
PY
ls -la $O | head -30
