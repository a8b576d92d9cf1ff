#!/bin/bash
# round-2 re-entry GPU pass: full GPU test-suite (new: reference-held numbers, Relax family, non-linear affect), smoke,
# default bench line (new: C1-ensemble and C3 secondary legs), reference arm
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -25 > $O/r2c_gputests.txt; tail -5 $O/r2c_gputests.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/r2c_smoke.txt
timeout 300 python bench.py --steps 20 --warmup 5 > $O/r2c_bench_n1.json 2> $O/r2c_bench_n1.err; echo "bench rc=$?"; tail -3 $O/r2c_bench_n1.err
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2c_bench_reference.json 2>/dev/null; echo "reference rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c_bench_n1.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "parity", d["parity"]["ok"])
for k, v in d.get("secondary", {}).items():
    print(k, v.get("error") or (v["ms_per_step"], (v.get("parity") or {}).get("dp_rel"), (v.get("parity") or {}).get("ok")))
PY
