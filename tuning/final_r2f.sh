#!/bin/bash
# 2-GPU pass: multi-GPU parity test + the driver-style bench line at N = 2 (all secondary legs, fused all-reduce)
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_multigpu.py -q 2>&1 | tail -5 > $O/r2f_multigpu_tests.txt; cat $O/r2f_multigpu_tests.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2f_bench_n2.json 2> $O/r2f_bench_n2.err; echo "bench n2 rc=$?"; tail -4 $O/r2f_bench_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2f_bench_n2.json"))
print("value", d["value"], "ms", d["ms_per_step"], "allreduce", d["allreduce"], "parity", d["parity"], "strong", d.get("strong", {}).get("ms_per_step"))
for k, v in d.get("secondary", {}).items():
    print(k, v.get("error") or (v["ms_per_step"], (v.get("parity") or {}).get("dp_rel"), (v.get("parity") or {}).get("ok"), (v.get("cpu_baseline") or {}).get("value")))
PY
