#!/bin/bash
# 2-GPU (or N-GPU) check of the fused all-reduce: tests, then the bench line with the peer-mailbox path and with NCCL
# usage: gpurun --gpus N -- bash tuning/multigpu_r2.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests/test_gpu_multigpu.py -x -q 2>&1 | tail -15 > $O/r2_multigpu_tests_n$N.txt; cat $O/r2_multigpu_tests_n$N.txt
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 20 --warmup 5 "${@:3}" > $O/$2.json 2> $O/$2.err; echo "$2 rc=$?"; tail -3 $O/$2.err; }
run 29611 r2_bench_n${N}_fused
run 29612 r2_bench_n${N}_nccl --nccl-allreduce
python - <<PY
import json
for f in ("fused","nccl"):
    try:
        d=json.loads(open("$O/r2_bench_n${N}_%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d.get("parity"), {k:v for k,v in d.get("config",{}).items() if "reduce" in k or "comm" in k}, d.get("strong"))
    except Exception as e: print(f, "ERR", e)
PY
