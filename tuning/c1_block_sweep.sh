for b in 32 64 128 256; do python bench.py --workload c1 --steps 5 --warmup 3 --block $b 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('block', $b, d['ms_per_step'], d['phases_ms'])"; done
