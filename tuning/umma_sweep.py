import os, sys, subprocess
for c in (1, 2, 4, 6, 8):
    env = dict(os.environ, B200ADJ_UMMA_CTAS_PER_SM=str(c))
    out = subprocess.run([sys.executable, "bench.py", "--workload", "c4", "--dtype", "bf16_f32acc", "--steps", "10", "--warmup", "3"], env=env, capture_output=True, text=True).stdout
    import json; d = json.loads(out.strip().splitlines()[-1]); print(c, "%.4e" % d["value"], d["phases_ms"])
