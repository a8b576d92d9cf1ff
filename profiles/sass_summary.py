#!/usr/bin/env python
"""SASS evidence for the shipped library: per kernel, the counts of the mnemonics that prove which hardware paths the code
uses (B200_PROFILING.md): UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA
bulk copy), SYNCS = mbarrier, REDUX = warp reduce, DFMA/DMUL/DADD = fp64 pipe, MUFU.TANH.
usage: python profiles/sass_summary.py [lib.so] > profiles/r2_sass_summary.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "scimlsensitivity.jl_b200/libb200adj.so"
KEYS = ["UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "SYNCS", "REDUX", "CREDUX", "DFMA", "DMUL", "DADD", "MUFU.TANH", "SHFL", "BAR.SYNC", "LDG", "STG"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda names: subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
counts, cur, archs = collections.OrderedDict(), None, set()
for line in sass.split("\n"):
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        archs.add(m.group(1))
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                counts[cur][k] += 1
        counts[cur]["_total"] += 1
names = demangle(list(counts))
print(f"# cuobjdump -sass {lib}: {len(counts)} kernels, arch {sorted(archs)}")
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("# library totals: " + ", ".join(f"{k} {tot[k]}" for k in KEYS if tot[k]))
print("# per kernel (only non-zero columns; kernels grouped by template name)")
groups = collections.OrderedDict()
for mang, name in zip(counts, names):
    base = re.sub(r"<.*", "", name.replace("void ", "")).strip()
    g = groups.setdefault(base, dict(n=0, c=collections.Counter(), ex=name))
    g["n"] += 1; g["c"].update(counts[mang])
for base, g in groups.items():
    c = g["c"]
    print(f"{base:40s} x{g['n']:<4d} instr {c['_total']:>8d}  " + "  ".join(f"{k} {c[k]}" for k in KEYS if c[k]))
# the headline kernels individually
print("# headline instantiations")
for mang, name in zip(counts, names):
    if ("tsit5_reverse_kernel<b200adj::Lorenz, 1, true, 1, false, double, false>" in name or "mlp_tc_reverse_kernel<1>" in name
            or "mlp_tc_forward_kernel" in name or "ros23_quadrature_kernel<b200adj::Robertson, false>" in name
            or "sde_backsolve_kernel<b200adj::SdeLotkaVolterra<true>, false, true, 1, false>" in name):
        c = counts[mang]
        print(f"{name[:150]}\n    instr {c['_total']}  " + "  ".join(f"{k} {c[k]}" for k in KEYS if c[k]))
