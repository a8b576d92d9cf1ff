#!/usr/bin/env python
"""Turn an ncu report (gpurun_out/*.ncu-rep, captured with --set full --import-source on) into the small tracked
summaries under profiles/: key metrics (json + txt) and the top stall sites of the source page.
usage: python profiles/summarize.py gpurun_out/r1_reverse.ncu-rep profiles/r1_reverse"""
import collections
import csv
import os
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum",
    "sm__cycles_active.avg", "sm__cycles_active.min", "sm__cycles_active.max", "sm__cycles_elapsed.avg",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__icc_request_hit_rate.pct",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2]


def main(rep, prefix):
    hdr, units, vals = raw(rep)
    m = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS or ("issue_stalled" in h and h.endswith("_per_issue_active.ratio")) or h == "Kernel Name":
            try:
                m[h] = {"value": float(v.replace(",", "")), "unit": u}
            except ValueError:
                m[h] = {"value": v, "unit": u}
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr2 = rows[1]
    ix = {h: i for i, h in enumerate(hdr2)}
    ops = collections.Counter()
    stall_tot = collections.Counter()
    sites = []
    for r in rows[2:]:
        s = r[ix["Source"]]
        toks = s.split()
        op = (toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "")).split(".")[0]
        ops[op] += int(r[ix["Instructions Executed"]] or 0)
        for k in hdr2:
            if k.startswith("stall_") and "(Not Issued)" not in k:
                stall_tot[k] += int(r[ix[k]] or 0)
        sites.append((int(r[ix["# Samples"]] or 0), r[ix["Address"]][-6:], s[:80]))
    m["opcode_mix_warp_instructions"] = dict(ops.most_common(16))
    m["stall_samples"] = dict(stall_tot.most_common(10))
    m["top_sample_sites"] = [dict(samples=a, addr=b, sass=c) for a, b, c in sorted(sites, reverse=True)[:10]]
    # CUDA-C correlation of the same page (needs -lineinfo + --import-source on): samples per source line
    m["top_source_lines"] = []
    try:
        cs = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
        agg = collections.Counter()
        text = {}
        fname, h3 = "", None
        for r in csv.reader(io.StringIO(cs)):
            if len(r) == 2 and r[0] == "File Path":
                fname, h3 = os.path.basename(r[1]), None
            elif len(r) > 6 and r[0] == "Line No":
                h3 = r.index("# Samples")
            elif h3 is not None and len(r) > h3 and r[0].isdigit():
                if r[1].strip():
                    text[(fname, int(r[0]))] = r[1].strip()[:110]
                try:
                    agg[(fname, int(r[0]))] += int(r[h3] or 0)
                except ValueError:
                    pass
        m["top_source_lines"] = [dict(samples=n, file=k[0], line=k[1], source=text.get(k, "")) for k, n in agg.most_common(25) if n]
    except Exception as e:                                   # older ncu: no CUDA correlation
        m["top_source_lines_error"] = str(e)
    with open(prefix + "_ncu_summary.json", "w") as f:
        json.dump(m, f, indent=1)
    with open(prefix + "_ncu_summary.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on  ({rep})\n")
        for k, v in m.items():
            if isinstance(v, dict) and "value" in v:
                f.write(f"{k:85s} {v['value']!s:>22s} {v['unit']}\n")
        f.write("\n# warp instructions executed by opcode\n")
        for k, v in m["opcode_mix_warp_instructions"].items():
            f.write(f"  {k:12s} {v}\n")
        f.write("\n# warp stall samples (all)\n")
        for k, v in m["stall_samples"].items():
            f.write(f"  {k:24s} {v}\n")
        f.write("\n# top sampled instructions\n")
        for s in m["top_sample_sites"]:
            f.write(f"  {s['samples']:7d} {s['addr']} {s['sass']}\n")
        if m["top_source_lines"]:
            f.write("\n# top sampled source lines (CUDA-C view)\n")
            for s_ in m["top_source_lines"]:
                f.write(f"  {s_['samples']:7d} {s_['file']}:{s_['line']}  {s_['source']}\n")
    print("wrote", prefix + "_ncu_summary.{json,txt}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
