#!/usr/bin/env python
"""Turn gpurun_out/{launches_TAG.csv, prof_TAG.ncu-rep} into the committed summaries under profiles/:
   profiles/launches_TAG.csv      per-launch gpu__time_duration (ncu --metrics pass, cold cache, serialised)
   profiles/ncu_TAG_summary.txt   key raw metrics + stall mix + hottest SASS lines of the dominant kernel
   profiles/traffic.json          dram bytes per launch of the dominant kernel, scaled to the bench batch
Usage: python profiles/summarize.py TAG [blocks_in_profiled_run] [blocks_in_bench_batch]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
prof_blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
bench_blocks = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
out = os.path.join(ROOT, "gpurun_out")
rep = os.path.join(out, "prof_%s.ncu-rep" % tag)

# ---- launch list ----
rows = [r for r in csv.reader(l for l in open(os.path.join(out, "launches_%s.csv" % tag)) if not l.startswith("=="))]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
with open(os.path.join(ROOT, "profiles", "launches_%s.csv" % tag), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 2 --warmup 3 (4 GiB batch)\n")
    f.write("kernel,duration_ns\n")
    tot = {}
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("<unnamed>::", "")
        f.write("%s,%s\n" % (name, r[vi]))
        tot.setdefault(name, []).append(float(r[vi].replace(",", "")))
    f.write("# share of the decode step (scan + expand_fast + expand_generic), mean per launch:\n")
    step = sum(sum(v) / len(v) for k, v in tot.items() if k.startswith("lz4_scan") or k.startswith("lz4_expand"))
    for k, v in tot.items():
        if k.startswith("lz4_scan") or k.startswith("lz4_expand"):
            f.write("#   %s: %.1f us  (%.1f %% of the step)\n" % (k, sum(v) / len(v) / 1e3, 100 * sum(v) / len(v) / step))

# ---- raw metrics ----
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h, v = rr[0], rr[2] if len(rr) > 2 else rr[1]
m = dict(zip(h, v))
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
lines = ["ncu --set full --clock-control none --import-source on -k regex:expand_fast  (bench.py --gib 0.5: %d blocks of 64 KB)" % prof_blocks, ""]
for k in keys:
    if k in m:
        lines.append("%-85s %s" % (k, m[k]))
lines.append("")
lines.append("warp stall samples (smsp__pcsamp_warps_issue_stalled_*):")
for k in sorted(h):
    if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k:
        lines.append("  %-30s %s" % (k.replace("smsp__pcsamp_warps_issue_stalled_", ""), m[k]))

# units: dram bytes are reported in MB by this ncu build for this size
def num(x):
    return float(x.replace(",", ""))
unit_scale = 1e6
rd, wr = num(m["dram__bytes_read.sum"]) * unit_scale, num(m["dram__bytes_write.sum"]) * unit_scale
per_block = (rd + wr) / prof_blocks
traffic = {"expand_dram_bytes_per_launch": int(per_block * bench_blocks), "source": "ncu --set full, tag %s" % tag,
           "profiled_blocks": prof_blocks, "dram_read_bytes": int(rd), "dram_write_bytes": int(wr),
           "dram_bytes_per_block": round(per_block, 1)}
json.dump(traffic, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
lines.append("")
lines.append("dram bytes per 64 KB block: %.0f  (algorithmic C+U for P50 = %.0f)" % (per_block, 6945762471 / 65536))

# ---- hottest SASS lines ----
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
sr = list(csv.reader(io.StringIO(src)))
sh = sr[1]
i_src, i_smp, i_ex = sh.index("Source"), sh.index("# Samples"), sh.index("Instructions Executed")
data = [(int(r[i_smp]), int(r[i_ex]), r[i_src].strip()) for r in sr[2:] if len(r) > i_ex]
lines.append("")
lines.append("SASS: %d instructions in the kernel, %.0f M warp-instructions executed, %d stall samples" %
             (len(data), sum(d[1] for d in data) / 1e6, sum(d[0] for d in data)))
lines.append("hottest SASS lines (samples, executions, instruction):")
for smp, ex, s in sorted(data, reverse=True)[:25]:
    lines.append("  %7d %11d  %s" % (smp, ex, s[:100]))
tma = [d for d in data if "UBLKCP" in d[2] or "UTMA" in d[2] or "SYNCS" in d[2]]
lines.append("")
lines.append("TMA / mbarrier instructions present in SASS: " + ", ".join(sorted(set(d[2].split()[0] if not d[2].startswith("@") else d[2].split()[1] for d in tma))))
open(os.path.join(ROOT, "profiles", "ncu_%s_summary.txt" % tag), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:45]))
