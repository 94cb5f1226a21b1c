#!/usr/bin/env python
"""Turn gpurun_out/{launches_TAG.csv, prof_TAG.ncu-rep} into the committed summaries under profiles/:
   profiles/launches_TAG.csv      per-launch gpu__time_duration (ncu --metrics pass, cold cache, serialised) + shares of the step
   profiles/ncu_TAG_summary.txt   per profiled kernel: key raw metrics, stall mix, hottest SASS lines
   profiles/traffic.json          dram bytes per 64 KB block of the scan and the expand kernel (bench.py scales them to its batch)
Usage: python profiles/summarize.py TAG [blocks_in_profiled_run]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
prof_blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
out = os.path.join(ROOT, "gpurun_out")
rep = os.path.join(out, "prof_%s.ncu-rep" % tag)


def num(x):
    return float(str(x).replace(",", ""))


# ---- launch list ----
lpath = os.path.join(out, "launches_%s.csv" % tag)
if os.path.exists(lpath):
    rows = [r for r in csv.reader(l for l in open(lpath) if not l.startswith("=="))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    with open(os.path.join(ROOT, "profiles", "launches_%s.csv" % tag), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e (4 GiB batch)\n")
        f.write("kernel,duration_ns\n")
        tot = {}
        for r in rows[1:]:
            name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
            f.write("%s,%s\n" % (name, r[vi]))
            tot.setdefault(name, []).append(num(r[vi]))
        dec = {k: v for k, v in tot.items() if k.startswith("lz4_scan") or k.startswith("lz4_expand")}
        # the decode launches of the timed steps are the LAST ones of each kind (warm-up and setup come first)
        step = sum(min(v) for v in dec.values())
        f.write("# share of the decode step (fastest launch of each kernel; ncu serialises and flushes caches: compare shares, not absolutes):\n")
        for k, v in dec.items():
            f.write("#   %s: %.1f us  (%.1f %% of the step), %d launches\n" % (k, min(v) / 1e3, 100 * min(v) / step, len(v)))

# ---- raw metrics, one block per profiled kernel ----
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h = rr[0]
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
lines = ["ncu --set full --clock-control none --import-source on   python bench.py --gib 0.5 --steps 2 --warmup 3 --no-cpu --no-e2e"
         "  (%d blocks of 64 KB per launch; one launch per kernel)" % prof_blocks, ""]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# the source page concatenates one table per launch, each introduced by a "Kernel Name" line
tables, cur = [], None
for row in csv.reader(io.StringIO(src)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        tables.append(cur)
    elif cur is not None:
        cur["rows"].append(row)
tpath = os.path.join(ROOT, "profiles", "traffic.json")
traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
traffic.update({"proba": 0.5, "block_bytes": 65536})
traffic.setdefault("sources", {})
seen = set()
for idx, v in enumerate(rr[2:]):
    m = dict(zip(h, v))
    name = m["Kernel Name"].replace("<unnamed>::", "").replace("void ", "")
    short = name.split("(")[0].split("<")[0]
    if short in seen:
        continue
    seen.add(short)
    lines.append("=" * 100)
    lines.append("kernel: " + name)
    for k in keys:
        if k in m:
            lines.append("  %-85s %s" % (k, m[k]))
    rd, wr = num(m["dram__bytes_read.sum"]) * 1e6, num(m["dram__bytes_write.sum"]) * 1e6      # this ncu build reports MB here
    per_block = (rd + wr) / prof_blocks
    lines.append("  dram bytes per 64 KB block: %.0f   (algorithmic C+U for P50 = %.0f)" % (per_block, 6945762471 / 65536))
    lines.append("  warp-instructions per 64 KB block: %.0f" % (num(m["smsp__inst_executed.sum"]) / prof_blocks))
    src_note = "ncu --set full, tag %s, %d blocks of 64 KB (datagen P50) per launch" % (tag, prof_blocks)
    if "expand_rows" in short:
        traffic["expand_dram_bytes_per_block"] = round(per_block, 1); traffic["sources"]["expand"] = src_note
    elif short.endswith("lz4_scan_kernel"):
        traffic["scan_dram_bytes_per_block"] = round(per_block, 1); traffic["sources"]["scan"] = src_note
    elif "encode_par" in short:
        traffic["encode_par_dram_bytes_per_block"] = round(per_block, 1); traffic["sources"]["encode_par"] = src_note
    lines.append("  stalls per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active):")
    for k in sorted(h):
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and num(m[k]) >= 0.2:
            lines.append("    %-24s %.2f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), num(m[k])))
    tb = next((t for t in tables if t["name"].replace("<unnamed>::", "").replace("void ", "") == name), None)
    if tb and len(tb["rows"]) > 2:
        sh = tb["rows"][0]
        i_src, i_smp, i_ex = sh.index("Source"), sh.index("# Samples"), sh.index("Instructions Executed")
        data = [(int(r[i_smp]), int(r[i_ex]), r[i_src].strip()) for r in tb["rows"][1:] if len(r) > i_ex]
        lines.append("  SASS: %d instructions, %.0f M warp-instructions executed, %d stall samples" %
                     (len(data), sum(d[1] for d in data) / 1e6, sum(d[0] for d in data)))
        lines.append("  hottest SASS lines (samples, executions, instruction):")
        for smp, ex, s in sorted(data, reverse=True)[:16]:
            lines.append("    %7d %11d  %s" % (smp, ex, s[:96]))
        tma = [d for d in data if any(t in d[2] for t in ("UBLKCP", "UTMA", "SYNCS", "LDGSTS"))]
        mn = sorted(set(d[2].split()[0] if not d[2].startswith("@") else d[2].split()[1] for d in tma))
        lines.append("  TMA / mbarrier / cp.async instructions in the SASS: " + (", ".join(mn) if mn else "-"))
    lines.append("")
json.dump(traffic, open(tpath, "w"), indent=1)
open(os.path.join(ROOT, "profiles", "ncu_%s_summary.txt" % tag), "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:70]))
