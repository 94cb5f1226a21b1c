#!/usr/bin/env python
"""Attribute the stall samples / executed instructions of one profiled kernel to SOURCE LINES.
ncu's source page of a report gives per-SASS-instruction samples; nvdisasm -g of the library built from the same
sources gives the file:line of each SASS instruction; the two are joined by instruction index (and checked by opcode).
Usage: python profiles/lines.py REPORT.ncu-rep KERNEL_NAME [LIB.so] [top]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, kern = os.path.abspath(sys.argv[1]), sys.argv[2]
lib = os.path.abspath(sys.argv[3]) if len(sys.argv) > 3 else os.path.join(ROOT, "lz4_b200", "liblz4_b200.so")
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.startswith("lz4_kernels")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
ins, cur, on = [], ("?", 0), False
for ln in dis.splitlines():
    if ln.startswith(".text."):
        on = kern in ln
        continue
    if not on:
        continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)(.*)', ln)
    if m:
        if "inlined at" not in m.group(3) or True:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        ins.append((cur, m.group(2).strip()))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
i_src, i_smp, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = []
for r in rows[h + 1:]:
    if r and r[0] == "Address":
        break                                   # next launch of the same kernel
    if len(r) > i_ex and r[0].startswith("0x"):
        data.append((r[i_src].strip(), int(r[i_smp] or 0), int(r[i_ex] or 0)))
if len(data) != len(ins):
    print("# warning: %d profiled instructions, %d in the library (sources changed since the profile?)" % (len(data), len(ins)))
agg = {}
bad = 0
for (loc, txt), (s, smp, ex) in zip(ins, data):
    if txt.split()[0].split(".")[0] != s.split()[0].split(".")[0] and not s.startswith("@"):
        bad += 1
    a = agg.setdefault(loc, [0, 0, 0])
    a[0] += smp; a[1] += ex; a[2] += 1
if bad:
    print("# warning: %d opcodes differ between report and library" % bad)
ts, te = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print("# %s: %d samples, %.1f M warp-instructions" % (kern, ts, te / 1e6))
srcs = {}
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        p = os.path.join(ROOT, "lz4_b200", "csrc", f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = srcs[f][l - 1].strip()[:100] if 0 < l <= len(srcs[f]) else ""
    print("%5.1f%% smp %5.1f%% inst %4d sass  %s:%d  %s" % (100.0 * a[0] / max(ts, 1), 100.0 * a[1] / max(te, 1), a[2], f, l, text))
