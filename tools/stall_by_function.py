#!/usr/bin/env python
"""Attribute ncu warp-stall samples and executed instructions of a kernel to source functions.

usage: stall_by_function.py <report.ncu-rep> <kernel mangled-name prefix> <warps*steps per launch>
Joins the SASS page of the report (per-instruction samples) with `nvdisasm --print-line-info` of the current
libsplendor_b200.so by instruction order (the .so must be the build that was profiled)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, prefix, denom = sys.argv[1], sys.argv[2], float(sys.argv[3])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iA, iS, iW, iE = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
KINDS = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_barrier", "stall_mio", "stall_lg", "stall_math", "stall_no_inst", "stall_branch_resolving"]
iK = [hdr.index(k) for k in KINDS]
prof, last = [], -1
for r in rows[2:]:
    if len(r) <= iE or not r[iA].startswith("0x"):
        continue
    a = int(r[iA], 16)
    if a < last:
        break
    last = a
    prof.append((r[iS].strip(), int(r[iW] or 0), int(r[iE] or 0), [int(r[i] or 0) for i in iK]))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "splendor_gym_b200", "libsplendor_b200.so")], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.startswith("spl_kernels.") and f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
sect = [x for x in re.split(r"\n\.text\.", dis) if x.startswith(prefix)][-1]
cur, lines = None, []
for line in sect.splitlines():
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m2 = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m2:
        lines.append((cur, m2.group(2).strip()))
n = min(len(prof), len(lines))
mism = sum(1 for i in range(n) if prof[i][0].split()[0].lstrip("@!P0123456789 ") != lines[i][1].split()[0].lstrip("@!P0123456789 "))
print(f"# {len(prof)} profiled / {len(lines)} disassembled instructions, {mism} opcode mismatches")
src = {f: open(os.path.join(root, "splendor_gym_b200", "csrc", f)).read().splitlines() for f in ("spl_core.cuh", "spl_kernels.cu")}


def owner(fn, ln):
    if fn not in src:
        return fn
    L = src[fn]
    for i in range(min(ln, len(L)) - 1, -1, -1):
        if re.match(r"^(SPL_HD|SPL_HD_NOINLINE|template|struct|__device__|__global__|static|\t__device__)", L[i]) and "(" in L[i]:
            m = re.search(r"(spl_\w+|operator\(\)|first|put|last|seed|next|randbelow)\s*\(", L[i])
            return m.group(1) if m else L[i][:40]
    return fn


stall, execd = collections.Counter(), collections.Counter()
kinds = collections.defaultdict(lambda: [0] * len(KINDS))
for i in range(n):
    o = owner(*lines[i][0]) if lines[i][0] else "?"
    stall[o] += prof[i][1]
    execd[o] += prof[i][2]
    for j, v in enumerate(prof[i][3]):
        kinds[o][j] += v
ts, te = sum(stall.values()), sum(execd.values())
print(f"{'function':30s} {'stall%':>7s} {'exec%':>7s} {'exec/warp-step':>14s}  " + " ".join(f"{k[6:10]:>5s}" for k in KINDS) + "   (stall kinds in % of all samples)")
for k, v in stall.most_common(24):
    print(f"{k:30s} {100*v/ts:7.1f} {100*execd[k]/te:7.1f} {execd[k]/denom:14.1f}  " + " ".join(f"{100*x/ts:5.1f}" for x in kinds[k]))
tot = [sum(kinds[k][j] for k in kinds) for j in range(len(KINDS))]
print(f"{'all':30s} {100.0:7.1f} {100.0:7.1f} {te/denom:14.1f}  " + " ".join(f"{100*x/ts:5.1f}" for x in tot))
