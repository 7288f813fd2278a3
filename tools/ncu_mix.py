#!/usr/bin/env python
"""Instruction mix + stall reasons of one kernel from `ncu -i X.ncu-rep --page source --csv` output.
usage: ncu_mix.py source.csv [particles_per_launch]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
npart = float(sys.argv[2]) if len(sys.argv) > 2 else 134217728.0
h = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[h]
ix = {k: i for i, k in enumerate(hdr)}
data = []
for r in rows[h + 1:]:
    if len(r) != len(hdr) or not r[0].startswith("0x"):
        break
    data.append(r)
tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
tot_samp = sum(int(r[ix["# Samples"]]) for r in data)
print("warp-instructions %d  (%.1f per 32 particles)  samples %d  sass lines %d" % (tot_inst, tot_inst / (npart / 32), tot_samp, len(data)))
c, s = Counter(), Counter()
for r in data:
    toks = r[ix["Source"]].strip().split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "LDG", "STG", "STS", "ATOM", "RED", "SHFL")) else op.split(".")[0]
    c[op] += int(r[ix["Instructions Executed"]])
    s[op] += int(r[ix["# Samples"]])
for op, n in c.most_common(32):
    print("%-14s %6.2f%% inst  %6.2f%% samples  per32p=%.1f" % (op, 100 * n / tot_inst, 100 * s[op] / max(tot_samp, 1), n / (npart / 32)))
for k in hdr:
    if k.startswith("stall_") and "Not Issued" not in k:
        v = sum(int(r[ix[k]]) for r in data)
        if v * 200 > tot_samp:
            print("%-24s %.1f%%" % (k, 100 * v / tot_samp))
