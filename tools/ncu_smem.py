#!/usr/bin/env python
"""Shared-memory wavefronts per opcode class (and ideal) of one kernel from `ncu --page source --csv`.
usage: ncu_smem.py source.csv [particles_per_launch]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
npart = float(sys.argv[2]) if len(sys.argv) > 2 else 134217728.0
h = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[h]
ix = {k: i for i, k in enumerate(hdr)}
wf, ideal, n = Counter(), Counter(), Counter()
for r in rows[h + 1:]:
    if len(r) != len(hdr) or not r[0].startswith("0x"):
        break
    w = r[ix["L1 Wavefronts Shared"]]
    if not w or w == "0":
        continue
    toks = r[ix["Source"]].strip().split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    wf[op] += int(w)
    ideal[op] += int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
    n[op] += int(r[ix["Instructions Executed"]])
it = npart / 32
print("%-22s %10s %10s %10s %8s" % ("op", "inst/32p", "wf/32p", "ideal/32p", "wf/inst"))
for op, v in wf.most_common():
    print("%-22s %10.1f %10.1f %10.1f %8.2f" % (op, n[op] / it, v / it, ideal[op] / it, v / max(n[op], 1)))
print("%-22s %10.1f %10.1f %10.1f" % ("total", sum(n.values()) / it, sum(wf.values()) / it, sum(ideal.values()) / it))
