#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel from an `ncu --page source --csv` dump.  usage: ncu_hot.py source.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[h]
ix = {k: i for i, k in enumerate(hdr)}
data = []
for r in rows[h + 1:]:
    if len(r) != len(hdr) or not r[0].startswith("0x"):
        break
    data.append(r)
tot = sum(int(r[ix["# Samples"]]) for r in data)
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:N]
for i in sorted(order):
    r = data[i]
    top = sorted(((int(r[ix[k]]), k[6:]) for k in stalls), reverse=True)[:2]
    print("%5d %5.1f%% x%-9s %-70s %s" % (i, 100 * int(r[ix["# Samples"]]) / tot, r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:70],
                                     " ".join("%s=%d" % (k, v) for v, k in top if v)))
