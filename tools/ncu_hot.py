#!/usr/bin/env python
"""Top stall-sampled SASS instructions of an .ncu-rep (needs -lineinfo builds): python tools/ncu_hot.py rep [N]"""
import csv
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out[1:]))
hdr = rows[0]
ci, si, ei = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
data = []
for k, r in enumerate(rows[1:]):
    try:
        data.append((int(r[ci]), k, r[si].strip(), int(r[ei])))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data) or 1
print(out[0][:150], "total samples", tot)
for n, k, s, e in sorted(data, reverse=True)[:top]:
    print(f"{100 * n / tot:5.1f}%  #{k:<5d} exec {e:>9d}  {s[:110]}")
