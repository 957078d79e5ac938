#!/usr/bin/env python
"""Top warp-stall sites of one captured launch: python tools/ncu_stalls.py REP LAUNCH_INDEX [min_pct]"""
import csv, subprocess, sys
rep, li = sys.argv[1], int(sys.argv[2])
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.8
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(li), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:150])
hdr, data = rows[1], rows[2:]
seen, d2 = set(), []
for r in data:
    if r[0] in seen:
        continue
    seen.add(r[0]); d2.append(r)
data = d2
ia, isamp, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
tot = sum(int(r[isamp]) for r in data if r[isamp].isdigit())
print("total samples", tot, "instructions", len(data))
names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
for idx, r in enumerate(data):
    if not r[isamp].isdigit():
        continue
    for n in names:
        v = r[hdr.index(n)]
        if v.isdigit():
            agg[n] = agg.get(n, 0) + int(v)
    if int(r[isamp]) < tot * minp / 100:
        continue
    st = {n: int(r[hdr.index(n)]) for n in names if r[hdr.index(n)].isdigit() and int(r[hdr.index(n)]) > 0}
    st = sorted(st.items(), key=lambda x: -x[1])[:2]
    print(f"{idx:5d} {int(r[isamp]):6d} {100*int(r[isamp])/tot:5.1f}% ex={r[iex]:>8s} {r[ia].strip()[:64]:64s} {st}")
print({k: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]})
