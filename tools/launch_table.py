"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list per kernel name.
    python tools/launch_table.py gpurun_out/train_launches.csv"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name, v, unit = r[4].split("(")[0], float(r[-1].replace(",", "")), r[-2]
    v = v / 1e3 if unit == "ns" else v * 1e3 if unit == "ms" else v
    d = agg.setdefault(name, [0, 0.0])
    d[0] += 1
    d[1] += v
tot = sum(d[1] for d in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:72]:72s} n={n:4d} total={t:9.1f}us avg={t / n:8.1f}us {100 * t / tot:5.1f}%")
print(f"total {tot:.1f} us over {sum(d[0] for d in agg.values())} launches")
