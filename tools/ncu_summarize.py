#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep + launches csv into committed text summaries under profiles/.
    python tools/ncu_summarize.py r01 gpurun_out/prof_*.ncu-rep [--launches gpurun_out/launches_b4.csv] [--workload TEXT]"""
import csv
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
           "smsp__inst_executed.sum"]


def raw(rep):
    """-> header, units, [one row of values per captured launch]; accepts an .ncu-rep or its `--page raw --csv` export."""
    if rep.endswith(".csv"):
        out = Path(rep).read_text()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = [r for r in csv.reader(out.splitlines()) if r]
    return rows[0], rows[1], rows[2:]


def main():
    tag, args = sys.argv[1], sys.argv[2:]
    launches = None
    if "--launches" in args:
        i = args.index("--launches")
        launches = args[i + 1]
        args = args[:i] + args[i + 2:]
    workload = "tools/ncu_target.py 4 (FULL forward, batch 4 = 20 images, BASELINE.json configs[1])"
    if "--workload" in args:
        i = args.index("--workload")
        workload = args[i + 1]
        args = args[:i] + args[i + 2:]
    out_dir = ROOT / "profiles"
    out_dir.mkdir(exist_ok=True)
    lines = [f"# ncu --set full --clock-control none, one launch per kernel (round {tag}); workload: {workload}", ""]
    traffic = {}
    for rep in args:
        hdr, units, launches_ = raw(rep)
        for li, vals in enumerate(launches_):
            name = vals[hdr.index("Kernel Name")]
            lines.append(f"## {Path(rep).name.replace('.raw.csv', '').replace('.ncu-rep', '')} launch {li}: {name}")
            d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
            for m in METRICS:
                if m in d:
                    lines.append(f"  {m:70s} {d[m][0]:>16s} {d[m][1]}")
            try:
                t = float(d["gpu__time_duration.sum"][0].replace(",", ""))
                tu = d["gpu__time_duration.sum"][1]
                us = t * {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(tu.replace("second", "s").replace("usecond", "us").replace("msecond", "ms").replace("nsecond", "ns"), 1)
                def b(x):
                    v, u = d[x]
                    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
                tr = b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
                lines.append(f"  {'traffic = dram read + write':70s} {tr / 1e6:16.1f} MB  -> {tr / us / 1e3:.0f} GB/s over the launch")
                traffic[f"{Path(rep).name.split('.')[0]}[{li}]"] = {"kernel": name.split("(")[0], "traffic_bytes": tr, "duration_us": us}
            except Exception as e:  # noqa: BLE001
                lines.append(f"  (traffic n/a: {e})")
            stall = sorted(((float(vals[i].replace(",", "")), h) for i, h in enumerate(hdr)
                            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h and vals[i].replace(",", "").replace(".", "").isdigit()), reverse=True)
            tot = sum(v for v, _ in stall) or 1
            lines.append("  top stall reasons: " + ", ".join(f"{h.split('stalled_')[1]} {100 * v / tot:.0f}%" for v, h in stall[:5]))
            lines.append("")
    import json
    (out_dir / f"ncu_{tag}_traffic_by_capture.json").write_text(json.dumps(traffic, indent=1))
    (out_dir / f"ncu_{tag}_summary.txt").write_text("\n".join(lines))
    print("\n".join(lines))
    if launches:
        rows = list(csv.reader(l for l in open(launches) if l.startswith('"')))
        hdr = rows[0]
        ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
        agg = defaultdict(lambda: [0, 0.0])
        for r in rows[1:]:
            try:
                agg[r[ki].split("(")[0]][0] += 1
                agg[r[ki].split("(")[0]][1] += float(r[vi].replace(",", ""))
            except (ValueError, IndexError):
                pass
        tot = sum(v[1] for v in agg.values()) or 1
        unit = rows[1][hdr.index("Metric Unit")]
        txt = [f"# ncu --metrics gpu__time_duration.sum --clock-control none on `python bench.py --steps 2 --warmup 3` (batch 4); "
               f"cold-cache serialised launches: compare SHARES. {len(rows) - 1} launches, unit {unit}", "",
               f"{'kernel':60s} {'launches':>8s} {'total':>14s} {'share':>7s}"]
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            txt.append(f"{k[:60]:60s} {n:8d} {t:14.1f} {100 * t / tot:6.1f}%")
        (out_dir / f"ncu_{tag}_launches.txt").write_text("\n".join(txt))
        print("\n".join(txt))


if __name__ == "__main__":
    main()
