#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libmds_b200.so (cuobjdump -sass): python tools/sass_histogram.py > profiles/sass_<round>.txt
The mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,
UTMALDG = TMA tensor load, HMMA = mma.sync (legacy tensor path), LDGSTS = cp.async, FFMA2 = packed fp32."""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
lib = ROOT / "ball_action_spotting_b200" / "libmds_b200.so"
out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
KEY = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "LDSM", "LDGSTS", "FFMA2", "FMUL2", "FADD2",
       "MUFU", "USETMAXREG", "UCGABAR_ARV", "ATOM", "RED", "MEMBAR", "STL", "LDL"]
name, counts, kernels = None, Counter(), []
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        if name:
            kernels.append((name, counts))
        name, counts = m.group(1), Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m and name:
        counts[m.group(1)] += 1
if name:
    kernels.append((name, counts))
demangle = subprocess.run(["c++filt"] + [k for k, _ in kernels], capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {lib.name}: opcode counts per kernel (static instruction counts)")
print("# tcgen05.mma = UTC*MMA, tcgen05.ld = LDTM, TMA = UTMALDG, mma.sync = HMMA, cp.async = LDGSTS, packed fp32 = FFMA2/FMUL2/FADD2\n")
for (mangled, c), dn in sorted(zip(kernels, demangle), key=lambda t: t[1]):
    short = re.sub(r"\(.*", "", dn).replace("void ", "").replace("mds::", "")
    total = sum(c.values())
    keys = ", ".join(f"{k} {c[k]}" for k in KEY if c[k])
    top = ", ".join(f"{k} {v}" for k, v in c.most_common(6))
    print(f"{short}\n    total {total}; key: {keys or '-'}\n    top: {top}")
