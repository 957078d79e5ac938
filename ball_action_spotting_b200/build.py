"""Build recipe for libmds_b200.so (nvcc, sm_100a only, in-tree so the .so travels with the repo snapshot)."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libmds_b200.so"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def sources():
    return sorted(CSRC.glob("*.cu")), sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.inl")) + [PKG.parent / "include" / "mds_b200.h"])


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    cus, hdrs = sources()
    return any(p.stat().st_mtime > t for p in cus + hdrs)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cus, _ = sources()
    cmd = [nvcc, *NVCC_FLAGS, "-o", str(LIB), *map(str, cus)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libmds_b200.so")
    (PKG / "build_ptxas.log").write_text(res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
