#!/usr/bin/env python
"""Stand-alone timing of the 1x1-conv GEMMs (tcgen05 kernel) at their real shapes: us per launch, nothing else running.
    python tools/bench_gemm.py [n_images]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ball_action_spotting_b200 import _lib  # noqa: E402
from ball_action_spotting_b200.packer import bias_matrix  # noqa: E402

import ctypes, os
if os.environ.get("MDS_TRACE_LIB"):          # instrumented build (tools/conv_trace.sh): timeline on stderr
    lib = ctypes.CDLL(os.environ["MDS_TRACE_LIB"])
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(lib, name); fn.restype, fn.argtypes = res, args
else:
    lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
DEV = "cuda:0"
# (name, rows per image, N, K, gated(pre-gated weights), res, act)
SHAPES = [("b3.0.pw", 184 * 320, 192, 48, 0, 0, 1), ("b3.1.pw", 92 * 160 // 4 * 4 // 4, 384, 96, 0, 0, 1),
          ("b4.1.pw", 46 * 80, 672, 112, 0, 0, 1), ("b4.1.pw-noact", 46 * 80, 672, 112, 0, 0, 0), ("b4.1.pwl", 46 * 80, 112, 672, 1, 1, 0),
          ("b5.1.pw", 23 * 40, 1152, 192, 0, 0, 1), ("b5.1.pw-noact", 23 * 40, 1152, 192, 0, 0, 0), ("b5.1.pwl", 23 * 40, 192, 1152, 1, 1, 0),
          ("b3.1.pwl", 46 * 80, 96, 384, 1, 1, 0)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
for name, rows, N, K, gated, res, act in SHAPES:
    M = rows * n
    A = torch.randn(M, K, device=DEV).half()
    W = (torch.randn(n if gated else 1, N, K, device=DEV) * K ** -0.5).half()
    bias = torch.randn(N) * 0.1
    bm = bias_matrix(bias).to(DEV)
    r = torch.randn(M, N, device=DEV).half() if res else None
    out = torch.zeros(M, N, device=DEV).half()
    def run():
        if gated:
            rc = lib.mds_k_gemm_gated(A.data_ptr(), W.data_ptr(), bm.data_ptr(), r.data_ptr() if res else None, out.data_ptr(), rows, n, N, K, act, None)
        else:
            rc = lib.mds_k_gemm1x1(A.data_ptr(), W.data_ptr(), bias.to(DEV).data_ptr(), bm.data_ptr(), None, None, out.data_ptr(), rows, n, N, K, act, None)
        assert rc == 0, lib.mds_last_error()
    for _ in range(3):
        run()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    us = sorted(ts)[len(ts) // 2]
    byts = 2.0 * M * (K + N * (2 if res else 1))
    print(f"{name:14s} n={n}: {us:7.1f} us  {byts / us * 1e-3:7.1f} GB/s  {2.0 * M * N * K / us * 1e-6:6.1f} TFLOP/s  "
          f"{us * 1.965e3 / (M / 128 / 148):7.0f} clk per 128-row tile per SM")
