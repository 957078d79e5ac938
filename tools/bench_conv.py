#!/usr/bin/env python
"""Stand-alone timing of the dense 3x3 blocks (mds_k_conv3x3) at their real resolutions: us per launch, no other kernel running.
    python tools/bench_conv.py [n_images] [conv_mode]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from ball_action_spotting_b200 import _lib  # noqa: E402

import ctypes, os
if os.environ.get("MDS_TRACE_LIB"):          # instrumented build (tools/conv_trace.sh): timeline on stderr
    lib = ctypes.CDLL(os.environ["MDS_TRACE_LIB"])
    for name, (res, args) in _lib.SIGNATURES.items():
        fn = getattr(lib, name); fn.restype, fn.argtypes = res, args
else:
    lib = _lib.load()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lib.mds_set_conv_mode(mode)
DEV = "cuda:0"
SHAPES = [("b0.0", 32, 16, 1, 0, 0, 368, 640), ("b1.0", 16, 64, 2, 32, 0, 368, 640), ("b1.1", 32, 128, 1, 32, 1, 184, 320),
          ("b2.0", 32, 128, 2, 48, 0, 184, 320), ("b2.1", 48, 192, 1, 48, 1, 92, 160)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
for name, cin, cmid, stride, cproj, res, H, W in SHAPES:
    x = torch.randn(n, H, W, cin, device=DEV).half()
    w1 = (torch.randn(cmid, 9 * cin, device=DEV) * 0.05).half()
    b1 = torch.randn(cmid, device=DEV) * 0.1
    w2 = (torch.randn(max(cproj, 16), cmid, device=DEV) * 0.05).half()
    b2 = torch.randn(max(cproj, 16), device=DEV) * 0.1
    Ho, Wo = H // stride, W // stride
    out = torch.zeros(n, Ho, Wo, cproj or cmid, device=DEV).half()
    def run():
        rc = lib.mds_k_conv3x3(x.data_ptr(), out.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr() if cproj else None,
                               b2.data_ptr() if cproj else None, n, H, W, cin, cmid, stride, cproj, res, None)
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
    flops = 2.0 * n * Ho * Wo * (9 * cin * cmid + cmid * cproj)
    byts = 2.0 * n * (H * W * cin + Ho * Wo * (cproj or cmid))
    print(f"{name} n={n} mode={mode}: {us:7.1f} us  {flops / us * 1e-6:7.1f} TFLOP/s  {byts / us * 1e-3:7.1f} GB/s  "
          f"{us * 1.965e3 / (n * Ho * Wo / 128 / 148):7.0f} clk per 128 px per SM")
