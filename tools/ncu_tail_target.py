#!/usr/bin/env python
"""Workload for ncu captures of the fused MBConv tail kernel: python tools/ncu_tail_target.py [n] [H W C rd N kt T stride]
Launches mds_k_mbconv_tail on random data: once depthwise + SE only (N = 0) and once complete."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from ball_action_spotting_b200 import _lib  # noqa: E402
from ball_action_spotting_b200.packer import bias_matrix  # noqa: E402

a = [int(v) for v in sys.argv[1:]]
n = a[0] if a else 20
H, W, C_, rd, N, kt, T, stride = (a[1:9] if len(a) >= 9 else (46, 80, 672, 28, 112, 1, 1, 1))
lib = _lib.load()
dev = "cuda:0"
Ho, Wo = H // stride, W // stride
g = torch.Generator(device=dev).manual_seed(0)
m1 = torch.randn((n, T, H, W, C_), device=dev, generator=g).half()
m2 = torch.zeros((n, T, Ho, Wo, C_), dtype=torch.float16, device=dev)
dw_w = torch.randn((9 * kt, C_), device=dev, generator=g) * 0.2
dw_b = torch.randn(C_, device=dev, generator=g) * 0.1
parts = torch.zeros((n, 64, C_), device=dev)
w1 = torch.randn((rd, C_), device=dev, generator=g) * 0.05
b1 = torch.zeros(rd, device=dev)
w2t = torch.randn((rd, C_), device=dev, generator=g) * 0.05
b2 = torch.zeros(C_, device=dev)
gate = torch.zeros((n, C_), device=dev)
sync = torch.zeros((3, n), dtype=torch.int32, device=dev)
wp = (torch.randn((N, C_), device=dev, generator=g) * C_ ** -0.5).half()
bm = bias_matrix(torch.zeros(N)).to(dev)
res = torch.randn((n, T * Ho * Wo, N), device=dev, generator=g).half()
out = torch.zeros((n, T * Ho * Wo, N), dtype=torch.float16, device=dev)
for NN in (0, N, 0, N):
    rc = lib.mds_k_mbconv_tail(m1.data_ptr(), m2.data_ptr(), dw_w.data_ptr(), dw_b.data_ptr(), parts.data_ptr(), w1.data_ptr(),
                               b1.data_ptr(), w2t.data_ptr(), b2.data_ptr(), gate.data_ptr(), sync.data_ptr(), wp.data_ptr(),
                               bm.data_ptr(), res.data_ptr(), out.data_ptr(), n, T, H, W, C_, kt, stride, rd, NN, 0, None)
    assert rc == 0, lib.mds_last_error().decode()
    torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
