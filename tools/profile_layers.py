#!/usr/bin/env python
"""Per-layer CUDA-event profile of one FULL forward (uses the library's mds_profile_* hooks).
    python tools/profile_layers.py [--batch B] [--chunk-images C] [--reps R]
Prints per launch: ms, achieved GB/s (algorithmic bytes) and TFLOP/s."""
import argparse
import sys
from collections import defaultdict
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from ball_action_spotting_b200 import MultiDimStacker, accounting as acc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--chunk-images", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--tail-mode", type=int, default=3)
ap.add_argument("--conv-mode", type=int, default=2)
a = ap.parse_args()
H, W, SH, T = 736, 1280, 720, 5
dev = torch.device("cuda:0")
net = MultiDimStacker("tf_efficientnetv2_b0.in1k", 2, num_3d_blocks=4, expansion_3d_ratio=3, chunk_images=a.chunk_images).init_random_(1).to(dev).eval()
eng = net.engine(dev)
eng.lib.mds_set_tail_mode(a.tail_mode)
eng.lib.mds_set_conv_mode(a.conv_mode)
x = torch.randint(0, 256, (a.batch, 15, SH, W), dtype=torch.uint8, device=dev)
desc = eng.frames_desc(x, H, W, 3 * SH * W, SH * W)
for _ in range(2):
    eng.forward(desc, a.batch)
torch.cuda.synchronize()
eng.profile_begin()
for _ in range(a.reps):
    eng.forward(desc, a.batch)
recs = eng.profile_end()
# group by (kind, tag, ordinal within tag)
per = defaultdict(float)
order = []
seen = defaultdict(int)
n_per_rep = len(recs) // a.reps
for i, (kind, tag, ms) in enumerate(recs):
    j = i % n_per_rep
    per[j] += ms
launch = [(recs[j][0], recs[j][1]) for j in range(n_per_rep)]
# expected launch list for one chunk of the encoder + 3D
enc = acc.encoder_launches(H, W, SH, tail_mode=0 if a.tail_mode == 3 else a.tail_mode)
s3d = acc.stack3d_launches(H // 32, W // 32, T, tail_mode=0 if a.tail_mode == 3 else a.tail_mode)
chunk = eng.cfg.chunk_images or 160
n_img = a.batch * T
chunks = -(-n_img // chunk)
rows = []
j = 0
for ci in range(chunks):
    cs = min(chunk, n_img - ci * chunk)
    for l in enc:
        k, tg = launch[j]
        assert k == l.kind, (j, k, l.kind, l.name)
        rows.append((l.name, l.kind, per[j] / a.reps, l.bytes * cs, l.flops * cs)); j += 1
for l in s3d:
    reps_here = 2 if l.kind == 6 else 1
    ms = 0.0
    for _ in range(reps_here):
        ms += per[j] / a.reps; j += 1
    rows.append((l.name, l.kind, ms, l.bytes * a.batch, l.flops * a.batch))
agg = defaultdict(lambda: [0.0, 0.0, 0.0])
for name, kind, ms, by, fl in rows:
    agg[name][0] += ms; agg[name][1] += by; agg[name][2] += fl
tot = sum(v[0] for v in agg.values())
print(f"batch {a.batch} chunk {chunk}: {tot:.2f} ms per forward -> {a.batch / tot * 1e3:.0f} stacks/s (sum of kernel times)")
print(f"{'layer':16s} {'ms':>8s} {'%':>6s} {'GB/s':>8s} {'TFLOP/s':>8s}")
for name, (ms, by, fl) in agg.items():
    print(f"{name:16s} {ms:8.3f} {100 * ms / tot:6.2f} {by / ms / 1e6 if ms else 0:8.0f} {fl / ms / 1e9 if ms else 0:8.1f}")
