#!/usr/bin/env python
"""SLIDING-mode throughput (BASELINE.json configs[3] on one GPU): the batched sliding-window sweep over a synthetic
clip — one encoder pass per new triple, one 3D/head pass per prediction, like the streaming predictor.
    python tools/bench_sweep.py [frames] [tta]"""
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from ball_action_spotting_b200 import MultiDimStacker  # noqa: E402
from ball_action_spotting_b200.sweep import SlidingSweep, prediction_bounds  # noqa: E402

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
tta = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
dev = torch.device("cuda:0")
net = MultiDimStacker("tf_efficientnetv2_b0.in1k", 2, num_3d_blocks=4, expansion_3d_ratio=3).init_random_(1).to(dev).eval()
frames = torch.randint(0, 256, (n_frames, 720, 1280), dtype=torch.uint8, device=dev)
sweep = SlidingSweep(net, 15, 2, (1280, 736), tta=tta, max_stacks=128)
lo, hi = prediction_bounds(sweep.gen, n_frames, 1)
for _ in range(2):
    sweep.predict_range(frames, 0, lo, min(hi + 1, lo + 256))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = sweep.predict_range(frames, 0, lo, hi + 1)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(json.dumps({"metric": "SLIDING predictions/sec (15x1280x736 window, step 2)", "value": (hi - lo + 1) / (ms / 1e3), "frames": n_frames,
                  "predictions": hi - lo + 1, "ms": ms, "tta": tta, "fps_equivalent": n_frames / (ms / 1e3)}))
