#!/usr/bin/env python
"""Streaming-predictor throughput: frames/s of MultiDimStackerPredictor.predict fed one uint8 frame at a
time, as scripts/ball_action/predict.py:29-55 does -- (a) with the reference script's per-frame `.cpu()` (one host sync per
frame, predict.py:48) and (b) keeping the predictions on the device and copying them once at the end.
    python tools/bench_predictor.py [frames] [tta]"""
import json
import sys
import tempfile
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from ball_action_spotting_b200 import MultiDimStacker, MultiDimStackerPredictor  # noqa: E402

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 600
tta = bool(int(sys.argv[2])) if len(sys.argv) > 2 else False
dev = torch.device("cuda:0")
kwargs = dict(model_name="tf_efficientnetv2_b0.in1k", num_classes=2, num_frames=15, stack_size=3, index_2d_features=4, pretrained=False,
              num_3d_blocks=4, num_3d_features=192, expansion_3d_ratio=3, se_reduce_3d_ratio=24, num_3d_stack_proj=256,
              drop_rate=0.2, drop_path_rate=0.2, act_layer="silu")
net = MultiDimStacker(**kwargs).init_random_(1)
params = {"nn_module": ("multidim_stacker", kwargs), "frames_processor": ("pad_normalize", {"size": (1280, 736), "pad_mode": "constant", "fill_value": 0}),
          "frame_stack_size": 15, "frame_stack_step": 2}
with tempfile.TemporaryDirectory() as d:
    path = Path(d) / "model.pth"
    torch.save({"model_name": "BallActionModel", "params": params, "nn_state_dict": net.state_dict()}, path)
    predictor = MultiDimStackerPredictor(path, device="cuda:0", tta=tta)
frames = torch.randint(0, 256, (64, 720, 1280), dtype=torch.uint8, device=dev)
out = {}
for mode in ("cpu_per_frame", "device_resident"):
    predictor.reset_buffers()
    for i in range(60):                                   # warm-up: fill the window
        predictor.predict(frames[i % 64], i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    kept = []
    for i in range(60, 60 + n_frames):
        pred, idx = predictor.predict(frames[i % 64], i)
        if pred is not None:
            kept.append(pred.cpu().numpy() if mode == "cpu_per_frame" else pred)
    if mode == "device_resident":
        torch.stack(kept).cpu()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out[mode] = {"frames_per_s": n_frames / dt, "ms_per_frame": 1e3 * dt / n_frames}
print(json.dumps({"metric": "streaming predict() frames/s (1280x720 uint8 frame in, probabilities out), steady state", "tta": tta,
                  "frames": n_frames, **out}))
