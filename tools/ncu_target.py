#!/usr/bin/env python
"""One FULL forward at a given batch, no warm-up: the workload ncu captures kernels from (see profiles/README.md)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from ball_action_spotting_b200 import MultiDimStacker  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
net = MultiDimStacker("tf_efficientnetv2_b0.in1k", 2, num_3d_blocks=4, expansion_3d_ratio=3, chunk_images=batch * 5).init_random_(1).to(dev).eval()
x = torch.randint(0, 256, (batch, 15, 720, 1280), dtype=torch.uint8, device=dev)
y = net(x)
torch.cuda.synchronize()
print(y[0].tolist())
