#!/usr/bin/env python
"""Distribution of the end-to-end logit error over N full-size stacks (GPU engine vs CPU fp32 oracle).
    python tests/parity_sweep.py [N] [seed]
Lives under tests/ because it uses the oracle as the checker."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from ball_action_spotting_b200 import MultiDimStacker  # noqa: E402
from oracle import mds_oracle as O  # noqa: E402  (checker)

N = int(sys.argv[1]) if len(sys.argv) > 1 else 12
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 100
cfg = O.ModelConfig()
sd = O.make_state_dict(cfg, seed=1234)
u8 = torch.randint(0, 256, (N, 15, 720, 1280), dtype=torch.uint8, generator=torch.Generator().manual_seed(seed))
torch.set_num_threads(torch.get_num_threads())
with torch.no_grad():
    ref = torch.cat([O.forward(sd, O.pad_normalize(u8[i:i + 1], (1280, 736)), cfg) for i in range(N)])
res = {}
for bc in (False, True):
    net = MultiDimStacker("tf_efficientnetv2_b0.in1k", 2, num_3d_blocks=4, expansion_3d_ratio=3, bias_correction=bc)
    net.load_state_dict(sd)
    net.to("cuda:0").eval()
    got = net(u8.to("cuda:0")).cpu()
    err = (got - ref).abs() / ref.abs().max()
    perr = (torch.sigmoid(got) - torch.sigmoid(ref)).abs()
    res[bc] = {"max": err.max().item(), "rms": err.pow(2).mean().sqrt().item(), "per_stack_max": [round(v, 6) for v in err.max(1).values.tolist()],
               "prob_max": perr.max().item()}
    print("bias_correction", bc, json.dumps(res[bc]))
print("ref absmax", ref.abs().max().item())
