"""ncu workload: two training steps (BASELINE.json configs[4] shape) on synthetic encoder features.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv python tools/ncu_train_target.py"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from ball_action_spotting_b200 import FrozenEncoderTrainer, MultiDimStacker  # noqa: E402

b, frames = (int(sys.argv[1]) if len(sys.argv) > 1 else 4), 33
net = MultiDimStacker("tf_efficientnetv2_b0.in1k", 2, num_frames=frames, stack_size=3, num_3d_blocks=4, expansion_3d_ratio=3,
                      se_reduce_3d_ratio=24, drop_rate=0.2, drop_path_rate=0.2).init_random_(0).to("cuda:0").eval()
tr = FrozenEncoderTrainer(net, lr=1e-3)
g = torch.Generator().manual_seed(0)
feats = (torch.randn((b, frames // 3, 23, 40, 192), generator=g) * 0.7).half().to("cuda:0")
targets = (torch.rand((b, 2), generator=g) > 0.7).float().to("cuda:0")
for _ in range(2):
    loss, _ = tr.step_on_features(feats, targets)
torch.cuda.synchronize()
print("loss", loss.item())
