"""Parity of the CUDA training step (mds_train_step) against the CPU fp32 oracle, tensor by tensor.

Usage (GPU box): python tests/train_parity.py            # prints a table for a few shapes
Test helper (imported by tests/test_train_gpu.py); lives under tests/ because it uses the oracle as the checker.
"""
from __future__ import annotations

import sys
from pathlib import Path
from typing import Dict, Tuple

import torch

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from oracle import mds_oracle as O            # noqa: E402
from oracle import mds_train_oracle as TO     # noqa: E402

ALPHA, GAMMA = 0.4, 1.2


def build(cfg: O.ModelConfig, sd, device="cuda:0", amp=True, lr=0.01, init_scale=65536.0):
    from ball_action_spotting_b200 import FrozenEncoderTrainer, MultiDimStacker
    net = MultiDimStacker("tf_efficientnetv2_b0.in1k", cfg.num_classes, num_frames=cfg.num_frames, stack_size=3,
                          num_3d_blocks=cfg.num_3d_blocks, expansion_3d_ratio=cfg.expansion_3d_ratio,
                          se_reduce_3d_ratio=cfg.se_reduce_3d_ratio, num_3d_stack_proj=cfg.num_3d_stack_proj,
                          drop_rate=0.2, drop_path_rate=0.2)
    net.load_state_dict(sd)
    net.to(device).eval()
    return net, FrozenEncoderTrainer(net, lr=lr, focal_alpha=ALPHA, focal_gamma=GAMMA, amp=amp, init_scale=init_scale)


def to_nhwc16(enc: torch.Tensor, b: int, T: int, device="cuda:0") -> torch.Tensor:
    n, c, h, w = enc.shape
    return enc.view(b, T, c, h, w).permute(0, 1, 3, 4, 2).contiguous().to(device, torch.float16)


def rel_l2(a: torch.Tensor, r: torch.Tensor, floor: float = 0.0) -> float:
    a, r = a.double().flatten(), r.double().flatten()
    return float((a - r).norm() / max(float(r.norm()), floor))


def compare(cfg: O.ModelConfig, b: int, hw: Tuple[int, int], seed: int = 7, mask_seed: int = 11, sd=None,
            drop: bool = True) -> Dict[str, float]:
    """One forward + backward (no update) on identical inputs / masks; relative L2 error per tensor."""
    sd = sd if sd is not None else O.make_state_dict(cfg, seed=1234, calib_hw=(96, 160))
    enc, targets = TO.make_case(cfg, b, hw, seed)
    enc = enc.half().float()                                  # both sides see the same fp16-representable input
    dp, do = TO.make_masks(cfg, b, 0.2 if drop else 0.0, 0.2 if drop else 0.0, mask_seed)
    loss, logits, grads, stats = TO.loss_and_grads(sd, enc, targets, cfg, dp, do, ALPHA, GAMMA)
    net, tr = build(cfg, sd)
    g_loss, g_logits = tr.step_on_features(to_nhwc16(enc, b, cfg.num_stacks), targets, dp, do, apply_update=False)
    gmax = max(float(g.abs().max()) for g in grads.values())
    out = {"loss": abs(g_loss.item() - loss.item()) / abs(loss.item()),
           "logits": float((g_logits.cpu() - logits).abs().max() / logits.abs().max())}
    for k, g in grads.items():
        # gradients that are exactly zero in exact arithmetic (see make_train_golden.py) are compared on an absolute floor
        out["grad:" + k] = rel_l2(tr.get(k, "grad"), g, floor=1e-3 * gmax * g.numel() ** 0.5)
    for k, v in stats.items():
        if not k.endswith("num_batches_tracked"):
            out["stat:" + k] = rel_l2(tr.get(k), v)
    tr.close()
    return out


def main():
    cases = [("t3 b2 3x5", O.ModelConfig(num_frames=9), 2, (3, 5)),
             ("t11 b2 4x6", O.ModelConfig(num_frames=33), 2, (4, 6)),
             ("t5 b3 23x40", O.ModelConfig(num_frames=15), 3, (23, 40)),
             ("t11 b4 23x40", O.ModelConfig(num_frames=33), 4, (23, 40))]
    for name, cfg, b, hw in cases:
        try:
            errs = compare(cfg, b, hw)
        except Exception as e:      # keep going: one GPU call should report every case
            print(f"== {name}: FAILED {type(e).__name__}: {e}")
            continue
        worst = max(errs, key=errs.get)
        print(f"== {name}: worst {worst} = {errs[worst]:.3e}")
        for k, v in errs.items():
            flag = "  <<<" if v > 2e-2 else ""
            print(f"   {k:55s} {v:.3e}{flag}")


if __name__ == "__main__":
    main()
