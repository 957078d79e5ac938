"""Benchmark of the frozen-encoder training step (BASELINE.json configs[4]: 33 frames, T = 11, batch 4, 1280x736).

    python tools/bench_train.py [--batch 4] [--frames 33] [--steps 20] [--warmup 5] [--cpu-steps 1] [--out FILE]

Reports, as one JSON object:
* ``step_3d``: optimizer steps/s and frame-stacks/s of mds_train_step alone (features resident in HBM) -- forward in
  train mode, backward, SGD-Nesterov + GradScaler;
* ``step_full``: the same through ``FrozenEncoderTrainer.train_step`` (uint8 frames -> frozen encoder -> step ->
  loss read back with .item(), as the reference's train loop does, src/argus_models.py:63);
* ``by_kind``: share of the step per kernel class from the library's CUDA-event hooks, with the achieved HBM GB/s of
  the BatchNorm column kernels and the TFLOP/s of the GEMMs (algorithmic bytes / flops stated in DESIGN.md §10);
* ``cpu_baseline``: the oracle (reference modules restated, torch CPU fp32, all host threads) on the same step.
The timed CPU baseline is bench.py's `train_step_cpu_baseline` (oracle port).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

KINDS = {2: "gemm_fwd_dgrad", 7: "gemm_wgrad", 8: "bn_columns", 9: "dw3d", 10: "small"}


def accounting(b, T, P, nb=4, c3=192, mid=576, pj=256):
    """Algorithmic fp16 bytes (each kernel reads its inputs once, writes its outputs once) and flops of one step."""
    M = b * T * P
    gemm_flops = 2 * M * (c3 * 192 + nb * 2 * c3 * mid + c3 * pj)          # forward
    fl = {"gemm_fwd_dgrad": gemm_flops + (gemm_flops - 2 * M * c3 * 192),  # + data gradients (none into the encoder)
          "gemm_wgrad": gemm_flops}
    # column kernels: per BN layer fwd = stats (read C) + apply (read C, write C); bwd = reduce (read 2C) + apply (read 2C, write C)
    per_bn = lambda C, extra=0: 2 * M * C * (1 + 2 + 2 + 3 + extra)
    by = per_bn(c3) + per_bn(pj) + nb * (per_bn(mid) + per_bn(mid, extra=1 + 2) + per_bn(c3, extra=1))
    return fl, by


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--frames", type=int, default=33)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--cpu-steps", type=int, default=1)
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    from ball_action_spotting_b200 import FrozenEncoderTrainer, MultiDimStacker
    dev = torch.device("cuda:0")
    H, W, SH = 736, 1280, 720
    b, T = args.batch, args.frames // 3
    net = MultiDimStacker("tf_efficientnetv2_b0.in1k", 2, num_frames=args.frames, stack_size=3, num_3d_blocks=4,
                          expansion_3d_ratio=3, se_reduce_3d_ratio=24, drop_rate=0.2, drop_path_rate=0.2).init_random_(0)
    net.to(dev).eval()
    tr = FrozenEncoderTrainer(net, lr=1e-3 * b / 4)
    g = torch.Generator(device="cpu").manual_seed(0)
    frames = [torch.randint(0, 256, (b, args.frames, SH, W), dtype=torch.uint8, generator=g).to(dev) for _ in range(2)]
    targets = (torch.rand((b, 2), generator=g) > 0.7).float().to(dev)
    feats = [tr.encoder_features(f) for f in frames]
    fh, fw = feats[0].shape[2], feats[0].shape[3]

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    ms_3d = timed(lambda i: tr.step_on_features(feats[i % 2], targets), args.steps, args.warmup)
    ms_full = timed(lambda i: tr.train_step((frames[i % 2], targets)), args.steps, args.warmup)
    ms_enc = timed(lambda i: tr.encoder_features(frames[i % 2]), args.steps, args.warmup)

    eng = net.engine(dev)
    tr.step_on_features(feats[0], targets)
    torch.cuda.synchronize()
    eng.profile_begin()
    nprof = 5
    for i in range(nprof):
        tr.step_on_features(feats[i % 2], targets)
    recs = eng.profile_end()
    by_kind = {}
    for kind, tag, ms in recs:
        d = by_kind.setdefault(KINDS.get(kind, str(kind)), {"ms_per_step": 0.0, "launches_per_step": 0})
        d["ms_per_step"] += ms / nprof
        d["launches_per_step"] += 1 / nprof
    fl, col_bytes = accounting(b, T, fh * fw)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    for k, d in by_kind.items():
        if k in fl:
            d["tflops"] = fl[k] / (d["ms_per_step"] * 1e-3) / 1e12
        if k == "bn_columns":
            d["algorithmic_gbs"] = col_bytes / (d["ms_per_step"] * 1e-3) / 1e9
            if peaks.get("hbm_gbs"):
                d["frac_of_hbm_peak"] = d["algorithmic_gbs"] / peaks["hbm_gbs"]
    total_kernel_ms = sum(d["ms_per_step"] for d in by_kind.values())
    for d in by_kind.values():
        d["share"] = d["ms_per_step"] / total_kernel_ms

    out = {"workload": f"frozen-encoder training step, batch {b}, {args.frames} frames (T={T}), 1280x736, fp16 storage / fp32 accumulate + master weights",
           "step_3d": {"ms_per_step": ms_3d, "steps_per_s": 1e3 / ms_3d, "frame_stacks_per_s": 1e3 * b / ms_3d},
           "step_full": {"ms_per_step": ms_full, "steps_per_s": 1e3 / ms_full, "frame_stacks_per_s": 1e3 * b / ms_full,
                         "includes": "uint8 frames -> frozen encoder (eval) -> train step -> loss.item()"},
           "encoder_only_ms": ms_enc, "sum_of_kernel_ms": total_kernel_ms, "by_kind": by_kind,
           "gpu_launches_per_step": sum(d["launches_per_step"] for d in by_kind.values()),
           "scaler": tr.scaler_state()}

    if args.cpu_steps > 0:
        import bench                                   # the CPU-baseline leg (the only code that executes oracle/) lives in bench.py
        sd = {k: v.detach().float().cpu() for k, v in net.state_dict().items()}
        out["cpu_baseline"] = bench.train_step_cpu_baseline(sd, args.frames, b, (fh, fw), args.cpu_steps)
        out["speedup_step_3d_vs_cpu"] = out["cpu_baseline"]["ms_per_step"] / ms_3d
    s = json.dumps(out, indent=1)
    print(s)
    if args.out:
        Path(args.out).parent.mkdir(parents=True, exist_ok=True)
        Path(args.out).write_text(s)


if __name__ == "__main__":
    main()
