"""GPU: the end-to-end error over 16 random full-size stacks, and the staged tensors at full size.

north_star: logits within 1e-3 relative (max|got - ref| / max|ref|) of the fp32 CPU forward, asserted here UNSCALED on every
one of the 16 stacks, on the logits and on the sigmoid probabilities.  The engine stores activations and GEMM weights in fp16
(BASELINE.json configs[1]); a CPU emulation of exactly those roundings on the fp32 oracle (tests/precision_study.py, DESIGN.md
"Numerics") predicts an RMS logit error of ~4e-4 of max|logit| on these random-weight networks, so the 1e-3 bound sits at about
2.4 sigma: this build measures RMS 4.2e-4, max 9.7e-4 (16 of 16 stacks within 1e-3); the forward is bit-reproducible, so the
outcome is a fixed property of the build, and a kernel change that moves any stack past 1e-3 fails this test.  The per-stack
values are written to gpurun_out/parity_report.jsonl ("sweep16.*").
forward_2d / forward_3d have no pooling after them; their full-size errors are recorded and held to 1e-3 RMS / 1e-2 max
(the emulation's prediction for un-pooled tensors of a random-weight network)."""
import json
from pathlib import Path

import pytest
import torch

from oracle import mds_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
REPORT = Path(__file__).resolve().parents[1] / "gpurun_out" / "parity_report.jsonl"


def record(name, value, bound):
    try:
        REPORT.parent.mkdir(exist_ok=True)
        with REPORT.open("a") as f:
            f.write(json.dumps({"test": name, "err": value, "tol": bound}) + "\n")
    except OSError:
        pass
    return value


def make_net(cfg, sd, **kw):
    from ball_action_spotting_b200 import MultiDimStacker
    net = MultiDimStacker("tf_efficientnetv2_b0.in1k", cfg.num_classes, num_frames=cfg.num_frames, stack_size=3,
                          num_3d_blocks=cfg.num_3d_blocks, expansion_3d_ratio=cfg.expansion_3d_ratio,
                          se_reduce_3d_ratio=cfg.se_reduce_3d_ratio, **kw)
    net.load_state_dict(sd, strict=True)
    return net.to(DEV).eval()


def test_sixteen_random_full_size_stacks(oracle_sd):
    cfg = O.ModelConfig()
    n = 16
    u8 = torch.randint(0, 256, (n, 15, 720, 1280), dtype=torch.uint8, generator=torch.Generator().manual_seed(100))
    with torch.no_grad():
        ref = torch.cat([O.forward(oracle_sd, O.pad_normalize(u8[i:i + 1], (1280, 736)), cfg) for i in range(n)])
    net = make_net(cfg, oracle_sd)
    got = torch.cat([net(u8[i:i + 4].to(DEV)).cpu() for i in range(0, n, 4)])          # batch 4 = BASELINE.json configs[1]
    err = (got - ref).abs() / ref.abs().max()
    per_stack = err.max(1).values
    perr = (torch.sigmoid(got) - torch.sigmoid(ref)).abs()
    rms, mx, within = err.pow(2).mean().sqrt().item(), err.max().item(), int((per_stack <= 1e-3).sum())
    record("sweep16.logits_rms", rms, 6e-4)
    record("sweep16.logits_max", mx, 1e-3)
    record("sweep16.stacks_within_1e-3_of_16", within, 16)
    record("sweep16.probs_max", perr.max().item(), 1e-3)
    record("sweep16.per_stack_max", [round(v, 6) for v in per_stack.tolist()], 1e-3)
    assert mx <= 1e-3 and perr.max().item() <= 1e-3 and rms <= 6e-4, (rms, mx, within, per_stack.tolist())
    # the same stacks one by one and in one batch of 16: every kernel is batch-invariant, so the logits are bit-identical
    one = torch.cat([net(u8[i:i + 1].to(DEV)).cpu() for i in range(4)])
    assert torch.equal(one, got[:4])
    assert torch.equal(net(u8.to(DEV)).cpu(), got)


def test_staged_tensors_full_size(oracle_sd):
    """forward_2d / forward_3d / forward_head at 1280 x 736 (SURVEY.md 8(d) config 1 names them)."""
    cfg = O.ModelConfig()
    x = torch.rand((1, 15, 736, 1280), generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        f2r = O.forward_2d(oracle_sd, x, cfg)
        f3r = O.forward_3d(oracle_sd, f2r, cfg)
        ref = O.forward_head(oracle_sd, f3r)
    net = make_net(cfg, oracle_sd)
    rel = lambda a, b: ((a.float().cpu() - b).abs().max() / b.abs().max()).item()
    rms = lambda a, b: ((a.float().cpu() - b).pow(2).mean().sqrt() / b.abs().max()).item()
    f2 = net.forward_2d(x.to(DEV))
    f3 = net.forward_3d(f2)
    record("full.forward_2d.max", rel(f2, f2r), 1e-2); record("full.forward_2d.rms", rms(f2, f2r), 1e-3)
    record("full.forward_3d.max", rel(f3, f3r), 1e-2); record("full.forward_3d.rms", rms(f3, f3r), 1e-3)
    assert rel(f2, f2r) <= 1e-2 and rel(f3, f3r) <= 1e-2
    assert rms(f2, f2r) <= 1e-3 and rms(f3, f3r) <= 1e-3             # RMS over the tensor: within 1e-3 of max|ref|
    # each stage on the ORACLE's input (no accumulated drift): 1e-3 on the pooled output, 1e-2 un-pooled
    assert record("full.forward_3d_from_oracle_f2.max", rel(net.forward_3d(f2r.to(DEV)), f3r), 1e-2) <= 1e-2
    assert record("full.head_from_oracle_f3", rel(net.forward_head(f3r.to(DEV)), ref), 1e-3) <= 1e-3
    assert record("full.logits", rel(net(x.to(DEV)), ref), 1e-3) <= 1e-3


def test_raw_recipe_stress_case_is_recorded():
    """SURVEY.md A.5 recipe without BN calibration ("raw"): activations reach ~850 and SE pre-activations ~50, fp16 WEIGHTS alone
    move the logits by ~6e-2 (DESIGN.md "Numerics").  Recorded, not required: the engine must stay finite and within 0.25."""
    cfg = O.ModelConfig()
    sd = O.make_state_dict(cfg, seed=1234, recipe="raw")
    x = torch.rand((1, 15, 96, 160), generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        ref = O.forward(sd, x, cfg)
    got = make_net(cfg, sd)(x.to(DEV)).cpu()
    assert torch.isfinite(got).all()
    e = record("raw_recipe.small.logits", ((got - ref).abs().max() / ref.abs().max()).item(), 0.25)
    assert e <= 0.25
