"""Pin the oracle and write tests/golden/*.npz.  Run in the AUTHORING container (needs /root/reference):

    python oracle/make_golden.py

Steps:
 1. import the reference's ``src/models/multidim_stacker.py`` UNMODIFIED by file path, with ``oracle/timm_shim``
    standing in for the absent ``timm==0.9.2``;
 2. build ``MultiDimStacker(**nn_module_params)`` for the 15-frame and the 33-frame configs, load the oracle's
    seeded state dict with ``strict=True`` (pins key names and shapes against the reference + timm naming);
 3. check the functional oracle (``oracle/mds_oracle.py``) against the reference module: forward_2d, forward_3d,
    forward_head, forward — bit-exact on the same CPU;
 4. cross-check the encoder against an independent torchvision ``EfficientNet`` assembly with explicit TF-SAME
    (0,1,0,1) pre-padding on the five stride-2 convs (parameter count 5 610 384 and forward equality);
 5. check frames.py / indexes.py restatements against the reference's own functions (imported by path);
 6. write small fixtures (logits, per-tap checksums, first values) for the committed golden tests.
"""
from __future__ import annotations

import hashlib
import importlib.util
import json
import sys
from functools import partial
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))

from oracle import mds_oracle as O  # noqa: E402


def load_by_path(name: str, path: Path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_model_module():
    shim = str(ROOT / "oracle" / "timm_shim")
    if shim not in sys.path:
        sys.path.insert(0, shim)
    return load_by_path("ref_multidim_stacker", REF / "src/models/multidim_stacker.py")


def nn_module_params(cfg: O.ModelConfig) -> dict:
    """configs/ball_action/sampling_weights_001.py:30-45 (pretrained forced False, train.py:48-49)."""
    return dict(model_name="tf_efficientnetv2_b0.in1k", num_classes=cfg.num_classes, num_frames=cfg.num_frames,
                stack_size=cfg.stack_size, index_2d_features=4, pretrained=False, num_3d_blocks=cfg.num_3d_blocks,
                num_3d_features=cfg.num_3d_features, expansion_3d_ratio=cfg.expansion_3d_ratio,
                se_reduce_3d_ratio=cfg.se_reduce_3d_ratio, num_3d_stack_proj=cfg.num_3d_stack_proj,
                drop_rate=0.2, drop_path_rate=0.2, act_layer="silu")


def build_reference(cfg: O.ModelConfig, sd):
    ref = load_reference_model_module()
    m = ref.MultiDimStacker(**nn_module_params(cfg))
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return m.eval()


def torchvision_encoder(sd):
    """Independent assembly: torchvision EfficientNet blocks + explicit SAME padding on stride-2 convs."""
    from torchvision.models.efficientnet import EfficientNet, FusedMBConvConfig, MBConvConfig
    from torch import nn
    setting = [FusedMBConvConfig(1, 3, 1, 32, 16, 1), FusedMBConvConfig(4, 3, 2, 16, 32, 2),
               FusedMBConvConfig(4, 3, 2, 32, 48, 2), MBConvConfig(4, 3, 2, 48, 96, 3),
               MBConvConfig(6, 3, 1, 96, 112, 5), MBConvConfig(6, 3, 2, 112, 192, 8)]
    net = EfficientNet(setting, dropout=0.0, norm_layer=partial(nn.BatchNorm2d, eps=1e-3), last_channel=1280)
    n_full = sum(p.numel() for p in net.parameters())
    feats = net.features[:7]                       # stem + 6 stages, drops the 1280-wide head conv

    class SamePad(nn.Module):
        def __init__(self, conv):
            super().__init__()
            self.conv = conv
            conv.padding = (0, 0)

        def forward(self, x):
            return self.conv(torch.nn.functional.pad(x, (0, 1, 0, 1)))

    def patch(mod):
        for name, child in list(mod.named_children()):
            if isinstance(child, nn.Conv2d) and child.stride == (2, 2):
                setattr(mod, name, SamePad(child))
            else:
                patch(child)
    patch(feats)
    enc = {k[len("conv2d_encoder."):]: v for k, v in sd.items() if k.startswith("conv2d_encoder.")}
    tv = feats.state_dict()
    assert len(tv) == len(enc), (len(tv), len(enc))
    mapped = {}
    for (tk, tvv), (ok, ov) in zip(tv.items(), enc.items()):
        assert tvv.shape == ov.shape, (tk, ok, tvv.shape, ov.shape)
        mapped[tk] = ov
    feats.load_state_dict(mapped, strict=True)
    n_feat = sum(p.numel() for p in feats.parameters())
    return feats.eval(), n_feat, n_full


def digest(t: torch.Tensor) -> str:
    return hashlib.sha256(t.detach().contiguous().numpy().tobytes()).hexdigest()


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    report = {}
    out = ROOT / "tests" / "golden"
    out.mkdir(parents=True, exist_ok=True)

    # ---- frames.py / indexes.py --------------------------------------------------------------------
    frames_ref = load_by_path("ref_frames", REF / "src/frames.py")
    indexes_ref = load_by_path("ref_indexes", REF / "src/indexes.py")
    g = torch.Generator().manual_seed(7)
    u8 = torch.randint(0, 256, (2, 3, 720, 1280), dtype=torch.uint8, generator=g)
    proc = frames_ref.get_frames_processor("pad_normalize", dict(size=(1280, 736), pad_mode="constant", fill_value=0))
    assert torch.equal(proc(u8), O.pad_normalize(u8, (1280, 736)))
    for size, step in [(15, 2), (33, 2), (15, 1), (4, 3)]:
        gen = indexes_ref.StackIndexesGenerator(size, step)
        for i in (-3, 0, 17, 100):
            assert gen.make_stack_indexes(i) == O.make_stack_indexes(i, size, step)
            for sz in (0, 1):
                assert gen.clip_index(i, 200, sz) == O.clip_index(i, 200, size, step, sz)
    report["frames_indexes"] = "equal"

    golden = {}
    for tag, cfg, hw in [("t5_small", O.ModelConfig(), (96, 160)), ("t5_full", O.ModelConfig(), (736, 1280)),
                         ("t11_small", O.ModelConfig(num_frames=33), (96, 160))]:
        sd = O.make_state_dict(cfg, seed=1234)
        ref = build_reference(cfg, sd)
        gx = torch.Generator().manual_seed(0)
        x = torch.rand((1, cfg.num_frames, *hw), generator=gx)
        taps = {}

        def tap(name, t):
            taps[name] = t
        with torch.no_grad():
            f2_ref = ref.forward_2d(x)
            f3_ref = ref.forward_3d(f2_ref)
            lg_ref = ref.forward_head(f3_ref)
            lg_ref2 = ref(x)
            f2 = O.forward_2d(sd, x, cfg, tap)
            f3 = O.forward_3d(sd, f2, cfg, tap)
            lg = O.forward_head(sd, f3, tap)
        for a, b, n in [(f2, f2_ref, "forward_2d"), (f3, f3_ref, "forward_3d"), (lg, lg_ref, "logits"),
                        (lg, lg_ref2, "forward")]:
            err = (a - b).abs().max().item()
            report[f"{tag}.{n}.max_abs_diff_vs_reference"] = err
            assert err == 0.0, (tag, n, err)
        n_params = sum(p.numel() for p in ref.conv2d_encoder.parameters())
        report[f"{tag}.encoder_params"] = n_params
        assert n_params == 5_610_384

        if tag == "t5_small":
            tv, n_feat, n_full = torchvision_encoder(sd)
            assert n_feat == 5_610_384, n_feat
            report["torchvision.features_params"] = n_feat
            report["torchvision.full_params_1000cls"] = n_full
            with torch.no_grad():
                xi = x.reshape(cfg.num_stacks, 3, *hw)
                a = tv(xi)
                b = O.encoder_forward(sd, xi)
            err = (a - b).abs().max().item() / b.abs().max().item()
            report["torchvision.encoder_rel_diff"] = err
            assert err < 1e-5, err

        entry = {"logits": lg.numpy(), "forward_2d_head": f2.flatten()[:64].numpy(),
                 "forward_3d_head": f3.flatten()[:64].numpy(),
                 "forward_2d_absmax": np.float32(f2.abs().max().item()),
                 "forward_3d_absmax": np.float32(f3.abs().max().item()),
                 "forward_2d_mean": np.float64(f2.double().mean().item()),
                 "forward_3d_mean": np.float64(f3.double().mean().item())}
        for k, v in taps.items():
            entry["tap_mean." + k] = np.float64(v.double().mean().item())
            entry["tap_absmax." + k] = np.float32(v.abs().max().item())
        np.savez(out / f"oracle_{tag}.npz", **entry)
        golden[tag] = {"logits": lg.tolist(), "sha256_logits": digest(lg)}

    # ---- uint8 variant through the frames processor (a1) + streaming predictor semantics ------------
    cfg = O.ModelConfig()
    sd = O.make_state_dict(cfg, seed=1234)
    gu = torch.Generator().manual_seed(1)
    u8 = torch.randint(0, 256, (1, 15, 80, 160), dtype=torch.uint8, generator=gu)
    with torch.no_grad():
        lg = O.forward(sd, O.pad_normalize(u8, (160, 96)), cfg)
    np.savez(out / "oracle_u8_small.npz", logits=lg.numpy())
    golden["u8_small"] = {"logits": lg.tolist()}

    report["golden"] = golden
    (out / "oracle_pin_report.json").write_text(json.dumps(report, indent=1, default=float))
    print(json.dumps(report, indent=1, default=float))


if __name__ == "__main__":
    main()
