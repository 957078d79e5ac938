#!/usr/bin/env python
"""CPU emulation of the engine's storage precision (which fp16 roundings cost how much logit error).

    python tests/precision_study.py [N stacks] [variant ...]

Every variant applies a subset of the engine's roundings to the fp32 oracle: `act:<tap-regex>` rounds the tapped
activations to fp16 (or to an fp16 hi/lo pair with `act2:`), `w:<key-regex>` rounds the BN-folded conv weights whose
state-dict key matches to fp16.  Reports the logit error RMS / max relative to max|ref| over N full-size stacks.
Lives under tests/ because it drives the oracle (checker only)."""
import json
import re
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import mds_oracle as O  # noqa: E402

torch.set_grad_enabled(False)


def fold_round(sd, wkey, bn, eps, hilo=False):
    """Replace sd[wkey] by the weight whose BN-folded value is fp16-rounded (what the packer stores)."""
    w = sd[wkey].double()
    scale = sd[bn + ".weight"].double() / torch.sqrt(sd[bn + ".running_var"].double() + eps)
    s = scale.view(-1, *([1] * (w.ndim - 1)))
    wf = w * s
    hi = wf.to(torch.float16).double()
    if hilo:
        hi = hi + (wf - hi).to(torch.float16).double()
    sd[wkey] = (hi / s).float()


def conv_bn_pairs(cfg):
    e = "conv2d_encoder."
    pairs = [(e + "conv_stem.weight", e + "bn1", O.ENC_BN_EPS)]
    for si, stage in enumerate(O.encoder_arch()):
        for bi, b in enumerate(stage):
            p = f"{e}blocks.{si}.{bi}."
            if b.kind == "cn":
                pairs.append((p + "conv.weight", p + "bn1", O.ENC_BN_EPS))
            elif b.kind == "er":
                pairs += [(p + "conv_exp.weight", p + "bn1", O.ENC_BN_EPS), (p + "conv_pwl.weight", p + "bn2", O.ENC_BN_EPS)]
            else:
                pairs += [(p + "conv_pw.weight", p + "bn1", O.ENC_BN_EPS), (p + "conv_pwl.weight", p + "bn3", O.ENC_BN_EPS)]
    pairs.append(("conv2d_projection.0.weight", "conv2d_projection.1", O.REF_BN_EPS))
    for i in range(cfg.num_3d_blocks):
        p = f"conv3d_encoder.{i}."
        pairs += [(p + "conv_pw.weight", p + "bn1.bn3d", O.REF_BN_EPS), (p + "conv_pwl.weight", p + "bn3.bn3d", O.REF_BN_EPS)]
    pairs.append(("conv3d_projection.0.weight", "conv3d_projection.1", O.REF_BN_EPS))
    return pairs


def make_variant(sd0, cfg, spec):
    """spec: list of 'act:<re>', 'act2:<re>', 'w:<re>'.  -> (sd, tap)"""
    sd = dict(sd0)
    act_re = [re.compile(s[4:]) for s in spec if s.startswith("act:")]
    act2_re = [re.compile(s[5:]) for s in spec if s.startswith("act2:")]
    w_re = [re.compile(s[2:]) for s in spec if s.startswith("w:")]
    wc_re = [re.compile(s[3:]) for s in spec if s.startswith("wc:")]      # fp16 weights + the packer's data-free bias correction
    if wc_re:
        from ball_action_spotting_b200.packer import pack_state_dict
        p0 = pack_state_dict(sd0, cfg.num_3d_blocks, bias_correction=False)
        p1 = pack_state_dict(sd0, cfg.num_3d_blocks, bias_correction=True)
    for wkey, bn, eps in conv_bn_pairs(cfg):
        if any(r.search(wkey) for r in w_re):
            fold_round(sd, wkey, bn, eps)
        if any(r.search(wkey) for r in wc_re):
            fold_round(sd, wkey, bn, eps)
            m = re.match(r"conv2d_encoder\.blocks\.(\d)\.(\d)\.(conv|conv_exp|conv_pwl|conv_pw)\.weight", wkey)
            assert m, wkey
            pk = f"b{m.group(1)}.{m.group(2)}." + {"conv": "c3", "conv_exp": "c3", "conv_pwl": "pwl", "conv_pw": "pw"}[m.group(3)] + ".b"
            sd[bn + ".bias"] = sd[bn + ".bias"] + (p1[pk] - p0[pk])

    sr_re = [re.compile(s[3:]) for s in spec if s.startswith("sr:")]       # stochastic rounding to fp16 (unbiased)
    gen = torch.Generator().manual_seed(7)

    def sround(x):
        bits = x.contiguous().view(torch.int32)
        r = torch.randint(0, 1 << 13, bits.shape, generator=gen, dtype=torch.int32)
        return ((bits + r) & ~((1 << 13) - 1)).view(torch.float32)

    ed_re = [re.compile(s[3:]) for s in spec if s.startswith("ed:")]       # error diffusion along W (full row chain)
    ep_re = [re.compile(s[3:]) for s in spec if s.startswith("ep:")]       # error feedback inside pixel pairs (w, w+1)

    def diffuse(x, pair):
        out = torch.empty_like(x)
        carry = torch.zeros_like(x[..., 0])
        for w_ in range(x.shape[-1]):
            v = x[..., w_] + carry
            q = v.half().float()
            out[..., w_] = q
            carry = v - q
            if pair and (w_ & 1):
                carry = torch.zeros_like(carry)
        return out

    def tap(name, x):
        if any(r.fullmatch(name) for r in ed_re):
            return diffuse(x, False)
        if any(r.fullmatch(name) for r in ep_re):
            return diffuse(x, True)
        if any(r.fullmatch(name) for r in sr_re):
            return sround(x)
        if any(r.fullmatch(name) for r in act2_re):
            hi = x.half().float()
            return hi + (x - hi).half().float()
        if any(r.fullmatch(name) for r in act_re):
            return x.half().float()
        return None
    return sd, tap


VARIANTS = {
    "all": ["act:.*", "w:.*"],
    "act_all": ["act:.*"],
    "w_all": ["w:.*"],
    "act_stem": ["act:stem"],
    "act_stream_early": [r"act:b[012]\.\d"],
    "act_exp_early": [r"act:b[012]\.\d\.exp"],
    "act_stream_late": [r"act:b[345]\.\d"],
    "act_mid_late": [r"act:b[345]\.\d\.(exp|dw)"],
    "act_3d_stream": [r"act:(proj2d|c3d\.\d)"],
    "act_3d_mid": [r"act:c3d\.\d\.(exp|dw)"],
    "w_early": [r"w:blocks\.[012]\."],
    "w_late_pw": [r"w:blocks\.[345]\.\d\.conv_pw\."],
    "w_late_pwl": [r"w:blocks\.[345]\.\d\.conv_pwl"],
    "w_3d": [r"w:(conv3d|conv2d_projection)"],
    "wc_early": [r"wc:blocks\.[012]\."],
    "w_b0": [r"w:blocks\.0\."], "w_b10": [r"w:blocks\.1\.0\."], "w_b11": [r"w:blocks\.1\.1\."],
    "w_b20": [r"w:blocks\.2\.0\."], "w_b21": [r"w:blocks\.2\.1\."],
    "w_early_exp": [r"w:blocks\.[012]\.\d\.(conv|conv_exp)\."], "w_early_pwl": [r"w:blocks\.[012]\.\d\.conv_pwl"],
    "sr_all": ["sr:.*"], "sr_stem": ["sr:stem"], "sr_stream": [r"sr:(stem|b\d\.\d|proj2d|c3d\.\d)"],
    "ed_stem": ["ed:stem"], "ep_stem": ["ep:stem"],
    "ed_early": [r"ed:(stem|b[012]\.\d)"], "ep_early": [r"ep:(stem|b[012]\.\d)"], "rn_early": [r"act:(stem|b[012]\.\d)"],
    "ed_stream": [r"ed:(stem|b\d\.\d)"], "rn_stream": [r"act:(stem|b\d\.\d)"],
    "optX": ["act:.*", r"wc:blocks\.[012]\."],
    "optY": ["act:.*"],
    # candidate builds: what stays fp16
    "cand_A": [r"act:b[012]\.\d\.exp", r"act:b[345]\.\d\.(exp|dw)", r"act:c3d\.\d\.(exp|dw)", "act:proj3d"],
    "cand_B": [r"act:b[345]\.\d\.(exp|dw)", r"act:c3d\.\d\.(exp|dw)", "act:proj3d"],
}


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    names = sys.argv[2:] or list(VARIANTS)
    cfg = O.ModelConfig()
    sd0 = O.make_state_dict(cfg, seed=1234)
    u8 = torch.randint(0, 256, (n, 15, 720, 1280), dtype=torch.uint8, generator=torch.Generator().manual_seed(100))
    xs = [O.pad_normalize(u8[i:i + 1], (1280, 736)) for i in range(n)]
    ref = torch.cat([O.forward(sd0, x, cfg) for x in xs])
    scale = ref.abs().max()
    print("ref absmax", scale.item(), flush=True)
    for name in names:
        spec = VARIANTS.get(name) or name.split(",")
        sd, tap = make_variant(sd0, cfg, spec)
        got = torch.cat([O.forward(sd, x, cfg, tap) for x in xs])
        err = (got - ref).abs() / scale
        print(json.dumps({"variant": name, "rms": round(err.pow(2).mean().sqrt().item(), 7), "max": round(err.max().item(), 7),
                          "mean_signed": round(((got - ref) / scale).mean().item(), 7)}), flush=True)


if __name__ == "__main__":
    main()
