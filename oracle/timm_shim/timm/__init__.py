"""Minimal stand-in for ``timm==0.9.2`` (requirements.txt:9 of the reference) — TEST INFRASTRUCTURE ONLY.

timm is an un-vendored, un-installable dependency here.  The reference's ``src/models/multidim_stacker.py``
imports exactly five names from it (:11-17).  This package provides those five so that the *unmodified*
reference file can be imported by path in the authoring container (``oracle/make_golden.py``).  It is put on
``sys.path`` only by that script and by the CPU tests that re-validate the oracle when ``/root/reference``
exists; nothing in the product imports it.

The encoder is a module-style restatement of timm's ``EfficientNetFeatures`` for ``tf_efficientnetv2_b0``
with timm's parameter names, written independently from the functional oracle in ``oracle/mds_oracle.py``
so that the two can be checked against each other.
"""
import math

import torch
from torch import nn
import torch.nn.functional as F

from . import layers  # noqa: F401
from .layers import BatchNormAct2d, create_conv2d

__version__ = "0.9.2-shim"

_B0 = dict(
    stem=32,
    # (type, repeats, kernel, stride, expand, out, se_ratio)
    stages=[("cn", 1, 3, 1, 1, 16, 0.0), ("er", 2, 3, 2, 4, 32, 0.0), ("er", 2, 3, 2, 4, 48, 0.0),
            ("ir", 3, 3, 2, 4, 96, 0.25), ("ir", 5, 3, 1, 6, 112, 0.25), ("ir", 8, 3, 2, 6, 192, 0.25)],
)


class _SE(nn.Module):
    def __init__(self, chs, rd):
        super().__init__()
        self.conv_reduce = nn.Conv2d(chs, rd, 1, bias=True)
        self.act1 = nn.SiLU(inplace=True)
        self.conv_expand = nn.Conv2d(rd, chs, 1, bias=True)
        self.gate = nn.Sigmoid()

    def forward(self, x):
        s = x.mean((2, 3), keepdim=True)
        return x * self.gate(self.conv_expand(self.act1(self.conv_reduce(s))))


class _ConvBnAct(nn.Module):
    def __init__(self, cin, cout, k, s, bn):
        super().__init__()
        self.conv = create_conv2d(cin, cout, k, stride=s, padding="same")
        self.bn1 = bn(cout)

    def forward(self, x):
        return self.bn1(self.conv(x))


class _EdgeResidual(nn.Module):
    def __init__(self, cin, cout, k, s, e, bn):
        super().__init__()
        mid = cin * e
        self.has_skip = s == 1 and cin == cout
        self.conv_exp = create_conv2d(cin, mid, k, stride=s, padding="same")
        self.bn1 = bn(mid)
        self.se = nn.Identity()
        self.conv_pwl = create_conv2d(mid, cout, 1, padding="same")
        self.bn2 = bn(cout, apply_act=False)

    def forward(self, x):
        y = self.bn2(self.conv_pwl(self.se(self.bn1(self.conv_exp(x)))))
        return y + x if self.has_skip else y


class _InvertedResidual(nn.Module):
    def __init__(self, cin, cout, k, s, e, se_ratio, bn):
        super().__init__()
        mid = cin * e
        self.has_skip = s == 1 and cin == cout
        self.conv_pw = create_conv2d(cin, mid, 1, padding="same")
        self.bn1 = bn(mid)
        self.conv_dw = create_conv2d(mid, mid, k, stride=s, padding="same", depthwise=True)
        self.bn2 = bn(mid)
        self.se = _SE(mid, int(round(cin * se_ratio))) if se_ratio > 0 else nn.Identity()
        self.conv_pwl = create_conv2d(mid, cout, 1, padding="same")
        self.bn3 = bn(cout, apply_act=False)

    def forward(self, x):
        y = self.bn2(self.conv_dw(self.bn1(self.conv_pw(x))))
        y = self.bn3(self.conv_pwl(self.se(y)))
        return y + x if self.has_skip else y


class EfficientNetFeatures(nn.Module):
    """features_only wrapper: conv_stem, bn1, blocks; returns the stage outputs named by out_indices."""

    def __init__(self, in_chans=3, out_indices=(4,), eps=1e-3):
        super().__init__()
        def bn(c, apply_act=True):
            return BatchNormAct2d(c, eps=eps, apply_act=apply_act, act_layer=nn.SiLU)
        self.conv_stem = create_conv2d(in_chans, _B0["stem"], 3, stride=2, padding="same")
        self.bn1 = bn(_B0["stem"])
        stages, cin = [], _B0["stem"]
        self.feature_info, reduction = [], 2
        for kind, reps, k, s, e, cout, se in _B0["stages"]:
            blocks = []
            for r in range(reps):
                st = s if r == 0 else 1
                if kind == "cn":
                    blocks.append(_ConvBnAct(cin, cout, k, st, bn))
                elif kind == "er":
                    blocks.append(_EdgeResidual(cin, cout, k, st, e, bn))
                else:
                    blocks.append(_InvertedResidual(cin, cout, k, st, e, se, bn))
                cin = cout
            stages.append(nn.Sequential(*blocks))
        self.blocks = nn.Sequential(*stages)
        # timm feature levels for this arch: stage 0 (/2), 1 (/4), 2 (/8), 4 (/16), 5 (/32)
        self._feature_stage = [0, 1, 2, 4, 5]
        chs = [16, 32, 48, 112, 192]
        self.feature_info = [dict(num_chs=c, reduction=2 ** (i + 1), module=f"blocks.{s}")
                             for i, (c, s) in enumerate(zip(chs, self._feature_stage))]
        self.out_indices = tuple(out_indices)

    def forward(self, x):
        x = self.bn1(self.conv_stem(x))
        feats = {}
        for si, stage in enumerate(self.blocks):
            x = stage(x)
            feats[si] = x
        return [feats[self._feature_stage[i]] for i in self.out_indices]


def create_model(model_name, pretrained=False, num_classes=1000, in_chans=3, drop_rate=0.0, drop_path_rate=0.0,
                 features_only=False, out_indices=None, **kwargs):
    name = model_name.split(".")[0]
    if name != "tf_efficientnetv2_b0" or not features_only:
        raise NotImplementedError(f"timm shim only provides tf_efficientnetv2_b0 features_only, got {model_name}")
    if pretrained:
        raise RuntimeError("timm shim: no network, pretrained weights unavailable")
    return EfficientNetFeatures(in_chans=in_chans, out_indices=out_indices or (4,))
