"""state_dict (reference key names) -> packed tensors for libmds_b200: eval-mode BatchNorm folded into the
preceding conv (scale into the weights, shift as a bias), fp16 K-major GEMM operands, fp32 depthwise / SE /
classifier parameters.

BN eps: 1e-3 for the timm ``tf_`` encoder, 1e-5 (nn.BatchNorm2d/3d default) for the reference-owned layers
(multidim_stacker.py:59,178-185,198-205).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

ENC_EPS, REF_EPS = 1e-3, 1e-5
# (kind, repeats, stride, expand, cout, has_se) — timm arch_def of tf_efficientnetv2_b0
STAGES = [("cn", 1, 1, 1, 16, False), ("er", 2, 2, 4, 32, False), ("er", 2, 2, 4, 48, False),
          ("ir", 3, 2, 4, 96, True), ("ir", 5, 1, 6, 112, True), ("ir", 8, 2, 6, 192, True)]


def _fold(sd, wkey: str, bn: str, eps: float) -> Tuple[torch.Tensor, torch.Tensor]:
    w = sd[wkey].detach().double().cpu()
    gamma, beta = sd[bn + ".weight"].detach().double().cpu(), sd[bn + ".bias"].detach().double().cpu()
    mean, var = sd[bn + ".running_mean"].detach().double().cpu(), sd[bn + ".running_var"].detach().double().cpu()
    scale = gamma / torch.sqrt(var + eps)
    w = w * scale.view(-1, *([1] * (w.ndim - 1)))
    return w, beta - mean * scale


def _h(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.float16).contiguous()


def _f(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.float32).contiguous()


def stem_weights(w: torch.Tensor) -> torch.Tensor:
    """[32][3][3][3] folded stem weights -> fp16 [2][32][32]: (hi, lo) x cout x k with k = (ci*3 + r)*3 + s, k >= 27 zero.
    hi + lo reproduces the fp32 weight to ~2^-22, so the tensor-core stem is as accurate as an fp32 convolution."""
    w = w.detach().double().cpu().reshape(32, 27)
    hi = w.to(torch.float16)
    lo = (w - hi.double()).to(torch.float16)
    m = torch.zeros((2, 32, 32), dtype=torch.float16)
    m[0, :, :27], m[1, :, :27] = hi, lo
    return m.contiguous()


def bias_matrix(b: torch.Tensor) -> torch.Tensor:
    """[N][64] fp16 operand of the tcgen05 GEMM's bias MMA: col 0 = fp16(b), col 1 = fp16(b - col 0), rest 0.
    Multiplied by a ones tile on the tensor core, hi + lo reproduces the fp32 bias to ~2^-22."""
    b = b.detach().double().cpu()
    hi = b.to(torch.float16)
    lo = (b - hi.double()).to(torch.float16)
    m = torch.zeros((b.numel(), 64), dtype=torch.float16)
    m[:, 0], m[:, 1] = hi, lo
    return m.contiguous()


_GH = None


def _silu_mean(beta: torch.Tensor, gamma: torch.Tensor) -> torch.Tensor:
    """E[SiLU(z)], z ~ N(beta, gamma^2) per channel (48-point Gauss-Hermite): the mean of a BN+SiLU output whose
    running statistics match its input, i.e. of any trained BatchNorm layer."""
    global _GH
    if _GH is None:
        import numpy as np
        x, w = np.polynomial.hermite_e.hermegauss(48)
        _GH = (torch.tensor(x, dtype=torch.float64), torch.tensor(w / w.sum(), dtype=torch.float64))
    z = beta.double()[:, None] + gamma.double().abs()[:, None] * _GH[0][None, :]
    return (torch.nn.functional.silu(z) * _GH[1][None, :]).sum(1)


def _rounding_bias(w: torch.Tensor, mu) -> torch.Tensor:
    """Expected output shift caused by rounding the folded weights to fp16: sum_k (fp16(w) - w)[n, k] * E[x_k].
    With few, mostly positive (post-SiLU) input channels this shift is the same at every pixel, so it survives the
    GeM average; subtracting it from the bias is the data-free "bias correction" of post-training quantisation
    (Nagel et al., DFQ, ICCV 2019), with E[x] taken from the preceding BatchNorm's (beta, gamma)."""
    if mu is None:
        return torch.zeros(w.shape[0], dtype=torch.float64)
    dw = w.to(torch.float16).double() - w
    return (dw.reshape(w.shape[0], w.shape[1], -1).sum(-1) * mu[None, :]).sum(1)


def pack_state_dict(sd: Dict[str, torch.Tensor], num_3d_blocks: int, bias_correction: bool = True) -> Dict[str, torch.Tensor]:
    out: Dict[str, torch.Tensor] = {}
    e = "conv2d_encoder."
    w, b = _fold(sd, e + "conv_stem.weight", e + "bn1", ENC_EPS)           # [32][3][3][3]
    out["stem.wh"] = stem_weights(w)
    out["stem.b"] = _f(b)

    def beta(bn):
        return sd[bn + ".bias"].detach().double().cpu()

    def act_mean(bn):
        return _silu_mean(sd[bn + ".bias"].detach().cpu(), sd[bn + ".weight"].detach().cpu())

    def conv3(name, wkey, bn, mu=None):
        w, b = _fold(sd, wkey, bn, ENC_EPS)                                # [co][ci][3][3]
        if bias_correction:
            b = b - _rounding_bias(w, mu)
        out[name + ".w"] = _h(w.permute(0, 2, 3, 1).reshape(w.shape[0], -1))   # [co][(r*3+s)*ci + c]
        out[name + ".b"] = _f(b)

    def pw(name, wkey, bn, eps, bias_mat=False, gated=False, mu=None):
        w, b = _fold(sd, wkey, bn, eps)
        if bias_correction and not gated:      # gated weights are rounded per image at run time: nothing static to correct
            b = b - _rounding_bias(w, mu)
        out[name + ".w"] = _h(w.reshape(w.shape[0], w.shape[1]))            # [co][ci]
        out[name + ".b"] = _f(b)
        if bias_mat or gated:
            out[name + ".bm"] = bias_matrix(b)
        if gated:       # SE-gated projection: the SE kernel multiplies these fp32 weights by the gate, one rounding to fp16
            out[name + ".w32"] = _f(w.reshape(w.shape[0], w.shape[1]))

    def dw(name, wkey, bn, eps):
        w, b = _fold(sd, wkey, bn, eps)                                    # [C][1][(kt)][3][3]
        out[name + ".w"] = _f(w.reshape(w.shape[0], -1).t())                # [taps][C]
        out[name + ".b"] = _f(b)

    def se(name, prefix):
        w1, w2 = sd[prefix + ".conv_reduce.weight"], sd[prefix + ".conv_expand.weight"]
        out[name + ".w1"] = _f(w1.detach().cpu().reshape(w1.shape[0], w1.shape[1]))           # [rd][C]
        out[name + ".b1"] = _f(sd[prefix + ".conv_reduce.bias"].detach().cpu())
        out[name + ".w2t"] = _f(w2.detach().cpu().reshape(w2.shape[0], w2.shape[1]).t())      # [rd][C]
        out[name + ".b2"] = _f(sd[prefix + ".conv_expand.bias"].detach().cpu())

    # mu = expected per-channel mean of the tensor that flows between blocks (used only for the bias correction)
    mu = act_mean(e + "bn1")
    cin = 32
    for si, (kind, reps, stride, _, cout, _) in enumerate(STAGES):
        for bi in range(reps):
            p, n = f"{e}blocks.{si}.{bi}.", f"b{si}.{bi}"
            skip = kind != "cn" and (stride if bi == 0 else 1) == 1 and cin == cout
            if kind == "cn":
                conv3(n + ".c3", p + "conv.weight", p + "bn1", mu)
                mu = act_mean(p + "bn1")
            elif kind == "er":
                conv3(n + ".c3", p + "conv_exp.weight", p + "bn1", mu)
                pw(n + ".pwl", p + "conv_pwl.weight", p + "bn2", ENC_EPS, mu=act_mean(p + "bn1"))
                mu = beta(p + "bn2") + (mu if skip else 0)
            else:
                pw(n + ".pw", p + "conv_pw.weight", p + "bn1", ENC_EPS, bias_mat=True, mu=mu)
                dw(n + ".dw", p + "conv_dw.weight", p + "bn2", ENC_EPS)
                se(n + ".se", p + "se")
                pw(n + ".pwl", p + "conv_pwl.weight", p + "bn3", ENC_EPS, gated=True)
                mu = beta(p + "bn3") + (mu if skip else 0)
            cin = cout

    pw("proj2d", "conv2d_projection.0.weight", "conv2d_projection.1", REF_EPS, bias_mat=True, mu=mu)
    mu = act_mean("conv2d_projection.1")
    for i in range(num_3d_blocks):
        p, n = f"conv3d_encoder.{i}.", f"c3d.{i}"
        pw(n + ".pw", p + "conv_pw.weight", p + "bn1.bn3d", REF_EPS, bias_mat=True, mu=mu)
        dw(n + ".dw", p + "conv_dw.weight", p + "bn2.bn3d", REF_EPS)
        se(n + ".se", p + "se")
        pw(n + ".pwl", p + "conv_pwl.weight", p + "bn3.bn3d", REF_EPS, gated=True)
        mu = beta(p + "bn3.bn3d") + mu            # the 3D shortcut is unconditional (multidim_stacker.py:133)
    pw("proj3d", "conv3d_projection.0.weight", "conv3d_projection.1", REF_EPS, bias_mat=True, mu=mu)
    out["gem.p"] = _f(sd["global_pool.p"].detach().cpu().reshape(1))
    out["cls.w"] = _f(sd["classifier.weight"].detach().cpu())
    out["cls.b"] = _f(sd["classifier.bias"].detach().cpu())
    return out
