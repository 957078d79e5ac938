"""CPU fp32 oracle for the MultiDimStacker forward path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this file.  The product package (``ball_action_spotting_b200``) never does.

What it restates (citations relative to ``/root/reference``):

* ``src/models/multidim_stacker.py:210-243``  forward_2d / forward_3d / forward_head / forward
* ``src/models/multidim_stacker.py:93-134``   InvertedResidual3d (conv_pw, bn1, conv_dw, bn2, se, conv_pwl, bn3, + shortcut)
* ``src/models/multidim_stacker.py:72-90``    3D SqueezeExcite (mean over T,H,W; rd = C // ratio)
* ``src/models/multidim_stacker.py:20-45``    GeneralizedMeanPooling (clamp(eps).pow(p) -> mean -> pow(1/p))
* ``src/frames.py:7-54``                      pad_to_frames + normalize_frames (PadNormalizeFramesProcessor)
* timm==0.9.2 ``tf_efficientnetv2_b0`` features_only, out_indices=[4]  (requirements.txt:9; call site
  ``multidim_stacker.py:166-176``).  timm is an un-vendored dependency that is absent from the reference
  tree and from this image, so its published algorithm is restated here: TF-"SAME"
  padding on stride-2 convs, BN eps 1e-3, SiLU, SE reduce width round(block_in * 0.25).

The oracle is *functional* (plain ``torch.nn.functional`` calls over a ``state_dict`` whose keys are the
reference's), so every intermediate tensor can be tapped for per-kernel parity tests.

Parity pinning: the reference has no tests or golden vectors for this path.  The pins
are produced in the authoring container by ``oracle/make_golden.py``, which (i) imports the reference's own
``multidim_stacker.py`` unmodified by file path on top of the ``oracle/timm_shim`` encoder, (ii) checks this
functional restatement against it bit-for-bit on identical weights, (iii) cross-checks the encoder against
an independent torchvision ``EfficientNet`` assembly, and (iv) commits small fixtures under ``tests/golden``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
StateDict = Dict[str, Tensor]

ENC_BN_EPS = 1e-3   # timm "tf_" models
REF_BN_EPS = 1e-5   # nn.BatchNorm2d / nn.BatchNorm3d default used by the reference-owned layers


# --------------------------------------------------------------------------------------------------------
# Architecture table (timm arch_def for tf_efficientnetv2_b0)
# --------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class BlockSpec:
    kind: str          # "cn" ConvBnAct | "er" EdgeResidual (FusedMBConv) | "ir" InvertedResidual (MBConv)
    cin: int
    cout: int
    stride: int
    expand: int        # expansion ratio (1 for cn)
    se_rd: int         # SE reduce width (0 = no SE)

    @property
    def mid(self) -> int:
        return self.cin * self.expand

    @property
    def has_skip(self) -> bool:
        return self.kind != "cn" and self.stride == 1 and self.cin == self.cout


STEM_CH = 32
_STAGE_DEFS = [
    # kind, repeats, stride, expand, cout, se_ratio
    ("cn", 1, 1, 1, 16, 0.0),
    ("er", 2, 2, 4, 32, 0.0),
    ("er", 2, 2, 4, 48, 0.0),
    ("ir", 3, 2, 4, 96, 0.25),
    ("ir", 5, 1, 6, 112, 0.25),
    ("ir", 8, 2, 6, 192, 0.25),
]


def encoder_arch() -> List[List[BlockSpec]]:
    stages: List[List[BlockSpec]] = []
    cin = STEM_CH
    for kind, reps, stride, expand, cout, se in _STAGE_DEFS:
        blocks = []
        for r in range(reps):
            s = stride if r == 0 else 1
            rd = int(round(cin * se)) if se > 0 else 0
            blocks.append(BlockSpec(kind, cin, cout, s, expand, rd))
            cin = cout
        stages.append(blocks)
    return stages


ENCODER_OUT_CH = 192


@dataclass(frozen=True)
class ModelConfig:
    """kwargs of MultiDimStacker.__init__ that change shapes (multidim_stacker.py:138-153)."""
    num_classes: int = 2
    num_frames: int = 15
    stack_size: int = 3
    num_3d_blocks: int = 4
    num_3d_features: int = 192
    num_3d_stack_proj: int = 256
    expansion_3d_ratio: int = 3
    se_reduce_3d_ratio: int = 24

    @property
    def num_stacks(self) -> int:
        return self.num_frames // self.stack_size

    @property
    def num_features(self) -> int:
        return self.num_3d_stack_proj * self.num_stacks

    @property
    def mid_3d(self) -> int:
        return self.num_3d_features * self.expansion_3d_ratio

    @property
    def rd_3d(self) -> int:
        return self.mid_3d // self.se_reduce_3d_ratio


# --------------------------------------------------------------------------------------------------------
# Seeded "meaningful random weights" 
# --------------------------------------------------------------------------------------------------------
def _bn_keys(prefix: str) -> Tuple[str, str, str, str, str]:
    return (prefix + ".weight", prefix + ".bias", prefix + ".running_mean", prefix + ".running_var",
            prefix + ".num_batches_tracked")


def make_state_dict(cfg: ModelConfig = ModelConfig(), seed: int = 1234, recipe: str = "calibrated",
                    calib_hw: Tuple[int, int] = (192, 320)) -> StateDict:
    """State dict with the reference's key names and shapes, filled from a fixed torch.Generator.

    conv weights ~ N(0, sqrt(2 / fan_out)) (timm "goog" init), SE/classifier biases small random,
    every BN perturbed (weight U(0.5,1.5), bias N(0,0.1), mean N(0,0.1), var U(0.5,1.5)) so that BN folding
    is exercised and logits are not bias dominated.

    recipe="calibrated" (default) additionally (a) draws the last BN gamma of every residual branch from
    U(0.25, 0.75) and (b) replaces every BN's running_mean / running_var by the statistics its input really
    has on a seeded random clip (one sequential calibration pass), which is the defining property of a
    *trained* BN network: post-BN activations are O(1) at every depth.  recipe="raw" skips (a) and (b); that
    network is badly conditioned (activations grow to ~850, SE pre-activations ~50) and is kept as a stress
    case (see DESIGN.md "Numerics").
    """
    assert recipe in ("calibrated", "raw")
    g = torch.Generator().manual_seed(seed)
    sd: StateDict = {}

    def conv(name: str, cout: int, cin_per_group: int, *k: int, groups: int = 1, gain: float = 1.0):
        fan_out = cout * math.prod(k) // groups
        w = torch.randn((cout, cin_per_group, *k), generator=g) * (gain * math.sqrt(2.0 / fan_out))
        sd[name] = w

    def bn(prefix: str, c: int):
        kw, kb, km, kv, kn = _bn_keys(prefix)
        sd[kw] = torch.rand(c, generator=g) + 0.5
        sd[kb] = torch.randn(c, generator=g) * 0.1
        sd[km] = torch.randn(c, generator=g) * 0.1
        sd[kv] = torch.rand(c, generator=g) + 0.5
        sd[kn] = torch.tensor(0, dtype=torch.long)

    def se(prefix: str, c: int, rd: int, *k: int):
        conv(prefix + ".conv_reduce.weight", rd, c, *k)
        sd[prefix + ".conv_reduce.bias"] = torch.randn(rd, generator=g) * 0.1
        conv(prefix + ".conv_expand.weight", c, rd, *k)
        sd[prefix + ".conv_expand.bias"] = torch.randn(c, generator=g) * 0.1

    e = "conv2d_encoder."
    conv(e + "conv_stem.weight", STEM_CH, cfg.stack_size, 3, 3)
    bn(e + "bn1", STEM_CH)
    for si, stage in enumerate(encoder_arch()):
        for bi, b in enumerate(stage):
            p = f"{e}blocks.{si}.{bi}."
            if b.kind == "cn":
                conv(p + "conv.weight", b.cout, b.cin, 3, 3)
                bn(p + "bn1", b.cout)
            elif b.kind == "er":
                conv(p + "conv_exp.weight", b.mid, b.cin, 3, 3)
                bn(p + "bn1", b.mid)
                conv(p + "conv_pwl.weight", b.cout, b.mid, 1, 1)
                bn(p + "bn2", b.cout)
            else:
                conv(p + "conv_pw.weight", b.mid, b.cin, 1, 1)
                bn(p + "bn1", b.mid)
                conv(p + "conv_dw.weight", b.mid, 1, 3, 3, groups=b.mid)
                bn(p + "bn2", b.mid)
                se(p + "se", b.mid, b.se_rd, 1, 1)
                conv(p + "conv_pwl.weight", b.cout, b.mid, 1, 1)
                bn(p + "bn3", b.cout)

    c3, mid, rd = cfg.num_3d_features, cfg.mid_3d, cfg.rd_3d
    conv("conv2d_projection.0.weight", c3, ENCODER_OUT_CH, 1, 1)
    bn("conv2d_projection.1", c3)
    for i in range(cfg.num_3d_blocks):
        p = f"conv3d_encoder.{i}."
        conv(p + "conv_pw.weight", mid, c3, 1, 1, 1)
        bn(p + "bn1.bn3d", mid)
        conv(p + "conv_dw.weight", mid, 1, 3, 3, 3, groups=mid)
        bn(p + "bn2.bn3d", mid)
        se(p + "se", mid, rd, 1, 1, 1)
        conv(p + "conv_pwl.weight", c3, mid, 1, 1, 1)
        bn(p + "bn3.bn3d", c3)
    conv("conv3d_projection.0.weight", cfg.num_3d_stack_proj, c3, 1, 1)
    bn("conv3d_projection.1", cfg.num_3d_stack_proj)
    sd["global_pool.p"] = torch.ones(1) * 3.0
    sd["classifier.weight"] = torch.randn(cfg.num_classes, cfg.num_features, generator=g) * 0.05
    sd["classifier.bias"] = torch.randn(cfg.num_classes, generator=g) * 0.1
    if recipe == "calibrated":
        for si, stage in enumerate(encoder_arch()):
            for bi, b in enumerate(stage):
                if b.has_skip:
                    k = f"{e}blocks.{si}.{bi}.{'bn2' if b.kind == 'er' else 'bn3'}.weight"
                    sd[k] = torch.rand(sd[k].shape, generator=g) * 0.5 + 0.25
        for i in range(cfg.num_3d_blocks):
            k = f"conv3d_encoder.{i}.bn3.bn3d.weight"
            sd[k] = torch.rand(sd[k].shape, generator=g) * 0.5 + 0.25
        x = torch.rand((2, cfg.num_frames, *calib_hw), generator=g)
        global _CALIBRATING
        _CALIBRATING = True
        try:
            with torch.no_grad():
                forward(sd, x, cfg)
        finally:
            _CALIBRATING = False
    return sd


# --------------------------------------------------------------------------------------------------------
# Functional layers
# --------------------------------------------------------------------------------------------------------
Tap = Optional[Callable[[str, Tensor], Optional[Tensor]]]


def _tap(tap: Tap, name: str, x: Tensor) -> Tensor:
    """tap may observe an intermediate and optionally return a replacement (used for fp16 emulation)."""
    if tap is None:
        return x
    y = tap(name, x)
    return x if y is None else y


_CALIBRATING = False


def _bn(sd: StateDict, prefix: str, x: Tensor, eps: float) -> Tensor:
    kw, kb, km, kv, _ = _bn_keys(prefix)
    if _CALIBRATING:        # make_state_dict(recipe="calibrated"): running stats := stats of the real input
        dims = [0] + list(range(2, x.ndim))
        sd[km] = x.double().mean(dims).float()
        sd[kv] = x.double().var(dims, unbiased=False).float()
    return F.batch_norm(x, sd[km], sd[kv], sd[kw], sd[kb], training=False, eps=eps)


def _same_pad(x: Tensor, k: int, s: int) -> Tensor:
    """timm ``pad_same``: total = max((ceil(i/s)-1)*s + k - i, 0); left = total//2, right = total-left."""
    ih, iw = x.shape[-2:]
    ph = max((math.ceil(ih / s) - 1) * s + k - ih, 0)
    pw = max((math.ceil(iw / s) - 1) * s + k - iw, 0)
    if ph or pw:
        x = F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
    return x


def _conv_same(x: Tensor, w: Tensor, stride: int, groups: int = 1) -> Tensor:
    k = w.shape[-1]
    if stride == 1:
        return F.conv2d(x, w, None, 1, (k - 1) // 2, 1, groups)      # static symmetric padding
    return F.conv2d(_same_pad(x, k, stride), w, None, stride, 0, 1, groups)


def _se2d(sd: StateDict, p: str, x: Tensor) -> Tensor:
    s = x.mean((2, 3), keepdim=True)
    s = F.silu(F.conv2d(s, sd[p + ".conv_reduce.weight"], sd[p + ".conv_reduce.bias"]))
    s = F.conv2d(s, sd[p + ".conv_expand.weight"], sd[p + ".conv_expand.bias"])
    return x * torch.sigmoid(s)


def encoder_forward(sd: StateDict, x: Tensor, tap: Tap = None, prefix: str = "conv2d_encoder.") -> Tensor:
    """(n, 3, H, W) f32 -> (n, 192, H/32, W/32): stem + all six stages; returns feature index 4."""
    e = prefix
    x = F.silu(_bn(sd, e + "bn1", _conv_same(x, sd[e + "conv_stem.weight"], 2), ENC_BN_EPS))
    x = _tap(tap, "stem", x)
    for si, stage in enumerate(encoder_arch()):
        for bi, b in enumerate(stage):
            p = f"{e}blocks.{si}.{bi}."
            name = f"b{si}.{bi}"
            sc = x
            if b.kind == "cn":
                x = F.silu(_bn(sd, p + "bn1", _conv_same(x, sd[p + "conv.weight"], b.stride), ENC_BN_EPS))
            elif b.kind == "er":
                x = F.silu(_bn(sd, p + "bn1", _conv_same(x, sd[p + "conv_exp.weight"], b.stride), ENC_BN_EPS))
                x = _tap(tap, name + ".exp", x)
                x = _bn(sd, p + "bn2", F.conv2d(x, sd[p + "conv_pwl.weight"]), ENC_BN_EPS)
            else:
                x = F.silu(_bn(sd, p + "bn1", F.conv2d(x, sd[p + "conv_pw.weight"]), ENC_BN_EPS))
                x = _tap(tap, name + ".exp", x)
                x = _conv_same(x, sd[p + "conv_dw.weight"], b.stride, groups=b.mid)
                x = F.silu(_bn(sd, p + "bn2", x, ENC_BN_EPS))
                x = _tap(tap, name + ".dw", x)
                x = _se2d(sd, p + "se", x)
                x = _tap(tap, name + ".se", x)
                x = _bn(sd, p + "bn3", F.conv2d(x, sd[p + "conv_pwl.weight"]), ENC_BN_EPS)
            if b.has_skip:
                x = x + sc
            x = _tap(tap, name, x)
    return x


def forward_2d(sd: StateDict, x: Tensor, cfg: ModelConfig = ModelConfig(), tap: Tap = None) -> Tensor:
    """multidim_stacker.py:210-219.  (b, t, H, W) -> (b, t/stack, 192, h, w)."""
    b, t, h, w = x.shape
    assert t % cfg.stack_size == 0
    k = t // cfg.stack_size
    x = x.reshape(b * k, cfg.stack_size, h, w)
    x = encoder_forward(sd, x, tap)
    x = F.conv2d(x, sd["conv2d_projection.0.weight"])
    x = F.silu(_bn(sd, "conv2d_projection.1", x, REF_BN_EPS))
    x = _tap(tap, "proj2d", x)
    return x.reshape(b, k, cfg.num_3d_features, x.shape[-2], x.shape[-1])


def block3d(sd: StateDict, p: str, x: Tensor, cfg: ModelConfig, tap: Tap = None, name: str = "") -> Tensor:
    """multidim_stacker.py:124-134 on (b, C, T, h, w)."""
    sc = x
    x = F.silu(_bn(sd, p + "bn1.bn3d", F.conv3d(x, sd[p + "conv_pw.weight"]), REF_BN_EPS))
    x = _tap(tap, name + ".exp", x)
    x = F.conv3d(x, sd[p + "conv_dw.weight"], None, 1, 1, 1, cfg.mid_3d)
    x = F.silu(_bn(sd, p + "bn2.bn3d", x, REF_BN_EPS))
    x = _tap(tap, name + ".dw", x)
    s = x.mean((2, 3, 4), keepdim=True)
    s = F.silu(F.conv3d(s, sd[p + "se.conv_reduce.weight"], sd[p + "se.conv_reduce.bias"]))
    s = F.conv3d(s, sd[p + "se.conv_expand.weight"], sd[p + "se.conv_expand.bias"])
    x = x * torch.sigmoid(s)
    x = _tap(tap, name + ".se", x)
    x = _bn(sd, p + "bn3.bn3d", F.conv3d(x, sd[p + "conv_pwl.weight"]), REF_BN_EPS)
    return _tap(tap, name, x + sc)          # shortcut is unconditional (:133)


def forward_3d(sd: StateDict, x: Tensor, cfg: ModelConfig = ModelConfig(), tap: Tap = None) -> Tensor:
    """multidim_stacker.py:221-230.  (b, T, 192, h, w) -> (b, 256*T, h, w), channel = t*256 + c."""
    b, t, c, h, w = x.shape
    assert c == cfg.num_3d_features and t == cfg.num_stacks
    x = x.transpose(1, 2)
    for i in range(cfg.num_3d_blocks):
        x = block3d(sd, f"conv3d_encoder.{i}.", x, cfg, tap, f"c3d.{i}")
    x = x.transpose(1, 2).reshape(b * t, c, h, w)
    x = F.conv2d(x, sd["conv3d_projection.0.weight"])
    x = F.silu(_bn(sd, "conv3d_projection.1", x, REF_BN_EPS))
    x = _tap(tap, "proj3d", x)
    return x.reshape(b, cfg.num_features, h, w)


def gem(x: Tensor, p: Tensor, eps: float = 1e-6) -> Tensor:
    """multidim_stacker.py:42-45."""
    x = x.clamp(min=eps).pow(p)
    x = F.adaptive_avg_pool2d(x, 1).pow(1.0 / p)
    return x.view(x.size(0), -1)


def forward_head(sd: StateDict, x: Tensor, tap: Tap = None) -> Tensor:
    """multidim_stacker.py:232-237 in eval mode (dropout is a no-op)."""
    x = _tap(tap, "gem", gem(x, sd["global_pool.p"]))
    return F.linear(x, sd["classifier.weight"], sd["classifier.bias"])


def forward(sd: StateDict, x: Tensor, cfg: ModelConfig = ModelConfig(), tap: Tap = None) -> Tensor:
    """multidim_stacker.py:239-243.  (b, num_frames, H, W) f32 -> (b, num_classes) logits."""
    return forward_head(sd, forward_3d(sd, forward_2d(sd, x, cfg, tap), cfg, tap), tap)


# --------------------------------------------------------------------------------------------------------
# Frame pre-processing (src/frames.py)
# --------------------------------------------------------------------------------------------------------
def pad_normalize(frames_u8: Tensor, size: Tuple[int, int] = (1280, 736), fill_value: int = 0) -> Tensor:
    """frames.py:12-31 then :7-9.  size is (W, H).  uint8 (..., h, w) -> float32 (..., H, W) in [0, 1]."""
    h, w = frames_u8.shape[-2:]
    hp, wp = size[1] - h, size[0] - w
    assert hp >= 0 and wp >= 0
    top, left = hp // 2, wp // 2
    x = F.pad(frames_u8, [left, wp - left, top, hp - top], mode="constant", value=fill_value)
    return x.to(torch.float32) / 255.0


# --------------------------------------------------------------------------------------------------------
# Window index arithmetic (src/indexes.py:6-32) and the streaming predictor (src/predictors.py:20-75)
# --------------------------------------------------------------------------------------------------------
def stack_offsets(size: int, step: int) -> Tuple[int, int]:
    behind = (size // 2) * step
    ahead = (size - size // 2 - 1) * step
    return behind, ahead


def make_stack_indexes(frame_index: int, size: int, step: int) -> List[int]:
    behind, ahead = stack_offsets(size, step)
    return list(range(frame_index - behind, frame_index + ahead + 1, step))


def clip_index(index: int, frame_count: int, size: int, step: int, save_zone: int = 0) -> int:
    behind, ahead = stack_offsets(size, step)
    lo, hi = behind + save_zone, ahead + save_zone
    if index < lo:
        return lo
    if index >= frame_count - hi:
        return frame_count - hi - 1
    return index


class StreamingPredictorOracle:
    """Per-frame semantics of MultiDimStackerPredictor.predict (predictors.py:50-75), without the cache.

    The reference caches forward_2d per triple; results are identical to recomputing (eval-mode network is
    per-sample deterministic), so the oracle recomputes — it is the *checker*, not a fast path.
    """

    def __init__(self, sd: StateDict, cfg: ModelConfig = ModelConfig(), frame_stack_step: int = 2,
                 size: Tuple[int, int] = (1280, 736), tta: bool = False):
        self.sd, self.cfg, self.step, self.size, self.tta = sd, cfg, frame_stack_step, size, tta
        self.frames: Dict[int, Tensor] = {}
        self.offset = make_stack_indexes(0, cfg.num_frames, frame_stack_step)[-1]

    def reset_buffers(self):
        self.frames = {}

    def predict(self, frame_u8: Tensor, index: int):
        self.frames[index] = pad_normalize(frame_u8[None, None], self.size)[0, 0]
        p = index - self.offset
        idx = make_stack_indexes(p, self.cfg.num_frames, self.step)
        for k in [k for k in self.frames if k < idx[0]]:
            del self.frames[k]
        if not set(idx) <= set(self.frames):
            return None, p
        x = torch.stack([self.frames[i] for i in idx], 0)[None]
        if self.tta:
            x = torch.cat([x, x.flip(-1)], 0)          # kornia hflip == flip(-1) (predictors.py:63)
        with torch.no_grad():
            y = torch.sigmoid(forward(self.sd, x, self.cfg))   # prediction_transform (argus_models.py:26)
        return y.mean(0), p
