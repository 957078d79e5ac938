"""Drop-in for ``src/models/multidim_stacker.py::MultiDimStacker`` (boundary B1).

Same constructor kwargs (multidim_stacker.py:138-153), same attributes (:156-161), same ``state_dict()`` keys and
shapes (so ``argus.load_model`` / ``load_state_dict`` / ``load_weights_from_pretrain`` work unchanged), same
``forward / forward_2d / forward_3d / forward_head`` shapes and dtypes at the edge.  Internally nothing is
computed by PyTorch: the submodules below only *hold parameters*; the forward passes call the sm_100a kernels
through the C-ABI (``engine.Engine``).  Inference only (eval mode), CUDA only — no fallback.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .engine import Engine, EngineConfig
from .packer import STAGES, pack_state_dict


class _Holder(nn.Module):
    """Parameter container; never called."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder: compute happens in libmds_b200")


def _conv2d(cin, cout, k, groups=1, bias=False):
    return nn.Conv2d(cin, cout, k, groups=groups, bias=bias)


def _bn2d(c, eps):
    return nn.BatchNorm2d(c, eps=eps)


def _encoder(in_chans: int) -> nn.Module:
    enc = _Holder()
    enc.conv_stem = _conv2d(in_chans, 32, 3)
    enc.bn1 = _bn2d(32, 1e-3)
    stages, cin = [], 32
    for kind, reps, stride, expand, cout, has_se in STAGES:
        blocks = []
        for r in range(reps):
            b, mid = _Holder(), cin * expand
            if kind == "cn":
                b.conv, b.bn1 = _conv2d(cin, cout, 3), _bn2d(cout, 1e-3)
            elif kind == "er":
                b.conv_exp, b.bn1 = _conv2d(cin, mid, 3), _bn2d(mid, 1e-3)
                b.conv_pwl, b.bn2 = _conv2d(mid, cout, 1), _bn2d(cout, 1e-3)
            else:
                b.conv_pw, b.bn1 = _conv2d(cin, mid, 1), _bn2d(mid, 1e-3)
                b.conv_dw, b.bn2 = _conv2d(mid, mid, 3, groups=mid), _bn2d(mid, 1e-3)
                b.se = _Holder()
                rd = int(round(cin * 0.25))
                b.se.conv_reduce, b.se.conv_expand = _conv2d(mid, rd, 1, bias=True), _conv2d(rd, mid, 1, bias=True)
                b.conv_pwl, b.bn3 = _conv2d(mid, cout, 1), _bn2d(cout, 1e-3)
            blocks.append(b)
            cin = cout
        stages.append(nn.Sequential(*blocks))
    enc.blocks = nn.Sequential(*stages)
    return enc


def _bn3d(c):
    h = _Holder()
    h.bn3d = nn.BatchNorm3d(c)
    return h


def _block3d(c, mid, rd):
    b = _Holder()
    b.conv_pw, b.bn1 = nn.Conv3d(c, mid, 1, bias=False), _bn3d(mid)
    b.conv_dw, b.bn2 = nn.Conv3d(mid, mid, 3, padding=1, groups=mid, bias=False), _bn3d(mid)
    b.se = _Holder()
    b.se.conv_reduce, b.se.conv_expand = nn.Conv3d(mid, rd, 1, bias=True), nn.Conv3d(rd, mid, 1, bias=True)
    b.conv_pwl, b.bn3 = nn.Conv3d(mid, c, 1, bias=False), _bn3d(c)
    return b


class _GeMParams(_Holder):
    def __init__(self, norm: float):
        super().__init__()
        self.p = nn.Parameter(torch.ones(1) * norm)      # multidim_stacker.py:38


class MultiDimStacker(nn.Module):
    def __init__(self, model_name: str, num_classes: int, num_frames: int = 15, stack_size: int = 3,
                 index_2d_features: int = 4, pretrained: bool = False, num_3d_blocks: int = 2,
                 num_3d_features: int = 192, num_3d_stack_proj: int = 256, expansion_3d_ratio: int = 6,
                 se_reduce_3d_ratio: int = 24, drop_rate: float = 0., drop_path_rate: float = 0.,
                 act_layer: str = "silu", chunk_images: int = 0, bias_correction: bool = True, **kwargs):
        super().__init__()
        assert num_frames > 0 and num_frames % stack_size == 0            # multidim_stacker.py:155
        if model_name.split(".")[0] != "tf_efficientnetv2_b0":
            raise NotImplementedError(f"only tf_efficientnetv2_b0 is built for B200, got {model_name}")
        if index_2d_features != 4 or act_layer != "silu" or stack_size != 3:
            raise NotImplementedError("only index_2d_features=4, act_layer='silu', stack_size=3 are built")
        if pretrained:
            raise RuntimeError("pretrained=True needs network access; load a checkpoint instead")
        self.num_frames, self.stack_size = num_frames, stack_size
        self.num_3d_features = num_3d_features
        self.num_stacks = num_frames // stack_size
        self.num_features = num_3d_stack_proj * self.num_stacks
        self.drop_rate = drop_rate
        self.drop_path_rate = drop_path_rate      # used by train.FrozenEncoderTrainer (DropPath of the 3D blocks, :122)
        mid = num_3d_features * expansion_3d_ratio
        self._cfg = EngineConfig(num_classes, num_frames, stack_size, num_3d_blocks, num_3d_features, num_3d_stack_proj,
                                 expansion_3d_ratio, se_reduce_3d_ratio, chunk_images)

        self.conv2d_encoder = _encoder(stack_size)
        self.conv2d_projection = nn.Sequential(_conv2d(192, num_3d_features, 1), _bn2d(num_3d_features, 1e-5))
        self.conv3d_encoder = nn.Sequential(*[_block3d(num_3d_features, mid, mid // se_reduce_3d_ratio)
                                              for _ in range(num_3d_blocks)])
        self.conv3d_projection = nn.Sequential(_conv2d(num_3d_features, num_3d_stack_proj, 1), _bn2d(num_3d_stack_proj, 1e-5))
        self.global_pool = _GeMParams(3.0)
        self.classifier = nn.Linear(self.num_features, num_classes, bias=True)
        self._bias_correction = bias_correction      # packer.py: data-free correction of the fp16 weight-rounding bias
        self._engine: Optional[Engine] = None
        self._dirty = True

    def init_random_(self, seed: int = 0) -> "MultiDimStacker":
        """Seeded random weights for benchmarks without a checkpoint: conv ~ N(0, sqrt(2/fan_out)) (timm's
        efficientnet init), BatchNorm affine and running statistics perturbed around identity."""
        import math
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for m in self.modules():
                if isinstance(m, (nn.Conv2d, nn.Conv3d)):
                    fan_out = m.out_channels * math.prod(m.kernel_size) // m.groups
                    m.weight.copy_(torch.randn(m.weight.shape, generator=g) * math.sqrt(2.0 / fan_out))
                    if m.bias is not None:
                        m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
                elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm3d)):
                    m.weight.copy_(torch.rand(m.weight.shape, generator=g) * 0.5 + 0.5)
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
                    m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                    m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
                elif isinstance(m, nn.Linear):
                    m.weight.copy_(torch.randn(m.weight.shape, generator=g) * 0.05)
                    m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
        self._dirty = True
        return self

    # ---- weight tracking -------------------------------------------------------------------------------------
    def load_state_dict(self, *a, **k):
        r = super().load_state_dict(*a, **k)
        self._dirty = True
        return r

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._dirty = True
        return r

    def repack(self) -> None:
        """Call after mutating parameters in place (load_state_dict / .to() are tracked automatically)."""
        self._dirty = True

    def engine(self, device: Optional[torch.device] = None) -> Engine:
        dev = torch.device(device) if device is not None else self.classifier.weight.device
        if dev.type != "cuda":
            raise RuntimeError("MultiDimStacker (B200): parameters must live on a CUDA device; there is no CPU path")
        if self.training:
            raise RuntimeError("MultiDimStacker (B200) is an inference engine: call .eval() first")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        if self._engine is not None and self._engine.device.index != idx:
            self._engine.close()
            self._engine = None
        if self._engine is None:
            self._engine = Engine(self._cfg, pack_state_dict(self.state_dict(), self._cfg.num_3d_blocks, self._bias_correction), dev)
            self._dirty = False
        elif self._dirty:
            self._engine.load_packed(pack_state_dict(self.state_dict(), self._cfg.num_3d_blocks, self._bias_correction))
            self._dirty = False
        return self._engine

    # ---- reference API ---------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_2d(self, x: torch.Tensor) -> torch.Tensor:
        b, t, h, w = x.shape
        assert t % self.stack_size == 0                                   # multidim_stacker.py:212
        k = t // self.stack_size
        eng = self.engine(x.device)
        x = x.contiguous()
        desc = eng.frames_desc(x, self._padded_h(x, h), w, self.stack_size * h * w, h * w)
        feats = eng.forward_2d(desc, b * k)                               # (b*k, fh, fw, 192) fp16
        out = eng.nhwc16_to_nchw32(feats)                                 # (b*k, 192, fh, fw) f32
        return out.view(b, k, self.num_3d_features, out.shape[-2], out.shape[-1])

    @torch.no_grad()
    def forward_3d(self, x: torch.Tensor) -> torch.Tensor:
        b, t, c, h, w = x.shape
        assert c == self.num_3d_features and t == self.num_stacks        # multidim_stacker.py:223
        eng = self.engine(x.device)
        feats = eng.nchw32_to_nhwc16(x.reshape(b * t, c, h, w).float()).view(b, t, h, w, c)
        out = eng.forward_3d(feats)                                       # (b, T, h, w, proj) fp16
        out = eng.nhwc16_to_nchw32(out.view(b * t, h, w, -1))             # (b*T, proj, h, w)
        return out.view(b, self.num_features, h, w)                       # channel = t*proj + c (:229)

    @torch.no_grad()
    def forward_head(self, x: torch.Tensor) -> torch.Tensor:
        b, f, h, w = x.shape
        assert f == self.num_features
        eng = self.engine(x.device)
        pj = f // self.num_stacks
        xh = eng.nchw32_to_nhwc16(x.reshape(b * self.num_stacks, pj, h, w).float()).view(b, self.num_stacks, h, w, pj)
        return eng.forward_head(xh)

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        b, t, h, w = x.shape
        assert t == self.num_frames, f"expected {self.num_frames} frames, got {t}"
        eng = self.engine(x.device)
        x = x.contiguous()
        desc = eng.frames_desc(x, self._padded_h(x, h), w, self.stack_size * h * w, h * w)
        return eng.forward(desc, b)

    @staticmethod
    def _padded_h(x: torch.Tensor, h: int) -> int:
        # float input is already padded (frames.py ran upstream); raw uint8 frames are padded to a multiple of 32
        return h if x.dtype != torch.uint8 else (h + 31) // 32 * 32
