"""The four ``timm.layers`` names the reference imports (multidim_stacker.py:12-17) — test infrastructure."""
import functools
import math

import torch
from torch import nn
import torch.nn.functional as F


class DropPath(nn.Module):
    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask


class Conv2dSame(nn.Conv2d):
    """TensorFlow 'SAME' padding computed from the input size at run time."""

    def __init__(self, cin, cout, k, stride=1, dilation=1, groups=1, bias=False):
        super().__init__(cin, cout, k, stride, 0, dilation, groups, bias)

    def forward(self, x):
        ih, iw = x.shape[-2:]
        kh, kw = self.kernel_size
        sh, sw = self.stride
        ph = max((math.ceil(ih / sh) - 1) * sh + (kh - 1) * self.dilation[0] + 1 - ih, 0)
        pw = max((math.ceil(iw / sw) - 1) * sw + (kw - 1) * self.dilation[1] + 1 - iw, 0)
        if ph > 0 or pw > 0:
            x = F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
        return F.conv2d(x, self.weight, self.bias, self.stride, (0, 0), self.dilation, self.groups)


def create_conv2d(in_channels, out_channels, kernel_size, **kwargs):
    stride = kwargs.pop("stride", 1)
    padding = kwargs.pop("padding", "")
    depthwise = kwargs.pop("depthwise", False)
    bias = kwargs.pop("bias", False)
    groups = in_channels if depthwise else kwargs.pop("groups", 1)
    dynamic = isinstance(padding, str) and padding.lower() == "same" and stride != 1
    if dynamic:
        return Conv2dSame(in_channels, out_channels, kernel_size, stride=stride, groups=groups, bias=bias)
    # '' and static-'same' both resolve to symmetric (k-1)//2 padding for stride 1
    return nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=(kernel_size - 1) // 2,
                     groups=groups, bias=bias)


class BatchNormAct2d(nn.BatchNorm2d):
    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True,
                 apply_act=True, act_layer=nn.ReLU, act_kwargs=None, inplace=True, drop_layer=None, **_):
        super().__init__(num_features, eps=eps, momentum=momentum, affine=affine,
                         track_running_stats=track_running_stats)
        self.drop = nn.Identity()
        if apply_act and act_layer is not None:
            try:
                self.act = act_layer(inplace=inplace)
            except TypeError:
                self.act = act_layer()
        else:
            self.act = nn.Identity()

    def forward(self, x):
        return self.act(self.drop(super().forward(x)))


_ACTS = dict(silu=nn.SiLU, swish=nn.SiLU, relu=nn.ReLU, gelu=nn.GELU, sigmoid=nn.Sigmoid)


def get_act_layer(name="relu"):
    if name is None or not isinstance(name, str):
        return name
    return _ACTS[name.lower()]


def get_norm_act_layer(norm_layer, act_layer=None):
    assert norm_layer in (nn.BatchNorm2d, BatchNormAct2d, "batchnorm", "batchnorm2d")
    return functools.partial(BatchNormAct2d, act_layer=act_layer)
