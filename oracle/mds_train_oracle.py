"""CPU fp32 oracle for the frozen-encoder training step (BASELINE.json configs[4]).  TEST INFRASTRUCTURE ONLY.

Only ``tests/`` and benchmark baselines may import this file; the product package never does.

What it restates (citations relative to ``/root/reference``):

* ``src/argus_models.py:41-74``  ``BallActionModel.train_step``: train mode, zero_grad, forward, loss, backward,
  optimizer step (``iter_size = 1``; the GradScaler only changes *when* a step is skipped, restated in
  ``GradScalerOracle``).
* ``src/argus_models.py:104-110`` ``freeze_conv2d_encoder``: the encoder parameters get no gradient.  The step
  restated here starts at the encoder *output* (b*T, 192, h, w): ``conv2d_projection`` -> ``forward_3d`` ->
  ``forward_head`` (``src/models/multidim_stacker.py:216-237``) with train-mode BatchNorm (batch statistics +
  running-stat update, momentum 0.1, unbiased running variance), DropPath on the residual branch (:133, timm
  ``drop_path`` with ``scale_by_keep``), Dropout on the pooled features (:234-235), learnable GeM ``p`` (:38).
* ``src/losses.py:6-50``         ``sigmoid_focal_loss`` (alpha, gamma, mean reduction).
* ``configs/ball_action/ball_finetune_long_004.py:51-55`` SGD, momentum 0.9, Nesterov (torch.optim.SGD update rule).

The random masks are *inputs* (``dp_masks[block][sample]`` in {0, 1/keep}, ``dropout_mask[sample][feature]`` in
{0, 1/(1-p)}) so that the CUDA path and the oracle see the same draw.  ``oracle/make_train_golden.py`` pins this
file against the unmodified reference module (same seed -> same masks) and torch.optim.SGD.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import mds_oracle as O

Tensor = torch.Tensor
BN_MOMENTUM = 0.1          # nn.BatchNorm2d / nn.BatchNorm3d default


def trainable_keys(cfg: O.ModelConfig) -> List[str]:
    """Parameters that receive a gradient when ``freeze_conv2d_encoder`` is set (argus_models.py:104-110)."""
    keys = ["conv2d_projection.0.weight", "conv2d_projection.1.weight", "conv2d_projection.1.bias"]
    for i in range(cfg.num_3d_blocks):
        p = f"conv3d_encoder.{i}."
        keys += [p + "conv_pw.weight", p + "bn1.bn3d.weight", p + "bn1.bn3d.bias",
                 p + "conv_dw.weight", p + "bn2.bn3d.weight", p + "bn2.bn3d.bias",
                 p + "se.conv_reduce.weight", p + "se.conv_reduce.bias",
                 p + "se.conv_expand.weight", p + "se.conv_expand.bias",
                 p + "conv_pwl.weight", p + "bn3.bn3d.weight", p + "bn3.bn3d.bias"]
    keys += ["conv3d_projection.0.weight", "conv3d_projection.1.weight", "conv3d_projection.1.bias",
             "global_pool.p", "classifier.weight", "classifier.bias"]
    return keys


def bn_prefixes(cfg: O.ModelConfig) -> List[str]:
    out = ["conv2d_projection.1"]
    for i in range(cfg.num_3d_blocks):
        out += [f"conv3d_encoder.{i}.bn{j}.bn3d" for j in (1, 2, 3)]
    return out + ["conv3d_projection.1"]


def _bn_train(sd: Dict[str, Tensor], prefix: str, x: Tensor, new_stats: Optional[Dict[str, Tensor]]) -> Tensor:
    """nn.BatchNorm in training mode: normalise with the biased batch variance; the running statistics move by
    momentum 0.1 towards (batch mean, *unbiased* batch variance)."""
    kw, kb, km, kv, kn = O._bn_keys(prefix)
    rm, rv = sd[km].clone(), sd[kv].clone()
    y = F.batch_norm(x, rm, rv, sd[kw], sd[kb], training=True, momentum=BN_MOMENTUM, eps=O.REF_BN_EPS)
    if new_stats is not None:
        new_stats[km], new_stats[kv] = rm, rv
        new_stats[kn] = sd[kn] + 1
    return y


def sigmoid_focal_loss(inputs: Tensor, targets: Tensor, alpha: float, gamma: float) -> Tensor:
    """src/losses.py:31-48, reduction='mean'."""
    inputs, targets = inputs.float(), targets.float()
    p = torch.sigmoid(inputs)
    ce = F.binary_cross_entropy_with_logits(inputs, targets, reduction="none")
    p_t = p * targets + (1 - p) * (1 - targets)
    loss = ce * ((1 - p_t) ** gamma)
    if alpha >= 0:
        loss = (alpha * targets + (1 - alpha) * (1 - targets)) * loss
    return loss.mean()


def train_forward(sd: Dict[str, Tensor], enc_feats: Tensor, cfg: O.ModelConfig, dp_masks: Tensor,
                  dropout_mask: Tensor, new_stats: Optional[Dict[str, Tensor]] = None,
                  taps: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """enc_feats (b*T, 192, h, w) f32 = frozen encoder output -> logits (b, classes), train mode.

    multidim_stacker.py:216 (conv2d_projection), :221-230 (forward_3d), :232-237 (forward_head)."""
    n, c, h, w = enc_feats.shape
    T = cfg.num_stacks
    b = n // T

    def tap(name, t):
        if taps is not None:
            taps[name] = t
            if t.requires_grad:
                t.retain_grad()
        return t

    x = F.conv2d(enc_feats, sd["conv2d_projection.0.weight"])
    x = F.silu(_bn_train(sd, "conv2d_projection.1", x, new_stats))
    x = x.contiguous().view(b, T, cfg.num_3d_features, h, w).transpose(1, 2)      # (b, C, T, h, w)
    x = tap("x0", x)
    for i in range(cfg.num_3d_blocks):
        p = f"conv3d_encoder.{i}."
        sc = x
        y = F.conv3d(x, sd[p + "conv_pw.weight"])
        y = F.silu(_bn_train(sd, p + "bn1.bn3d", y, new_stats))
        y = tap(f"b{i}.a1", y)
        y = F.conv3d(y, sd[p + "conv_dw.weight"], None, 1, 1, 1, cfg.mid_3d)
        y = tap(f"b{i}.y2", y)
        y = F.silu(_bn_train(sd, p + "bn2.bn3d", y, new_stats))
        s = y.mean((2, 3, 4), keepdim=True)
        s = F.silu(F.conv3d(s, sd[p + "se.conv_reduce.weight"], sd[p + "se.conv_reduce.bias"]))
        s = F.conv3d(s, sd[p + "se.conv_expand.weight"], sd[p + "se.conv_expand.bias"])
        y = y * torch.sigmoid(s)
        y = tap(f"b{i}.a2g", y)
        y = F.conv3d(y, sd[p + "conv_pwl.weight"])
        y = _bn_train(sd, p + "bn3.bn3d", y, new_stats)
        x = y * dp_masks[i].view(b, 1, 1, 1, 1) + sc                              # drop_path(x) + shortcut (:133)
        x = tap(f"x{i + 1}", x)
    x = x.transpose(1, 2).reshape(b * T, cfg.num_3d_features, h, w)
    x = F.conv2d(x, sd["conv3d_projection.0.weight"])
    x = F.silu(_bn_train(sd, "conv3d_projection.1", x, new_stats))
    x = tap("proj3d", x)
    x = x.view(b, cfg.num_features, h, w)
    feat = tap("feat", O.gem(x, sd["global_pool.p"]))
    feat = feat * dropout_mask                                                    # F.dropout(p, training=True)
    return F.linear(feat, sd["classifier.weight"], sd["classifier.bias"])


def loss_and_grads(sd: Dict[str, Tensor], enc_feats: Tensor, targets: Tensor, cfg: O.ModelConfig, dp_masks: Tensor,
                   dropout_mask: Tensor, alpha: float = 0.4, gamma: float = 1.2
                   ) -> Tuple[Tensor, Tensor, Dict[str, Tensor], Dict[str, Tensor]]:
    """-> (loss, logits, grads by state-dict key, updated BN buffers)."""
    keys = trainable_keys(cfg)
    work = {k: v.detach().clone() for k, v in sd.items()}
    for k in keys:
        work[k].requires_grad_(True)
    new_stats: Dict[str, Tensor] = {}
    logits = train_forward(work, enc_feats, cfg, dp_masks, dropout_mask, new_stats)
    loss = sigmoid_focal_loss(logits, targets, alpha, gamma)
    grads = torch.autograd.grad(loss, [work[k] for k in keys])
    return loss.detach(), logits.detach(), dict(zip(keys, grads)), new_stats


def sgd_nesterov_step(params: Dict[str, Tensor], grads: Dict[str, Tensor], bufs: Dict[str, Tensor], lr: float,
                      momentum: float = 0.9) -> None:
    """torch.optim.SGD(momentum, nesterov=True, dampening=0, weight_decay=0), in place: the first step seeds the
    momentum buffer with the gradient."""
    for k, g in grads.items():
        if k not in bufs:
            bufs[k] = g.clone()
        else:
            bufs[k].mul_(momentum).add_(g)
        params[k] = params[k] - lr * (g + momentum * bufs[k])


class GradScalerOracle:
    """torch.cuda.amp.GradScaler defaults (argus_models.py:36,65-66): scale 65536, growth 2x every 2000 clean steps,
    backoff 0.5 and a skipped optimizer step when any gradient is non-finite."""

    def __init__(self, init_scale: float = 65536.0, growth_factor: float = 2.0, backoff_factor: float = 0.5,
                 growth_interval: int = 2000):
        self.scale, self.growth_factor, self.backoff_factor = init_scale, growth_factor, backoff_factor
        self.growth_interval, self.tracker = growth_interval, 0

    def update(self, found_inf: bool) -> None:
        if found_inf:
            self.scale *= self.backoff_factor
            self.tracker = 0
        else:
            self.tracker += 1
            if self.tracker == self.growth_interval:
                self.scale *= self.growth_factor
                self.tracker = 0


def make_masks(cfg: O.ModelConfig, b: int, drop_path_rate: float, drop_rate: float, seed: int) -> Tuple[Tensor, Tensor]:
    g = torch.Generator().manual_seed(seed)
    keep = 1.0 - drop_path_rate
    dp = torch.bernoulli(torch.full((cfg.num_3d_blocks, b), keep), generator=g) / keep
    do = torch.bernoulli(torch.full((b, cfg.num_features), 1.0 - drop_rate), generator=g) / (1.0 - drop_rate)
    return dp, do


def make_case(cfg: O.ModelConfig, b: int, hw: Tuple[int, int], seed: int = 7) -> Tuple[Tensor, Tensor]:
    """Seeded synthetic encoder features (b*T, 192, h, w) and soft multilabel targets (b, classes)."""
    g = torch.Generator().manual_seed(seed)
    enc = torch.randn((b * cfg.num_stacks, 192, *hw), generator=g) * 0.7 + 0.1
    hard = (torch.rand((b, cfg.num_classes), generator=g) > 0.6).float()
    return enc, hard * torch.rand((b, cfg.num_classes), generator=g)
