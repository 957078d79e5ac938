"""Algorithmic bytes / FLOPs of every kernel launch of the path (fp16 activations; weights excluded — 13.5 MB,
amortised over the batch).  "Algorithmic" = each kernel reads its input tensor(s) once and writes its output once.
Used by bench.py for the roofline figures and quoted in DESIGN.md."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List

from .packer import STAGES

KINDS = {0: "stem", 1: "conv3x3", 2: "gemm1x1", 3: "dwconv2d", 4: "dwconv3d", 5: "se_fc", 6: "head", 11: "tail2d", 12: "tail3d"}


@dataclass
class Launch:
    kind: int
    tag: int
    name: str
    bytes: float
    flops: float


def encoder_launches(H: int, W: int, stored_h: int, in_elem_bytes: int = 1, tail_mode: int = 2) -> List[Launch]:
    """Launch list for ONE image through stem + 22 blocks + conv2d_projection.  tail_mode (mds_set_tail_mode): 2 = depthwise+SE
    kernel and gating GEMM, 0 = depthwise, SE kernel, pre-gated GEMM, 1 = the MBConv tail is ONE launch that reads the expanded
    tensor once, writes the block output and reads the shortcut (its depthwise output makes a round trip through L2 that is
    NOT counted as algorithmic bytes)."""
    out: List[Launch] = []
    h, w = H // 2, W // 2
    out.append(Launch(0, 0, "stem", 3 * stored_h * W * in_elem_bytes + h * w * 32 * 2, 2.0 * h * w * 27 * 32))
    cin, tag = 32, 0
    for si, (kind, reps, stride, expand, cout, has_se) in enumerate(STAGES):
        for r in range(reps):
            tag += 1
            s = stride if r == 0 else 1
            mid, ho, wo = cin * expand, h // s, w // s
            nm = f"b{si}.{r}"
            if kind == "cn":
                out.append(Launch(1, tag, nm + ".c3", (h * w * cin + ho * wo * cout) * 2, 2.0 * ho * wo * 9 * cin * cout))
            elif kind == "er":
                out.append(Launch(1, tag, nm + ".c3+pwl", (h * w * cin + ho * wo * cout) * 2,
                                  2.0 * ho * wo * (9 * cin * mid + mid * cout)))
            else:
                skip = s == 1 and cin == cout
                out.append(Launch(2, tag, nm + ".pw", h * w * (cin + mid) * 2, 2.0 * h * w * cin * mid))
                if tail_mode == 1:
                    out.append(Launch(11, tag, nm + ".tail", (h * w * mid + ho * wo * (cout + (cout if skip else 0))) * 2,
                                      2.0 * ho * wo * (9 * mid + mid * cout)))
                else:
                    out.append(Launch(3, tag, nm + ".dw", (h * w + ho * wo) * mid * 2, 2.0 * ho * wo * 9 * mid))
                    if tail_mode == 0:
                        out.append(Launch(5, tag, nm + ".se", 0, 0))
                    out.append(Launch(2, tag, nm + ".pwl", ho * wo * (mid + cout + (cout if skip else 0)) * 2, 2.0 * ho * wo * mid * cout))
            cin, h, w = cout, ho, wo
    tag += 1
    out.append(Launch(2, tag, "proj2d", h * w * (192 + 192) * 2, 2.0 * h * w * 192 * 192))
    return out


def stack3d_launches(fh: int, fw: int, T: int, c3: int = 192, mid: int = 576, proj: int = 256, blocks: int = 4,
                     tail_mode: int = 2) -> List[Launch]:
    """Launch list for ONE stack through the 3D blocks + conv3d_projection + head."""
    out: List[Launch] = []
    rows = T * fh * fw
    for i in range(blocks):
        tag = 101 + i
        out.append(Launch(2, tag, f"c3d.{i}.pw", rows * (c3 + mid) * 2, 2.0 * rows * c3 * mid))
        if tail_mode == 1:
            out.append(Launch(12, tag, f"c3d.{i}.tail", rows * (mid + 2 * c3) * 2, 2.0 * rows * (27 * mid + mid * c3)))
        else:
            out.append(Launch(4, tag, f"c3d.{i}.dw", rows * mid * 2 * 2, 2.0 * rows * 27 * mid))
            if tail_mode == 0:
                out.append(Launch(5, tag, f"c3d.{i}.se", 0, 0))
            out.append(Launch(2, tag, f"c3d.{i}.pwl", rows * (mid + 2 * c3) * 2, 2.0 * rows * mid * c3))
    out.append(Launch(2, 150, "proj3d", rows * (c3 + proj) * 2, 2.0 * rows * c3 * proj))
    out.append(Launch(6, 200, "gem+linear", rows * proj * 2, 4.0 * rows * proj))
    return out


def per_stack_totals(H: int = 736, W: int = 1280, stored_h: int = 720, T: int = 5, tail_mode: int = 2) -> Dict[str, Dict[str, float]]:
    """bytes / flops per frame-stack, by kernel kind, for the FULL forward (T encoder passes per stack)."""
    tot: Dict[str, Dict[str, float]] = {}
    for l in encoder_launches(H, W, stored_h, tail_mode=tail_mode):
        d = tot.setdefault(KINDS[l.kind], {"bytes": 0.0, "flops": 0.0, "launches": 0})
        d["bytes"] += l.bytes * T
        d["flops"] += l.flops * T
        d["launches"] += 1
    for l in stack3d_launches(H // 32, W // 32, T, tail_mode=tail_mode):
        d = tot.setdefault(KINDS[l.kind], {"bytes": 0.0, "flops": 0.0, "launches": 0})
        d["bytes"] += l.bytes
        d["flops"] += l.flops
        d["launches"] += 1
    return tot


def depthwise_only_bytes(H: int = 736, W: int = 1280, T: int = 5) -> Dict[str, float]:
    """SURVEY.md 8(d) "depthwise-stage-only" bytes per frame-stack (the north_star's >= 60 % figure): every depthwise layer
    reads mid x H x W and writes mid x Ho x Wo fp16 (2D: 16 layers x T images; 3D: 4 layers)."""
    d2 = sum(l.bytes for l in encoder_launches(H, W, 720, tail_mode=0) if l.kind == 3) * T
    d3 = sum(l.bytes for l in stack3d_launches(H // 32, W // 32, T, tail_mode=0) if l.kind == 4)
    return {"dwconv2d": d2, "dwconv3d": d3}
