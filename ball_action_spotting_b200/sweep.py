"""Batched sliding-window sweep (SLIDING semantics of ``scripts/ball_action/predict.py::get_raw_predictions`` +
``src/predictors.py``) and its multi-GPU sharding.

The streaming predictor computes, for every new frame, ONE encoder pass (the new triple) and one 3D/head pass.  The
sweep does the same work for a whole buffer of frames at once (plus hop*(T-1) = 24 re-encoded triples per buffer of
`buffer_frames` frames, 1.2 % at the default 2048):
  * every triple start s in the buffer is one encoder image (frames s, s+step, s+2*step): the stem kernel addresses
    them in place with img_stride = 1 frame and plane_stride = `step` frames — no gather, no copy;
  * prediction p stacks the cached features of the triples starting at p-behind + 3*step*j, j = 0..T-1;
  * TTA = a second encoder pass with the stem reading mirrored columns, batch of 2 in the 3D/head pass, mean of the
    sigmoids (predictors.py:62-63,71-72).
Multi-GPU: the prediction range is split into contiguous shards, each rank needs a halo of
behind/ahead frames (28 for 15x2), and the only exchange is one all-gather of the (n, num_classes) probabilities.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch

from .indexes import StackIndexesGenerator


def prediction_bounds(gen: StackIndexesGenerator, frame_count: int, save_zone: int = 1) -> Tuple[int, int]:
    """First / last predict_index of a video (scripts/ball_action/predict.py:36-37), inclusive."""
    return gen.clip_index(0, frame_count, save_zone), gen.clip_index(frame_count, frame_count, save_zone)


def shard_range(lo: int, hi: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [a, b) of the inclusive prediction range [lo, hi]; sizes differ by at most one."""
    n = max(0, hi - lo + 1)
    base, extra = divmod(n, world)
    a = lo + rank * base + min(rank, extra)
    return a, a + base + (1 if rank < extra else 0)


def frames_needed(gen: StackIndexesGenerator, a: int, b: int) -> Tuple[int, int]:
    """Inclusive frame range a shard [a, b) of predictions reads (its own range plus the behind/ahead halo)."""
    return a - gen.behind, b - 1 + gen.ahead


def gather_predictions(local: torch.Tensor, counts: List[int], group=None) -> torch.Tensor:
    """The path's single collective: all-gather of per-rank (n_r, C) predictions, padded to the largest shard."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    width = max(counts)
    pad = torch.zeros((width, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * width, local.shape[1]), dtype=local.dtype, device=local.device)
    if local.is_cuda:
        dist.all_gather_into_tensor(out, pad, group=group)
    else:   # gloo (CPU tests)
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        out = torch.cat(parts, 0)
    return torch.cat([out[r * width: r * width + counts[r]] for r in range(world)], 0)


class SlidingSweep:
    """Batched equivalent of feeding frames one by one to MultiDimStackerPredictor.predict."""

    def __init__(self, nn_module, frame_stack_size: int = 15, frame_stack_step: int = 2, image_size=(1280, 736),
                 tta: bool = False, max_stacks: int = 64):
        self.module = nn_module
        self.gen = StackIndexesGenerator(frame_stack_size, frame_stack_step)
        self.step = frame_stack_step
        self.stack_size = nn_module.stack_size
        self.T = frame_stack_size // self.stack_size
        self.image_size = tuple(image_size)   # (W, H)
        self.tta = tta
        self.max_stacks = max_stacks

    @torch.no_grad()
    def predict_range(self, frames: torch.Tensor, first_frame: int, a: int, b: int) -> torch.Tensor:
        """frames: uint8 (n, h, W) on the GPU holding video frames first_frame .. first_frame+n-1.
        Returns float32 (b - a, num_classes): sigmoid probabilities for predict indices a .. b-1."""
        if b <= a:
            return torch.empty((0, self.module._cfg.num_classes), dtype=torch.float32, device=frames.device)
        f_lo, f_hi = frames_needed(self.gen, a, b)
        n, h, w = frames.shape
        if f_lo < first_frame or f_hi > first_frame + n - 1:
            raise RuntimeError(f"predictions [{a},{b}) need frames [{f_lo},{f_hi}], buffer holds [{first_frame},{first_frame + n - 1}]")
        if frames.dtype != torch.uint8 or not frames.is_contiguous():
            raise RuntimeError("frames must be a contiguous uint8 tensor")
        eng = self.module.engine(frames.device)
        W, H = self.image_size
        hop = self.step * self.stack_size                      # frames between consecutive triples of one window
        # every triple start needed by [a, b) is encoded exactly once (the engine walks them in chunk_images passes);
        # the 3D blocks + head then run over windows of the cached features, max_stacks predictions at a time
        s_lo = a - self.gen.behind                             # first triple start
        s_hi = (b - 1) - self.gen.behind + hop * (self.T - 1)
        n_img = s_hi - s_lo + 1
        feats = []
        for flip in ((False, True) if self.tta else (False,)):
            desc = eng.frames_desc(frames, H, W, h * w, self.step * h * w, hflip=flip, offset_elems=(s_lo - first_frame) * h * w)
            feats.append(eng.forward_2d(desc, n_img))          # (n_img, fh, fw, 192)
        out = []
        for c0 in range(a, b, self.max_stacks):
            c1 = min(b, c0 + self.max_stacks)
            probs = None
            for f in feats:     # window p = cached features of the triples starting at p - behind + hop * t (predictors.py:58-68)
                x = eng.gather_stacks(f, c0 - self.gen.behind - s_lo, hop, c1 - c0, self.T)    # (n_pred, T, fh, fw, 192)
                pr = eng.forward_head(eng.forward_3d(x), sigmoid=True)
                probs = pr if probs is None else eng.axpby_(probs, pr, 1.0, 1.0)
            if len(feats) > 1:
                eng.axpby_(probs, probs, 1.0 / len(feats), 0.0)                               # mean over the TTA branches (:72)
            out.append(probs)
        return torch.cat(out, 0)


def sweep_video(sweep: SlidingSweep, frame_source: Callable[[int, int], torch.Tensor], frame_count: int, save_zone: int = 1,
                rank: int = 0, world: int = 1, group=None, buffer_frames: int = 2048):
    """Predict a whole video, sharded over `world` ranks.  frame_source(i0, i1) returns uint8 frames [i0, i1) on this
    rank's GPU (decode / synthetic).  Returns (frame_indexes, predictions (N, C)) on every rank."""
    lo, hi = prediction_bounds(sweep.gen, frame_count, save_zone)
    a, b = shard_range(lo, hi, rank, world)
    parts = []
    step = max(1, buffer_frames - sweep.gen.behind - sweep.gen.ahead)
    for p0 in range(a, b, step):
        p1 = min(b, p0 + step)
        f_lo, f_hi = frames_needed(sweep.gen, p0, p1)
        parts.append(sweep.predict_range(frame_source(f_lo, f_hi + 1), f_lo, p0, p1))
    dev = parts[0].device if parts else torch.device("cuda", torch.cuda.current_device())
    local = torch.cat(parts, 0) if parts else torch.empty((0, sweep.module._cfg.num_classes), device=dev)
    if world > 1:
        counts = [shard_range(lo, hi, r, world)[1] - shard_range(lo, hi, r, world)[0] for r in range(world)]
        local = gather_predictions(local, counts, group)
    return list(range(lo, hi + 1)), local
