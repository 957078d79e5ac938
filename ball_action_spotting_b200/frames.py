"""Host mirror of ``src/frames.py`` (registry + PadNormalizeFramesProcessor).  In the B200 engine the padding and
the /255 normalisation are fused into the stem kernel, so the processor here is only needed by callers that
want the float frames themselves; it is implemented with torch tensor ops on whatever device the frames are on."""
from __future__ import annotations

import abc
from typing import Type

import torch


def normalize_frames(frames: torch.Tensor) -> torch.Tensor:
    return frames.to(torch.float32) / 255.0


def pad_to_frames(frames: torch.Tensor, size: tuple, pad_mode: str = "constant", fill_value: int = 0) -> torch.Tensor:
    h, w = frames.shape[-2:]
    hp, wp = size[1] - h, size[0] - w
    assert hp >= 0 and wp >= 0
    top, left = hp // 2, wp // 2
    return torch.nn.functional.pad(frames, [left, wp - left, top, hp - top], mode=pad_mode, value=fill_value)


class FramesProcessor(metaclass=abc.ABCMeta):
    @abc.abstractmethod
    def __call__(self, frames: torch.Tensor) -> torch.Tensor:
        ...


class PadNormalizeFramesProcessor(FramesProcessor):
    def __init__(self, size: tuple, pad_mode: str = "constant", fill_value: int = 0):
        self.size, self.pad_mode, self.fill_value = tuple(size), pad_mode, fill_value

    def __call__(self, frames: torch.Tensor) -> torch.Tensor:
        return normalize_frames(pad_to_frames(frames, self.size, self.pad_mode, self.fill_value))


_FRAME_PROCESSOR_REGISTRY: dict = dict(pad_normalize=PadNormalizeFramesProcessor)


def get_frames_processor(name: str, processor_params: dict) -> FramesProcessor:
    assert name in _FRAME_PROCESSOR_REGISTRY
    return _FRAME_PROCESSOR_REGISTRY[name](**processor_params)
