"""Host mirror of ``src/indexes.py::StackIndexesGenerator`` (window index arithmetic, pure Python ints)."""
from __future__ import annotations


class StackIndexesGenerator:
    def __init__(self, size: int, step: int):
        self.size, self.step = size, step
        self.behind = (size // 2) * step
        self.ahead = (size - size // 2 - 1) * step

    def make_stack_indexes(self, frame_index: int) -> list:
        return list(range(frame_index - self.behind, frame_index + self.ahead + 1, self.step))

    def clip_index(self, index: int, frame_count: int, save_zone: int = 0) -> int:
        lo, hi = self.behind + save_zone, self.ahead + save_zone
        if index < lo:
            return lo
        if index >= frame_count - hi:
            return frame_count - hi - 1
        return index
