"""Window index arithmetic (integer, bit-exact) — same API as /root/reference/src/indexes.py:1-39."""
from __future__ import annotations


class IndexesGenerator:
    def __init__(self, size: int, step: int, position: str = "last"):
        self.size, self.step = size, step
        spans = {"first": (0, size - 1), "middle": (size // 2, size - size // 2 - 1), "last": (size - 1, 0)}
        if position not in spans:
            raise ValueError(f"Index position value should be one of {'first', 'middle', 'last'}")
        behind, ahead = spans[position]
        self.behind = behind * step
        self.ahead = ahead * step
        self.width = self.behind + self.ahead + 1

    def make_indexes(self, index: int) -> list[int]:
        return list(range(index - self.behind, index + self.ahead + 1, self.step))

    def clip_index(self, index: int, length: int, save_zone: int = 0) -> int:
        lo = self.behind + save_zone
        hi = length - (self.ahead + save_zone) - 1
        if index < lo:
            return lo
        if index > hi:
            return hi
        return index
