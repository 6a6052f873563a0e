"""CPU ORACLE (test infrastructure): `RingBuffer[T]`, the semantic model of the pending-task ring.

Restates pkg/abstractions/common/ring_buffer.go:8-96. PINNED by
pkg/abstractions/common/ring_buffer_test.go:7-169.
"""
from __future__ import annotations

from typing import Any, List, Tuple


class RingBuffer:
    def __init__(self, size: int):
        self.buffer: List[Any] = [None] * size
        self.size = size
        self.head = 0
        self.tail = 0
        self.count = 0
        self.overwrites = 0

    def push(self, item: Any, priority: bool = False) -> bool:
        overwritten = False
        if priority:
            if self.count == self.size:
                self.buffer[self.head] = item
                overwritten = True
            else:
                self.head = (self.head - 1 + self.size) % self.size
                self.buffer[self.head] = item
                self.count += 1
        else:
            self.buffer[self.tail] = item
            self.tail = (self.tail + 1) % self.size
            if self.count == self.size:
                self.head = (self.head + 1) % self.size
                overwritten = True
            else:
                self.count += 1
        if overwritten:
            self.overwrites += 1
        return overwritten

    def pop(self) -> Tuple[Any, bool]:
        if self.count == 0:
            return None, False
        item = self.buffer[self.head]
        self.buffer[self.head] = None
        self.head = (self.head + 1) % self.size
        self.count -= 1
        return item, True

    def __len__(self) -> int:
        return self.count

    def capacity(self) -> int:
        return self.size
