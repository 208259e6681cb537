"""Logging setup and the step latency tracker (reference torchstore/logging.py:13-54).

Step names used on the hot path match the reference's ("create transport buffer",
"put_to_storage_volume", "notify_put_batch", "fetch", ...), so DEBUG timelines line up.
"""

from __future__ import annotations

import logging
import os
import sys
import time


def init_logging() -> None:
    level = os.environ.get("TORCHSTORE_LOG_LEVEL", "INFO").upper()
    root = logging.getLogger()
    root.setLevel(level)
    for handler in root.handlers:
        if isinstance(handler, logging.StreamHandler) and getattr(handler, "stream", None) is sys.stdout:
            return
    handler = logging.StreamHandler(sys.stdout)
    handler.setLevel(level)
    root.addHandler(handler)


class LatencyTracker:
    """perf_counter deltas per named step, logged at DEBUG; GB/s when given a tensor or a byte count."""

    def __init__(self, name: str) -> None:
        self.name = name
        self.start_time = self.last_step = time.perf_counter()

    @staticmethod
    def _throughput(elapsed: float, tensor=None, nbytes: int | None = None) -> str:
        if nbytes is None and tensor is not None:
            nbytes = tensor.numel() * tensor.element_size()
        if not nbytes or elapsed <= 0:
            return ""
        return f" ({nbytes / 1e9 / elapsed:.2f} GB/s)"

    def track_step(self, step_name: str, tensor=None, nbytes: int | None = None) -> None:
        now = time.perf_counter()
        elapsed = now - self.last_step
        logging.debug("%s:%s took %.4fs%s", self.name, step_name, elapsed, self._throughput(elapsed, tensor, nbytes))
        self.last_step = now

    def track_e2e(self) -> None:
        logging.debug("%s took %s seconds", self.name, time.perf_counter() - self.start_time)
