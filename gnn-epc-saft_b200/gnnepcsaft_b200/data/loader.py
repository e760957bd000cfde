"""Host -> device batch feed that overlaps the H2D copies with the previous step's kernels.

Stands where the reference has `batch.to(device)` inside the Lightning loop after the PyG
`DataLoader` (reference gnnepcsaft/train/train.py:74-79, SURVEY.md section 3.1: "H2D boundary,
synchronous, pageable"): batches arrive as host `Data` objects (PyG layout, int64), are pinned
if they are not already, and are copied on a side stream `depth` batches ahead of the consumer;
the compute stream only waits on the copy-done event of the batch it is about to use.
"""

from __future__ import annotations

from collections import deque
from typing import Iterable, Iterator

import torch


def _apply_to_tensors(batch, fn):
    """Map `fn` over the tensors of a batch object: PyG `Data.apply` (public), this package's containers (`_apply`),
    or a plain dict."""
    if isinstance(batch, dict):
        return {k: fn(v) if isinstance(v, torch.Tensor) else v for k, v in batch.items()}
    if hasattr(batch, "_apply"):
        return batch._apply(fn)
    if hasattr(batch, "apply"):
        return batch.apply(fn)
    raise TypeError(f"DevicePrefetcher: cannot map over the tensors of {type(batch).__name__}")


class DevicePrefetcher:
    """Iterate device batches; the copy of batch i+1.. runs while batch i is being consumed."""

    def __init__(self, batches: Iterable, device, depth: int = 2):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("gnnepcsaft_b200: DevicePrefetcher needs a CUDA device (no CPU path)")
        self.device, self.depth = dev, max(int(depth), 1)
        self.source = iter(batches)
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.queue: deque = deque()
        self.h2d_bytes = 0

    def _enqueue(self) -> bool:
        try:
            host = next(self.source)
        except StopIteration:
            return False

        def to_dev(t):
            if not t.is_cuda and not t.is_pinned():
                t = t.pin_memory()
            self.h2d_bytes += t.numel() * t.element_size()
            return t.to(self.device, non_blocking=True)

        with torch.cuda.stream(self.copy_stream):
            dev_batch = _apply_to_tensors(host, to_dev)
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        self.queue.append((dev_batch, done))
        return True

    def __iter__(self) -> Iterator:
        while len(self.queue) < self.depth and self._enqueue():
            pass
        while self.queue:
            dev_batch, done = self.queue.popleft()
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(done)
            # the tensors were allocated on the copy stream: tell the caching allocator they are used on `cur`
            _apply_to_tensors(dev_batch, lambda t: (t.record_stream(cur), t)[1])
            self._enqueue()
            yield dev_batch
