"""Host -> device staging of observation batches for `select_action`.

The reference moves every observation with a blocking `.to(device)` (lerobot_fastvla/processor_fastvla.py:33,
fastvlm_adapter.py:488) and then bounces the frames through the CPU once more (:485).  Here a batch that arrives in
host memory is copied on a SIDE stream into one of two device staging slots and the forward is ordered after that copy
with an event: the call returns as soon as the work is queued, so the copy of observation i+1 overlaps the forward of
observation i whenever the caller keeps a step in flight (a serving loop reading actions back one step late, or the
environment stepping on the host).  Pinned host tensors make the copy truly asynchronous; pageable ones still work
(the driver stages them synchronously).  uint8 frames stay uint8 until the ingest kernel scales them.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch


class ObservationStager:
    SLOTS = 2

    def __init__(self, device: torch.device) -> None:
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._slots: List[Dict[str, torch.Tensor]] = [dict() for _ in range(self.SLOTS)]
        self._copied = [torch.cuda.Event() for _ in range(self.SLOTS)]
        self._consumed: List[Optional[torch.cuda.Event]] = [None] * self.SLOTS
        self._next = 0
        self.h2d_bytes = 0  # bytes of the last staged batch (what the e2e benchmark reports)

    def _buffer(self, slot: int, name: str, like: torch.Tensor) -> torch.Tensor:
        buf = self._slots[slot].get(name)
        if buf is None or buf.shape != like.shape or buf.dtype != like.dtype:
            buf = torch.empty(like.shape, dtype=like.dtype, device=self.device)
            self._slots[slot][name] = buf
        return buf

    def stage(self, tensors: Dict[str, torch.Tensor]) -> Tuple[Dict[str, torch.Tensor], int]:
        """Copy the host tensors of `tensors` to the device (device tensors pass through).  Returns the device views
        and the slot index to hand to `release()` once the consuming kernels are queued."""
        slot = self._next
        self._next = (self._next + 1) % self.SLOTS
        cur = torch.cuda.current_stream(self.device)
        out: Dict[str, torch.Tensor] = {}
        host = {k: v for k, v in tensors.items() if not v.is_cuda}
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
        if host:
            with torch.cuda.stream(self.copy_stream):
                if self._consumed[slot] is not None:
                    self.copy_stream.wait_event(self._consumed[slot])  # the forward that read this slot has finished
                for k, v in host.items():
                    out[k] = self._buffer(slot, k, v)
                    out[k].copy_(v, non_blocking=True)
                self._copied[slot].record(self.copy_stream)
            cur.wait_event(self._copied[slot])
        for k, v in tensors.items():
            if v.is_cuda:
                out[k] = v
        return out, slot

    def release(self, slot: int) -> None:
        ev = self._consumed[slot] or torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._consumed[slot] = ev
