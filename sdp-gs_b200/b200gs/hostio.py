"""Host -> device staging of per-step inputs (Gaussian parameters that live in pinned host memory).

`PinnedFeeder` keeps the named arrays in ONE pinned allocation and two device-side copies of it.  Each
`next()` hands out the device tensors of the current step and immediately enqueues, on a dedicated copy
stream, the single host->device copy of the following step's inputs into the other device buffer, so the
PCIe transfer of step i+1 runs while the kernels of step i execute.  Every step's inputs still cross the
bus exactly once; only the waiting is removed.  Ordering is enforced with CUDA events (no host sync):
  * the compute stream waits for the copy that filled the buffer it is about to read;
  * the copy stream waits for the last kernel that read the buffer it is about to overwrite.
"""
from __future__ import annotations

import numpy as np
import torch


class PinnedFeeder:
    def __init__(self, arrays: dict, device, prefetch: bool = True):
        self.device = torch.device(device)
        self.prefetch = bool(prefetch)
        self.layout = {}
        off = 0
        for k, a in arrays.items():
            a = np.ascontiguousarray(a, dtype=np.float32)
            self.layout[k] = (off, a.size, tuple(a.shape))
            off += (a.size + 63) // 64 * 64  # 256-byte aligned segments
        self.total = off
        self.host = torch.empty((self.total,), dtype=torch.float32).pin_memory()
        for k, a in arrays.items():
            o, n, _ = self.layout[k]
            self.host[o:o + n].copy_(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32).reshape(-1)))
        self.nbytes = int(sum(n for _, n, _ in self.layout.values()) * 4)
        self.dev = [torch.empty((self.total,), dtype=torch.float32, device=self.device) for _ in range(2)]
        # persistent leaf views (requires_grad) of the two device buffers: handing them out costs nothing per step
        self.leaves = [{k: buf[o:o + n].view(shape).requires_grad_(True) for k, (o, n, shape) in self.layout.items()}
                       for buf in self.dev]
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.filled = [torch.cuda.Event(), torch.cuda.Event()]   # copy into dev[i] finished
        self.released = [torch.cuda.Event(), torch.cuda.Event()]  # last reader of dev[i] finished
        self.cur = 0
        self._primed = False

    def host_view(self, name):
        """The pinned host array behind `name` (write to it to change what later steps upload)."""
        o, n, shape = self.layout[name]
        return self.host[o:o + n].view(shape)

    def _enqueue_copy(self, i):
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.released[i])
            with torch.no_grad():
                self.dev[i].copy_(self.host, non_blocking=True)
            self.filled[i].record(self.copy_stream)

    def next(self):
        """Device tensors for this step (leaf tensors with requires_grad=True and .grad reset; valid until the
        next-but-one call)."""
        cs = torch.cuda.current_stream(self.device)
        if not self._primed:
            for ev in self.released:
                ev.record(cs)
            self._enqueue_copy(self.cur)
            self._primed = True
        i = self.cur
        cs.wait_event(self.filled[i])
        j = i ^ 1
        if self.prefetch:
            self._enqueue_copy(j)  # next step's upload overlaps this step's kernels
        out = self.leaves[i]
        for v in out.values():
            v.grad = None
        self._last = i
        self.cur = j
        return out

    def done(self):
        """Call after the step's last kernel that reads the tensors from next() has been enqueued."""
        cs = torch.cuda.current_stream(self.device)
        self.released[self._last].record(cs)
        if not self.prefetch:
            self._enqueue_copy(self.cur)


class ResultReader:
    """Device -> host read-back of one scalar per step without stalling the step that produced it.

    `push(t)` enqueues, on the current stream, the copy of the 1-element tensor `t` into a pinned slot and records an event;
    `pop()` returns the oldest value still in flight once more than `lag` results are pending (it waits for THAT result's
    event only, so the host is already enqueueing step i+1 while step i executes); `drain()` returns the rest.  Every step's
    result still crosses the bus, inside the loop that produced it -- only the per-step host stall of `.item()` is gone
    (the reference's loop has it: `loss.item()` every iteration, train.py:197)."""

    def __init__(self, lag: int = 1, slots: int = 8):
        self.lag = int(lag)
        self.host = torch.empty((slots,), dtype=torch.float32).pin_memory()
        self.events = [torch.cuda.Event() for _ in range(slots)]
        self.head = self.tail = 0

    def push(self, t):
        n = len(self.events)
        assert self.head - self.tail < n, "ResultReader: too many results in flight"
        i = self.head % n
        with torch.no_grad():
            self.host[i:i + 1].copy_(t.detach().reshape(1).float(), non_blocking=True)
        self.events[i].record()
        self.head += 1

    def _take(self):
        i = self.tail % len(self.events)
        self.events[i].synchronize()
        self.tail += 1
        return float(self.host[i])

    def pop(self):
        return self._take() if self.head - self.tail > self.lag else None

    def drain(self):
        out = []
        while self.head > self.tail:
            out.append(self._take())
        return out
