"""Host -> device staging of index batches.

The reference builds every batch as Python lists and uploads six LongTensors synchronously
per step (knowledge_representation.py:179-184).  DevicePrefetcher keeps that contract -- the
caller hands over host (ideally pinned) index tensors, the model gets device tensors -- but
issues the copies of batch i+1 on a side stream while batch i is being scored, into a fixed
ring of device staging buffers (no allocation in steady state: an allocator round trip inside
the loop shows up as a multi-millisecond hiccup every few hundred launches).
"""
import torch


class DevicePrefetcher:
    """Iterate over host batches (tuples / lists of tensors), yielding device copies; the copy
    of the next batch overlaps the consumer's kernels.  Every tensor of a batch is copied inside
    the iteration that precedes its use, so a timed region around the loop contains all copies.
    The yielded tensors are views of a staging ring: they are valid until `depth` further batches
    have been requested (the consumer's kernels on them are ordered before the slot is reused)."""

    _streams = {}     # one copy stream per device for the life of the process
    _rings = {}       # staging rings keyed by (device, depth, shapes / dtypes): reused across epochs

    def __init__(self, batches, device, depth=2):
        self.it = iter(batches)
        self.device = torch.device(device)
        key = (self.device.type, self.device.index if self.device.index is not None else torch.cuda.current_device())
        if key not in DevicePrefetcher._streams:
            DevicePrefetcher._streams[key] = torch.cuda.Stream(device=self.device)
        self.copy_stream = DevicePrefetcher._streams[key]
        self.key = key
        self.queue = []
        self.depth = max(1, depth)
        self.slot = 0
        self.ring = None

    def _fresh(self, host):
        """New staging buffers.  The caching allocator may hand out memory whose previous owner's kernels are still
        queued on the CURRENT stream; the copy stream writes these buffers, so it must first catch up with it."""
        bufs = [torch.empty(x.shape, dtype=x.dtype, device=self.device) for x in host]
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.copy_stream.wait_event(ev)
        return bufs

    def _ring_for(self, host):
        sig = (self.key, self.depth, tuple((tuple(x.shape), x.dtype) for x in host))
        ring = DevicePrefetcher._rings.get(sig)
        if ring is None:
            n = self.depth + 1
            ring = {"bufs": [self._fresh(host) for _ in range(n)], "done": [None] * n}
            if len(DevicePrefetcher._rings) > 8:
                DevicePrefetcher._rings.clear()
            DevicePrefetcher._rings[sig] = ring
        return ring

    def _enqueue(self):
        try:
            host = next(self.it)
        except StopIteration:
            return False
        if self.ring is None:
            self.ring = self._ring_for(host)
        s = self.slot
        self.slot = (s + 1) % len(self.ring["bufs"])
        bufs = self.ring["bufs"][s]
        if len(bufs) != len(host) or any(b.shape != x.shape or b.dtype != x.dtype for b, x in zip(bufs, host)):
            bufs = self._fresh(host)                                                           # ragged last batch
            self.ring["done"][s] = None
        with torch.cuda.stream(self.copy_stream):
            if self.ring["done"][s] is not None:
                self.copy_stream.wait_event(self.ring["done"][s])      # the consumer's kernels on this slot are done
            for b, x in zip(bufs, host):
                b.copy_(x, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.queue.append((bufs, ev, host, s))
        return True

    def __iter__(self):
        while len(self.queue) < self.depth and self._enqueue():
            pass
        while self.queue:
            dev, ev, _host, s = self.queue.pop(0)
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            self._enqueue()
            yield dev
            done = torch.cuda.Event()          # the consumer has enqueued its work on this batch by now
            done.record(torch.cuda.current_stream(self.device))
            self.ring["done"][s] = done
