"""Host -> device staging of index batches.

The reference builds every batch as Python lists and uploads six LongTensors synchronously
per step (knowledge_representation.py:179-184).  DevicePrefetcher keeps that contract -- the
caller hands over host (ideally pinned) index tensors, the model gets device tensors -- but
issues the copies of batch i+1 on a side stream while batch i is being scored.
"""
import torch


class DevicePrefetcher:
    """Iterate over host batches (tuples / lists of tensors), yielding device copies; the copy
    of the next batch overlaps the consumer's kernels.  Every tensor of a batch is copied inside
    the iteration that precedes its use, so a timed region around the loop contains all copies."""

    _streams = {}     # one copy stream per device for the life of the process: the caching allocator keeps
                      # its free blocks per stream, a fresh stream per epoch would start every epoch with cudaMalloc

    def __init__(self, batches, device, depth=2):
        self.it = iter(batches)
        self.device = torch.device(device)
        key = (self.device.type, self.device.index if self.device.index is not None else torch.cuda.current_device())
        if key not in DevicePrefetcher._streams:
            DevicePrefetcher._streams[key] = torch.cuda.Stream(device=self.device)
        self.copy_stream = DevicePrefetcher._streams[key]
        self.queue = []
        self.depth = max(1, depth)

    def _enqueue(self):
        try:
            host = next(self.it)
        except StopIteration:
            return False
        with torch.cuda.stream(self.copy_stream):
            dev = [x.to(self.device, non_blocking=True) for x in host]
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.queue.append((dev, ev, host))
        return True

    def __iter__(self):
        while len(self.queue) < self.depth and self._enqueue():
            pass
        while self.queue:
            dev, ev, _host = self.queue.pop(0)
            torch.cuda.current_stream(self.device).wait_event(ev)
            for x in dev:                       # the consumer's stream now owns these buffers
                x.record_stream(torch.cuda.current_stream(self.device))
            self._enqueue()
            yield dev
