"""Frame-stream overlap.  The reference chains its per-frame stages as `processors=[callable, ...]`
with one Python thread per stage (RTL/dataloader.py:734-751, 1026-1054).  On one big GPU the useful overlap is between
FRAMES: frame k+1's encoder (PyTorch) can run while frame k's reconstruction kernels do.  `FramePipeline` keeps the
list-of-callables API, dispatches frames round-robin to `n_lanes` worker threads that each own a CUDA stream, and yields
results in order.  All monoport_b200 handles that carry per-call scratch are per-thread, so lanes do not share state."""
import threading
from collections import deque
from concurrent.futures import ThreadPoolExecutor

import torch


class FramePipeline:
    def __init__(self, processors, device="cuda:0", n_lanes=2):
        self.processors = list(processors)
        self.device = torch.device(device)
        self.n_lanes = int(n_lanes)
        self._pool = ThreadPoolExecutor(max_workers=self.n_lanes, thread_name_prefix="mp_lane")
        self._tls = threading.local()

    def _run_one(self, item):
        if not hasattr(self._tls, "stream"):
            torch.cuda.set_device(self.device)                 # RTL/dataloader.py:1031
            self._tls.stream = torch.cuda.Stream(self.device)
        with torch.cuda.stream(self._tls.stream):
            data = item
            for p in self.processors:
                data = p(data)
            self._tls.stream.synchronize()
        return data

    def run(self, iterable):
        """Generator: results in input order, up to n_lanes frames in flight."""
        inflight = deque()
        for item in iterable:
            inflight.append(self._pool.submit(self._run_one, item))
            if len(inflight) >= self.n_lanes:
                yield inflight.popleft().result()
        while inflight:
            yield inflight.popleft().result()

    def close(self):
        self._pool.shutdown(wait=True)
