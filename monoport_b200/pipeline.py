"""Frame-stream overlap.  The reference chains its per-frame stages as `processors=[callable, ...]`
with one Python thread per stage (RTL/dataloader.py:734-751, 1026-1054).  On one big GPU the useful overlap is between
FRAMES: frame k+1's encoder (PyTorch) can run while frame k's reconstruction kernels do.  `FramePipeline` keeps the
list-of-callables API, dispatches frames round-robin to `n_lanes` worker threads that each own a CUDA stream, and yields
results in order.  All monoport_b200 handles that carry per-call scratch are per-thread, so lanes do not share state."""
import ctypes
import threading
from collections import deque
from concurrent.futures import ThreadPoolExecutor

import torch

from . import _lib


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


class FrameGraph:
    """One frame = ONE CUDA-graph launch:  [image -> encoder (PyTorch) ->]  features -> coarse-to-fine engine -> visible surface.

    The reference overlaps its per-frame stages with one Python thread per stage (RTL/dataloader.py:734-751, 1026-1054) and
    pays a host round trip between every pair of them.  Here the whole frame step is captured once (the encoder's PyTorch
    kernels and this library's enqueue-only entry points mp_octree_run_fused_async / mp_forward_vertices_async on one
    stream) and replayed per frame: no Python, no launch gaps and no host synchronisation inside a frame -- even the choice
    between the tensor-core and the exact kernel is made on the device (range guard).  Several FrameGraphs own separate
    streams, workspaces and output buffers, so frame k+1's encoder overlaps frame k's reconstruction (`FrameGraphRing`).

    engine: a Seg3dLossless(faster=True) / Seg3dTopk built with make_query_func(net) (engines with a conflict loop read
    counts on the host and cannot be captured).  calib: [1,4,4] tensor (read once, at construction).
    `feature_hook(feat) -> feat` (optional) runs inside the captured step on the encoder's last-stage map.
    launch(x) enqueues a frame (x = image [1,3,H,W] with an encoder, else the [1,C,h,w] feature map); result() waits for it
    and returns (sdf or None, X, Y, Z, norm) exactly like engine(...) followed by forward_vertices(sdf, direction)."""

    def __init__(self, net, engine, calib, direction="front", with_encoder=True, input_shape=None, feature_hook=None):
        from .modeling.MonoPortNet import FeatureHandle
        from .modeling.geometry import perspective
        from .recon import _DIRS
        if not (engine.faster or engine.topk_points is not None):
            raise ValueError("FrameGraph needs an engine without a conflict loop (faster=True or Seg3dTopk)")
        if getattr(engine.query_func, "__monoport_fused__", None) is not net:
            raise ValueError("the engine's query_func must be make_query_func(net)")
        self.net, self.engine = net, engine
        dev = self.device = engine.b_min.device
        self.with_encoder = bool(with_encoder)
        self.feature_hook = feature_hook
        self.R = R = engine.resolutions[-1]
        self._dir = _DIRS[direction]
        self._proj = _lib.PROJ_PERSPECTIVE if net.projection is perspective else _lib.PROJ_ORTHOGONAL
        self._cal12 = _lib.calib12(calib.detach().cpu())
        if input_shape is None:
            input_shape = (1, 3, 512, 512) if with_encoder else (1, net.surface_classifier.filter_channels[0] - 1, 128, 128)
        lib = self._lib = _lib.load()
        with torch.cuda.device(dev):
            self.stream = torch.cuda.Stream(dev)
            self.static_in = torch.zeros(input_shape, dtype=torch.float32, device=dev)
            self.volume = torch.empty((1, 1, R, R, R), dtype=torch.float32, device=dev)
            cap = R * R
            self.X = torch.empty(cap, dtype=torch.int64, device=dev)
            self.Y = torch.empty(cap, dtype=torch.int64, device=dev)
            self.Z = torch.empty(cap, dtype=torch.float32, device=dev)
            self.N = torch.empty((cap, 3), dtype=torch.float32, device=dev)
            self.count = torch.zeros(1, dtype=torch.int64, device=dev)
            self._scratch = torch.empty(int(lib.mp_forward_vertices_scratch_bytes(R)), dtype=torch.uint8, device=dev)
            self._oct = engine._new_handle(dev)            # private workspace: graphs of different lanes overlap
            self._fh = None
            self._head = net.surface_classifier.handle()
            self._FeatureHandle = FeatureHandle
            # eager warm-up on the side stream (lazy allocations, cudnn autotune), then the capture
            self.stream.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(self.stream), torch.no_grad():
                for _ in range(2):
                    self._step()
            self.stream.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=self.stream), torch.no_grad():
                self._step()
        self.last_stats = None

    def _step(self):
        lib, dev = self._lib, self.device
        if self.with_encoder:
            feat = self.net.filter(self.static_in)[-1][0]
        else:
            feat = self.static_in
        if self.feature_hook is not None:
            feat = self.feature_hook(feat)
        if self._fh is None:
            _, C, H, W = feat.shape
            self._fh = self._FeatureHandle(C, H, W, dev)
        self._fh.key = None                      # always upload: the map changes with every replay
        self._fh.upload(feat)
        st = _lib.stream_ptr(dev)
        _lib.check(lib.mp_octree_run_fused_async(
            self._oct, self._head, self._fh.ptr, self._cal12, self._proj, ctypes.c_float(self.net.normalizer.scale),
            self.net._mode(), ctypes.c_void_p(self.volume.data_ptr()), st), "mp_octree_run_fused_async")
        _lib.check(lib.mp_forward_vertices_async(
            ctypes.c_void_p(self.volume.data_ptr()), self.R, self._dir, ctypes.c_void_p(self.X.data_ptr()),
            ctypes.c_void_p(self.Y.data_ptr()), ctypes.c_void_p(self.Z.data_ptr()), ctypes.c_void_p(self.N.data_ptr()),
            ctypes.c_void_p(self.count.data_ptr()), ctypes.c_void_p(self._scratch.data_ptr()), st), "mp_forward_vertices_async")

    def launch(self, x):
        """Enqueue one frame (returns at once)."""
        self.stream.wait_stream(torch.cuda.current_stream(self.device))     # x may still be produced on the caller's stream
        with torch.cuda.stream(self.stream):
            self.static_in.copy_(x, non_blocking=True)
            self.graph.replay()
        return self

    def result(self):
        """Wait for the enqueued frame: (sdf or None, X, Y, Z, norm) -- views of this graph's buffers, valid until its next launch."""
        n = len(self.engine.resolutions)
        nonempty = ctypes.c_int(0)
        stats = (ctypes.c_int64 * n)()
        with torch.cuda.stream(self.stream):
            _lib.check(self._lib.mp_octree_fetch(self._oct, ctypes.byref(nonempty), stats, _lib.stream_ptr(self.device)),
                       "mp_octree_fetch")                                    # (synchronises the graph's stream)
        self.last_stats = list(stats)
        if not nonempty.value:
            return None, None, None, None, None
        k = int(self.count.item())
        return self.volume, self.X[:k], self.Y[:k], self.Z[:k], self.N[:k]

    def close(self):
        if self._oct is not None:
            self.stream.synchronize()
            self._lib.mp_octree_destroy(self._oct)
            self._oct = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class FrameGraphRing:
    """`n_lanes` FrameGraphs used round-robin: up to n_lanes frames in flight, results in input order (the graph-captured
    counterpart of FramePipeline for the stages of this library)."""

    def __init__(self, make_graph, n_lanes=2):
        self.lanes = [make_graph() for _ in range(int(n_lanes))]

    def run(self, iterable):
        inflight = deque()
        for i, x in enumerate(iterable):
            lane = self.lanes[i % len(self.lanes)]
            if len(inflight) >= len(self.lanes):
                yield inflight.popleft().result()
            inflight.append(lane.launch(x))
        while inflight:
            yield inflight.popleft().result()

    def close(self):
        for lane in self.lanes:
            lane.close()
