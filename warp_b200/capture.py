"""Streams and CUDA-graph capture for the path's device work, after ``wp.Stream`` / ``wp.ScopedStream`` /
``wp.capture_begin`` / ``wp.capture_end`` / ``wp.capture_launch`` / ``wp.ScopedCapture``
(``warp/_src/context.py:12684-12900, 13532-13560``, ``warp/_src/utils.py:1946-2010``).

What can be captured: ``Mesh.refit()``, ``Mesh.points = ...``, ``Mesh.rebuild()`` / ``Bvh.rebuild()`` (in place, no
allocation -- the reason the reference has ``wp_bvh_rebuild_device``, ``bvh.cu:819-843``) and every query that takes
DEVICE arrays with a preallocated ``out=``.  A per-step collision loop (refit + queries, BASELINE config 4) becomes
one graph launch per frame.  Not capturable: constructors (they allocate), host-array queries (they synchronise),
and the CSR hit-list queries (the hit count comes back to the host between the count and the fill pass).
Run the loop body once before capturing so that grow-only scratch (query ordering) is already allocated and the
refit plan of a large tree exists.  A graph that contains ``refit()`` of a tree with 2**21 items or more replays the
plan of the build it was captured after: re-capture it after a ``rebuild()`` that ran outside the graph.
"""

from __future__ import annotations

import ctypes

from . import _lib
from .types import get_device

CAPTURE_MODE_GLOBAL, CAPTURE_MODE_THREAD_LOCAL, CAPTURE_MODE_RELAXED = 0, 1, 2


class Stream:
    """A CUDA stream on ``device`` (``wp.Stream``)."""

    def __init__(self, device=None, priority: int = 0):
        self.device = get_device(device)
        _lib.require_cuda()
        self.cuda_stream = _lib.core().wp_cuda_stream_create(self.device.context, int(priority))
        if not self.cuda_stream:
            raise RuntimeError(f"Failed to create stream: {_lib.error_string()}")

    def synchronize(self):
        _lib.core().wp_cuda_stream_synchronize(self.cuda_stream)

    def __del__(self):
        try:
            if getattr(self, "cuda_stream", None):
                _lib.core().wp_cuda_stream_destroy(self.device.context, self.cuda_stream)
        except Exception:
            pass


class ScopedStream:
    """Makes ``stream`` the device's current stream inside the ``with`` block (``wp.ScopedStream``)."""

    def __init__(self, stream: Stream, sync_enter: bool = True, sync_exit: bool = False):
        self.stream, self.sync_enter, self.sync_exit = stream, sync_enter, sync_exit

    def __enter__(self):
        c = _lib.core()
        ctx = self.stream.device.context
        self.saved = c.wp_cuda_context_get_stream(ctx)
        c.wp_cuda_context_set_stream(ctx, self.stream.cuda_stream, 1 if self.sync_enter else 0)
        return self.stream

    def __exit__(self, *exc):
        c = _lib.core()
        if self.sync_exit:
            self.stream.synchronize()
        c.wp_cuda_context_set_stream(self.stream.device.context, self.saved, 0)
        return False


class Graph:
    """An instantiated CUDA graph (``wp.Graph``, ``context.py:5989-6030``)."""

    def __init__(self, device, graph, graph_exec):
        self.device, self.graph, self.graph_exec = device, graph, graph_exec

    def __del__(self):
        try:
            c = _lib.core()
            if getattr(self, "graph_exec", None):
                c.wp_cuda_graph_exec_destroy(self.device.context, self.graph_exec)
            if getattr(self, "graph", None):
                c.wp_cuda_graph_destroy(self.device.context, self.graph)
        except Exception:
            pass


def _current_stream(device):
    return _lib.core().wp_cuda_context_get_stream(device.context)


def capture_begin(device=None, stream: Stream | None = None, mode: int = CAPTURE_MODE_THREAD_LOCAL) -> None:
    """Start capturing the work enqueued on ``stream`` (default: the device's current stream, which must be a
    created stream -- the legacy default stream cannot be captured)."""
    dev = stream.device if stream is not None else get_device(device)
    s = stream.cuda_stream if stream is not None else _current_stream(dev)
    if not s:
        raise RuntimeError("capture needs a non-default stream: use wp.ScopedStream(wp.Stream()) or wp.ScopedCapture()")
    if not _lib.core().wp_cuda_graph_begin_capture(dev.context, s, 0, int(mode)):
        raise RuntimeError(f"Failed to begin graph capture: {_lib.error_string()}")


def capture_end(device=None, stream: Stream | None = None) -> Graph:
    dev = stream.device if stream is not None else get_device(device)
    s = stream.cuda_stream if stream is not None else _current_stream(dev)
    c = _lib.core()
    graph, graph_exec = ctypes.c_void_p(), ctypes.c_void_p()
    if not c.wp_cuda_graph_end_capture(dev.context, s, ctypes.byref(graph)):
        raise RuntimeError(f"Error occurred during CUDA graph capture: {_lib.error_string()}")
    if not c.wp_cuda_graph_create_exec(dev.context, s, graph, ctypes.byref(graph_exec)):
        c.wp_cuda_graph_destroy(dev.context, graph)
        raise RuntimeError(f"Failed to instantiate the captured graph: {_lib.error_string()}")
    g = Graph(dev, graph, graph_exec)
    # the captured kernels point into the capture stream's grow-only scratch (query ordering), which is released when the
    # stream is destroyed: the graph keeps the stream alive
    g._captured_on = stream
    return g


def capture_launch(graph: Graph, stream: Stream | None = None) -> None:
    s = stream.cuda_stream if stream is not None else _current_stream(graph.device)
    if not _lib.core().wp_cuda_graph_launch(graph.graph_exec, s):
        raise RuntimeError(f"Graph launch error: {_lib.error_string()}")


class ScopedCapture:
    """``with wp.ScopedCapture() as capture: ...`` then ``wp.capture_launch(capture.graph)``.  Creates (and keeps) a
    stream of its own when the device's current stream is the default one; that stream is current inside the
    block, and graph launches default to the device's current stream at launch time."""

    def __init__(self, device=None, stream: Stream | None = None, mode: int = CAPTURE_MODE_THREAD_LOCAL):
        self.device = stream.device if stream is not None else get_device(device)
        self.stream, self.mode, self.graph = stream, mode, None
        self._scoped = None

    def __enter__(self):
        if self.stream is None and not _current_stream(self.device):
            self.stream = Stream(self.device)
        if self.stream is not None:
            self._scoped = ScopedStream(self.stream)
            self._scoped.__enter__()
        capture_begin(self.device, self.stream, self.mode)
        return self

    def __exit__(self, exc_type, exc, tb):
        try:
            if exc_type is None:
                self.graph = capture_end(self.device, self.stream)
            else:  # abandon the capture, keep the original exception
                g = ctypes.c_void_p()
                s = self.stream.cuda_stream if self.stream is not None else _current_stream(self.device)
                _lib.core().wp_cuda_graph_end_capture(self.device.context, s, ctypes.byref(g))
                if g:
                    _lib.core().wp_cuda_graph_destroy(self.device.context, g)
        finally:
            if self._scoped is not None:
                self._scoped.__exit__(None, None, None)
        return False
