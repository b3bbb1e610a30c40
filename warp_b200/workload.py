"""Device-side synthetic workloads of the shapes SURVEY.md 8(d) names (bench.py / tests only; not on the query path):
counter-RNG query batches (config C5: 1 B queries generated per shard on the device) and the per-frame cloth update of
config C4, written so that a whole frame -- vertex update, query generation, refit, queries -- is one CUDA graph."""

from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .types import array, empty, int32, vec3


def box_queries(out: array, first_index: int, seed: int, lower, upper) -> array:
    """``out[i]`` = uniform point of the box [lower, upper] drawn from (seed, first_index + i): a shard of a global
    batch is the same on whichever rank generates it."""
    lo = (ctypes.c_float * 3)(*[float(x) for x in lower])
    hi = (ctypes.c_float * 3)(*[float(x) for x in upper])
    if not _lib.core().wp_b200_gen_box_queries(ctypes.c_void_p(out.ptr), len(out), int(first_index), int(seed), lo, hi):
        raise RuntimeError("query generation failed")
    return out


class ClothFrames:
    """n x n cloth of config C4 whose frame index lives in device memory, so the update can sit inside a graph."""

    def __init__(self, n_side: int, device=None):
        self.n = int(n_side)
        self.points = empty(self.n * self.n, vec3, device)
        self.frame = array(np.zeros(1, np.int32), dtype=int32, device=device)
        self.update_points()

    def advance(self, by: int = 1):
        if not _lib.core().wp_b200_counter_add(ctypes.c_void_p(self.frame.ptr), int(by)):
            raise RuntimeError("frame counter update failed")

    def update_points(self):
        """Vertices of the current frame, in place."""
        if not _lib.core().wp_b200_gen_cloth_points(ctypes.c_void_p(self.points.ptr), self.n, ctypes.c_void_p(self.frame.ptr), 0):
            raise RuntimeError("cloth update failed")

    def queries(self, out: array, sigma: float = 0.01) -> array:
        """Vertices of the PREVIOUS frame jittered by N(0, sigma): the collision candidates of the current frame."""
        if not _lib.core().wp_b200_gen_cloth_queries(ctypes.c_void_p(out.ptr), len(out), self.n, ctypes.c_void_p(self.frame.ptr),
                                                     float(sigma)):
            raise RuntimeError("cloth query generation failed")
        return out
