"""Batched generic ``wp.Bvh`` queries: ``bvh_query_aabb`` / ``bvh_query_ray`` (+ the ``bvh_query_next`` loop)
of the reference (``warp/_src/builtins.py`` bvh_query_*, ``warp/native/bvh.h:494-600``) evaluated for a whole
batch.  The reference yields hits one at a time inside a user kernel; here every query's hits come back at
once in CSR form, in the order the reference iterator would produce them."""

from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .types import Bvh, Mesh, array, empty, float32, int32, vec3

FLT_MAX = float(np.finfo(np.float32).max)


class BvhQueryResult:
    """``offsets`` (int32, n + 1) and ``indices`` (int32, total): the items hit by query ``i`` are
    ``indices[offsets[i]:offsets[i + 1]]``."""

    __slots__ = ("offsets", "indices", "total")

    def __init__(self, offsets, indices, total):
        self.offsets, self.indices, self.total = offsets, indices, total

    def numpy(self):
        return self.offsets.numpy(), (self.indices.numpy() if self.total else np.zeros(0, np.int32))

    def lists(self):
        off, idx = self.numpy()
        return [idx[off[i] : off[i + 1]] for i in range(len(off) - 1)]


def _as_dev(a, dev, what):
    if isinstance(a, array):
        if a.dtype != vec3:
            raise RuntimeError(f"{what} should be an array of type wp.vec3")
        return a
    h = np.ascontiguousarray(a, dtype=np.float32)
    if h.ndim != 2 or h.shape[1] != 3:
        raise RuntimeError(f"{what} should have shape (n, 3)")
    return array(h, dtype=vec3, device=dev)


def _roots_arg(roots, n, dev):
    if roots is None:
        return None
    if not isinstance(roots, array):
        roots = array(np.ascontiguousarray(roots, dtype=np.int32), dtype=int32, device=dev)
    if roots.dtype != int32 or len(roots) != n:
        raise RuntimeError("roots should be an int32 array with one entry per query")
    return roots


def _radii_arg(radii, n, dev):
    """Per-query radii as a float32 device array; a scalar is broadcast."""
    if isinstance(radii, array):
        if radii.dtype != float32 or len(radii) != n:
            raise RuntimeError("radii should be a float32 array with one entry per query")
        return radii
    h = np.asarray(radii, dtype=np.float32)
    if h.ndim > 1 or (h.ndim == 1 and h.shape[0] != n):
        raise RuntimeError("radii should be one float or a float32 array with one entry per query")
    return array(np.ascontiguousarray(np.broadcast_to(h, (n,))), dtype=float32, device=dev)


def _run(bvh, qa, qb, kind, max_dist, mesh=False, roots=None, radii=None):
    """kind: "aabb" | "ray" | "sphere" | "capsule" (the reference's BvhQueryKind, bvh.h:420-427)."""
    if mesh:
        if not isinstance(bvh, Mesh) or not bvh.id:
            raise TypeError("expected a warp_b200.Mesh")
    elif not isinstance(bvh, Bvh) or not bvh.id:
        raise TypeError("expected a warp_b200.Bvh")
    dev = bvh.device
    qa = _as_dev(qa, dev, "first query array")
    qb = qa if qb is None else _as_dev(qb, dev, "second query array")
    if len(qa) != len(qb):
        raise RuntimeError("query arrays must have the same length")
    n = len(qa)
    c = _lib.core()
    p = lambda a: ctypes.c_void_p(a.ptr or 0)  # noqa: E731
    counts = empty(n, int32, dev)
    offsets = empty(n + 1, int32, dev)
    roots = _roots_arg(roots, n, dev)
    pr = ctypes.c_void_p(roots.ptr or 0) if roots is not None else None
    rad = _radii_arg(radii, n, dev) if kind in ("sphere", "capsule") else None

    def call(fill):
        tail = (p(offsets), p(indices)) if fill else (p(counts),)
        if mesh and kind == "sphere":
            fn = c.wp_b200_mesh_query_sphere_fill if fill else c.wp_b200_mesh_query_sphere_count
            return fn(bvh.id, p(qa), p(rad), n, *tail)
        if mesh:
            fn = c.wp_b200_mesh_query_aabb_fill if fill else c.wp_b200_mesh_query_aabb_count
            return fn(bvh.id, p(qa), p(qb), n, *tail)
        if kind == "ray":
            fn = c.wp_b200_bvh_query_ray_fill if fill else c.wp_b200_bvh_query_ray_count
            return fn(bvh.id, p(qa), p(qb), pr, n, max_dist, *tail)
        if kind == "sphere":
            fn = c.wp_b200_bvh_query_sphere_fill if fill else c.wp_b200_bvh_query_sphere_count
            return fn(bvh.id, p(qa), p(rad), pr, n, *tail)
        if kind == "capsule":
            fn = c.wp_b200_bvh_query_capsule_fill if fill else c.wp_b200_bvh_query_capsule_count
            return fn(bvh.id, p(qa), p(qb), p(rad), pr, n, max_dist, *tail)
        fn = c.wp_b200_bvh_query_aabb_fill if fill else c.wp_b200_bvh_query_aabb_count
        return fn(bvh.id, p(qa), p(qb), pr, n, *tail)

    indices = None
    ok = call(False) and c.wp_b200_exclusive_scan_i32(p(counts), p(offsets), n)
    if not ok:
        raise RuntimeError(f"bvh query failed: {_lib.error_string()}")
    total = int(offsets.numpy()[-1])  # the one host round trip: the hit list has to be allocated
    indices = empty(max(total, 1), int32, dev)
    if total and not call(True):
        raise RuntimeError(f"bvh query failed: {_lib.error_string()}")
    return BvhQueryResult(offsets, indices, total)


def bvh_query_aabb(bvh, lowers, uppers, roots=None) -> BvhQueryResult:
    """All items whose AABB overlaps ``[lowers[i], uppers[i]]`` (closed boxes, ``intersect.h:183-192``).
    ``roots`` (optional, one int per query): start the traversal at that node -- e.g. a group's subtree from
    :func:`bvh_get_group_root` -- instead of the tree root; ``-1`` means the tree root (``bvh.h:494-518``)."""
    return _run(bvh, lowers, uppers, "aabb", 0.0, roots=roots)


def bvh_query_ray(bvh, starts, dirs, max_dist: float = FLT_MAX, roots=None) -> BvhQueryResult:
    """All items whose AABB the ray ``starts[i] + t * dirs[i]`` enters at ``t < max_dist`` (``bvh.h:483-487``)."""
    return _run(bvh, starts, dirs, "ray", float(max_dist), roots=roots)


def bvh_query_sphere(bvh, centers, radii, roots=None) -> BvhQueryResult:
    """All items whose AABB lies within ``radii[i]`` (an array or one float) of ``centers[i]``: the exact sphere / box
    test of ``wp.bvh_query_sphere`` (``bvh.h:542-551``, ``intersect.h:197-205``); negative radii count as 0."""
    return _run(bvh, centers, None, "sphere", 0.0, roots=roots, radii=radii)


def bvh_query_capsule(bvh, starts, dirs, radii, max_dist: float = FLT_MAX, roots=None) -> BvhQueryResult:
    """All items whose AABB, inflated by ``radii[i]``, the ray ``starts[i] + t * dirs[i]`` enters at ``t <= max_dist``
    (closed, unlike the plain ray): ``wp.bvh_query_capsule`` + ``bvh_query_next(query, max_dist)`` (``bvh.h:472-482,
    529-540``).  With a unit ``dir`` and ``max_dist`` = segment length this is the swept-sphere broad phase."""
    return _run(bvh, starts, dirs, "capsule", float(max_dist), roots=roots, radii=radii)


def bvh_get_group_root(bvh, group_ids):
    """``wp.bvh_get_group_root`` for a batch of group ids (``bvh.h:376-390``): the node whose subtree holds exactly the
    items of each group (``-1`` where the group does not occur), as reference node indices usable as ``roots=``.
    Device array in -> device array out; anything else -> numpy array out."""
    if not isinstance(bvh, (Bvh, Mesh)) or not bvh.id:
        raise TypeError("expected a warp_b200.Bvh or warp_b200.Mesh")
    dev = bvh.device
    host = not isinstance(group_ids, array)
    g = array(np.ascontiguousarray(group_ids, dtype=np.int32), dtype=int32, device=dev) if host else group_ids
    if g.dtype != int32:
        raise RuntimeError("group_ids should be an int32 array")
    out = empty(len(g), int32, dev)
    if not _lib.core().wp_b200_bvh_get_group_root(bvh.id, ctypes.c_void_p(g.ptr or 0), len(g), ctypes.c_void_p(out.ptr or 0)):
        raise RuntimeError(f"bvh_get_group_root failed: {_lib.error_string()}")
    return out.numpy() if host else out


def mesh_query_aabb(mesh, lowers, uppers) -> BvhQueryResult:
    """All faces of ``mesh`` whose AABB (as of its last build / refit) overlaps ``[lowers[i], uppers[i]]``, in the
    order the reference's ``mesh_query_aabb`` / ``mesh_query_aabb_next`` loop yields them (``mesh.h:2476-2712``)."""
    return _run(mesh, lowers, uppers, "aabb", 0.0, mesh=True)


def mesh_query_sphere(mesh, centers, radii) -> BvhQueryResult:
    """All faces of ``mesh`` (as of its last build / refit) that intersect the sphere ``(centers[i], radii[i])`` --
    ``radii`` an array or one float -- in the order the reference's ``mesh_query_sphere`` / ``mesh_query_sphere_next``
    loop yields them (``mesh.h:2457-2737``): exact sphere / box broad phase, closest-point narrow phase."""
    return _run(mesh, centers, None, "sphere", 0.0, mesh=True, radii=radii)
