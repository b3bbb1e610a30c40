"""Batched equivalents of ``wp.mesh_query_point_no_sign``, ``wp.mesh_query_point`` and
``wp.mesh_query_ray`` (``warp/_src/builtins.py:8336-8440, 8569-8655, 8974-9081``).

The reference evaluates these per thread inside user kernels and returns a ``MeshQueryPoint`` /
``MeshQueryRay`` struct; here one call evaluates a whole batch and returns the same fields as
arrays (struct of arrays).  On a miss ``result`` is 0 and every other field is 0, as for the
struct-returning overloads (``warp/native/mesh.h:1514-1540, 2216-2257``).

Inputs may be device ``array``s (results are device ``array``s; asynchronous on the current stream)
or host numpy arrays (results are numpy arrays; the native ``*_host`` entry points stage the batch
through the GPU in chunks and return when the results are valid).
"""

from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .types import Mesh, array, empty, float32, from_numpy, int32, uint8, vec3, zeros


class MeshQueryPoint:
    """``result`` (uint8), ``sign``, ``face``, ``u``, ``v`` -- one entry per query (types.py:7786-7815)."""

    __slots__ = ("result", "sign", "face", "u", "v")

    def __init__(self, result, sign, face, u, v):
        self.result, self.sign, self.face, self.u, self.v = result, sign, face, u, v

    def numpy(self):
        return {k: (getattr(self, k).numpy() if isinstance(getattr(self, k), array) else getattr(self, k))
                for k in self.__slots__}  # fmt: skip


class MeshQueryRay:
    """``result``, ``sign``, ``face``, ``t``, ``u``, ``v``, ``normal`` -- one entry per ray (types.py:7820-7848)."""

    __slots__ = ("result", "sign", "face", "t", "u", "v", "normal")

    def __init__(self, result, sign, face, t, u, v, normal):
        self.result, self.sign, self.face, self.t, self.u, self.v, self.normal = result, sign, face, t, u, v, normal

    def numpy(self):
        return {k: (getattr(self, k).numpy() if isinstance(getattr(self, k), array) else getattr(self, k))
                for k in self.__slots__}  # fmt: skip


def _mesh_id(mesh):
    if isinstance(mesh, Mesh):
        if not mesh.id:
            raise RuntimeError("mesh has no native object")
        return mesh.id, mesh.device
    raise TypeError("expected a warp_b200.Mesh")


def _dev_vec3(a, what):
    if a.dtype != vec3:
        raise RuntimeError(f"{what} should be an array of type wp.vec3")
    return a


def _host_vec3(a, what):
    h = np.ascontiguousarray(a, dtype=np.float32)
    if h.ndim != 2 or h.shape[1] != 3:
        raise RuntimeError(f"{what} should have shape (n, 3)")
    return h


def _p(a):
    if isinstance(a, array):
        return ctypes.c_void_p(a.ptr or 0)
    return ctypes.c_void_p(a.ctypes.data)


def _check(ok, what):
    if not ok:
        raise RuntimeError(f"{what} failed: {_lib.error_string()}")


_OUT_DTYPES = {"result": uint8, "sign": float32, "face": int32, "t": float32, "u": float32, "v": float32, "normal": vec3}


def _check_out(out, n, dev, fields):
    """A caller-supplied ``out`` is written by the kernel without further checks: refuse short, mistyped or
    wrong-device buffers here instead of corrupting device memory."""
    for f in fields:
        a = getattr(out, f)
        if isinstance(a, array):
            if a.device != dev:
                raise RuntimeError(f"out.{f} lives on {a.device}, the mesh on {dev}")
            if a.dtype != _OUT_DTYPES[f] or len(a) < n:
                raise RuntimeError(f"out.{f} must be an array of at least {n} {_OUT_DTYPES[f]} entries, got {a}")
        else:
            want = _OUT_DTYPES[f].np_dtype
            shape_ok = a.shape[0] >= n and (a.ndim == 2 and a.shape[1] == 3 if f == "normal" else a.ndim == 1)
            if a.dtype != want or not shape_ok or not a.flags["C_CONTIGUOUS"]:
                raise RuntimeError(f"out.{f} must be a contiguous numpy array of at least {n} {want} entries")


def _point_query(mesh, points, max_dist, with_sign, out):
    id_, dev = _mesh_id(mesh)
    c = _lib.core()
    if isinstance(points, array):
        pts = _dev_vec3(points, "points")
        if pts.device != dev:
            raise RuntimeError(f"points live on {pts.device}, the mesh on {dev}")
        n = len(pts)
        if out is None:
            # `sign` of the unsigned query is 0 (mesh.h:1602-1608): zeroed once here; a caller-supplied `out` keeps
            # whatever its sign array holds (no 4-byte-per-query memset on the per-step path)
            out = MeshQueryPoint(empty(n, uint8, dev), empty(n, float32, dev) if with_sign else zeros(n, float32, dev),
                                 empty(n, int32, dev), empty(n, float32, dev), empty(n, float32, dev))  # fmt: skip
        else:
            _check_out(out, n, dev, ("result", "face", "u", "v") + (("sign",) if with_sign else ()))
        if with_sign:
            ok = c.wp_b200_mesh_query_point(id_, _p(pts), n, max_dist, _p(out.result), _p(out.sign), _p(out.face),
                                            _p(out.u), _p(out.v))  # fmt: skip
        else:
            ok = c.wp_b200_mesh_query_point_no_sign(id_, _p(pts), n, max_dist, _p(out.result), _p(out.face),
                                                    _p(out.u), _p(out.v))  # fmt: skip
        _check(ok, "mesh_query_point")
        return out
    pts = _host_vec3(points, "points")
    n = pts.shape[0]
    if out is None:
        out = MeshQueryPoint(np.zeros(n, np.uint8), np.zeros(n, np.float32), np.zeros(n, np.int32),
                             np.zeros(n, np.float32), np.zeros(n, np.float32))  # fmt: skip
    else:
        _check_out(out, n, dev, ("result", "face", "u", "v") + (("sign",) if with_sign else ()))
    if with_sign:
        ok = c.wp_b200_mesh_query_point_host(id_, _p(pts), n, max_dist, _p(out.result), _p(out.sign), _p(out.face),
                                             _p(out.u), _p(out.v))  # fmt: skip
    else:
        ok = c.wp_b200_mesh_query_point_no_sign_host(id_, _p(pts), n, max_dist, _p(out.result), _p(out.face),
                                                     _p(out.u), _p(out.v))  # fmt: skip
    _check(ok, "mesh_query_point")
    return out


def mesh_query_point_no_sign(mesh, points, max_dist: float, out: MeshQueryPoint | None = None) -> MeshQueryPoint:
    """Closest point on ``mesh`` to each of ``points`` within ``max_dist``; ``sign`` is 0 (mesh.h:501-676)."""
    return _point_query(mesh, points, float(max_dist), False, out)


def mesh_query_point(mesh, points, max_dist: float, out: MeshQueryPoint | None = None) -> MeshQueryPoint:
    """Closest point + inside/outside ``sign`` by three axis-ray probes (mesh.h:128-307, 2342-2359)."""
    return _point_query(mesh, points, float(max_dist), True, out)


def mesh_query_ray(mesh, starts, dirs, max_t: float, out: MeshQueryRay | None = None, roots=None) -> MeshQueryRay:
    """Closest hit of each ray ``starts[i] + t * dirs[i]``, ``0 <= t < max_t`` (mesh.h:1768-1891).  ``roots`` (device
    arrays only): per-ray start node, e.g. a group root from :func:`warp_b200.bvh_get_group_root`."""
    id_, dev = _mesh_id(mesh)
    c = _lib.core()
    max_t = float(max_t)
    if isinstance(starts, array) != isinstance(dirs, array):
        raise RuntimeError("starts and dirs must both be device arrays or both be host arrays")
    if isinstance(starts, array):
        s, d = _dev_vec3(starts, "starts"), _dev_vec3(dirs, "dirs")
        if len(s) != len(d):
            raise RuntimeError("starts and dirs must have the same length")
        n = len(s)
        if s.device != dev or d.device != dev:
            raise RuntimeError(f"starts / dirs live on {s.device} / {d.device}, the mesh on {dev}")
        if out is None:
            out = MeshQueryRay(empty(n, uint8, dev), empty(n, float32, dev), empty(n, int32, dev),
                               empty(n, float32, dev), empty(n, float32, dev), empty(n, float32, dev),
                               empty(n, vec3, dev))  # fmt: skip
        else:
            _check_out(out, n, dev, MeshQueryRay.__slots__)
        r = _roots(roots, n, dev)
        ok = c.wp_b200_mesh_query_ray(id_, _p(s), _p(d), n, max_t, _p(out.result), _p(out.sign), _p(out.face),
                                      _p(out.t), _p(out.u), _p(out.v), _p(out.normal), _p(r) if r is not None else None)  # fmt: skip
        _check(ok, "mesh_query_ray")
        return out
    if roots is not None:
        raise RuntimeError("roots= needs device arrays (wp.array) for starts and dirs")
    s, d = _host_vec3(starts, "starts"), _host_vec3(dirs, "dirs")
    if s.shape != d.shape:
        raise RuntimeError("starts and dirs must have the same length")
    n = s.shape[0]
    if out is None:
        out = MeshQueryRay(np.zeros(n, np.uint8), np.zeros(n, np.float32), np.zeros(n, np.int32),
                           np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32),
                           np.zeros((n, 3), np.float32))  # fmt: skip
    else:
        _check_out(out, n, dev, MeshQueryRay.__slots__)
    ok = c.wp_b200_mesh_query_ray_host(id_, _p(s), _p(d), n, max_t, _p(out.result), _p(out.sign), _p(out.face),
                                       _p(out.t), _p(out.u), _p(out.v), _p(out.normal))  # fmt: skip
    _check(ok, "mesh_query_ray")
    return out


def _roots(roots, n, dev):
    """Optional per-query int32 roots as a device array (host arrays are uploaded); the caller holds the result
    until the native call has returned -- freeing a device array synchronises, so an upload cannot vanish under
    a kernel that is still reading it."""
    if roots is None:
        return None
    if not isinstance(roots, array):
        roots = from_numpy(np.ascontiguousarray(roots, dtype=np.int32), int32, dev)
    if roots.dtype != int32 or len(roots) != n:
        raise RuntimeError("roots should be an int32 array with one entry per ray")
    return roots


def _stage(a, dtype, dev, what):
    """(device array, was_host): host arrays are uploaded; device arrays are type-checked and passed through."""
    if isinstance(a, array):
        if a.dtype != dtype:
            raise RuntimeError(f"{what} should be an array of type {dtype!r}")
        return a, False
    if dtype == vec3:
        return from_numpy(_host_vec3(a, what), vec3, dev), True
    h = np.ascontiguousarray(a, dtype=dtype.np_dtype)
    if h.ndim != 1:
        raise RuntimeError(f"{what} should be one-dimensional")
    return from_numpy(h, dtype, dev), True


def mesh_query_point_sign_parity(mesh, points, max_dist: float, n_sample: int = 1, perturbation_scale: float = 0.1):
    """Closest point + inside/outside by ray-crossing parity (``mesh.h:309-498, 2362-2392``): ``sign`` is -1 when at
    least half of ``n_sample`` rays along ``(1,1,1) + U(-s, s)^3`` (deterministic stream) cross an odd number of
    faces.  Device array in -> device arrays out; host array in -> numpy arrays out."""
    id_, dev = _mesh_id(mesh)
    pts, host = _stage(points, vec3, dev, "points")
    n = len(pts)
    out = MeshQueryPoint(empty(n, uint8, dev), empty(n, float32, dev), empty(n, int32, dev), empty(n, float32, dev),
                         empty(n, float32, dev))  # fmt: skip
    ok = _lib.core().wp_b200_mesh_query_point_sign_parity(id_, _p(pts), n, float(max_dist), int(n_sample),
                                                          float(perturbation_scale), _p(out.result), _p(out.sign),
                                                          _p(out.face), _p(out.u), _p(out.v))  # fmt: skip
    _check(ok, "mesh_query_point_sign_parity")
    if host:
        return MeshQueryPoint(*(getattr(out, k).numpy() for k in MeshQueryPoint.__slots__))
    return out


def mesh_query_point_sign_normal(mesh, points, max_dist: float, epsilon: float = 1e-3):
    """Closest point + inside/outside from angle-weighted normals (``mesh.h:860-1090``): faces whose distance is within
    ``average_edge_length * epsilon`` of the minimum are welded, ``sign`` is +1 when their accumulated normal points
    towards the query.  Device array in -> device arrays out; host array in -> numpy arrays out."""
    id_, dev = _mesh_id(mesh)
    pts, host = _stage(points, vec3, dev, "points")
    n = len(pts)
    out = MeshQueryPoint(empty(n, uint8, dev), empty(n, float32, dev), empty(n, int32, dev), empty(n, float32, dev),
                         empty(n, float32, dev))  # fmt: skip
    ok = _lib.core().wp_b200_mesh_query_point_sign_normal(id_, _p(pts), n, float(max_dist), float(epsilon), _p(out.result),
                                                          _p(out.sign), _p(out.face), _p(out.u), _p(out.v))  # fmt: skip
    _check(ok, "mesh_query_point_sign_normal")
    if host:
        return MeshQueryPoint(*(getattr(out, k).numpy() for k in MeshQueryPoint.__slots__))
    return out


def mesh_average_edge_length(mesh) -> float:
    """``Mesh.average_edge_length`` of the current points (``mesh.cu:38-60``), the welding scale of
    :func:`mesh_query_point_sign_normal`."""
    import ctypes

    id_, _ = _mesh_id(mesh)
    v = ctypes.c_float(0.0)
    _check(_lib.core().wp_b200_mesh_average_edge_length(id_, ctypes.byref(v)), "mesh_average_edge_length")
    return float(np.float32(v.value))


def _ray_pair(mesh, starts, dirs):
    id_, dev = _mesh_id(mesh)
    if isinstance(starts, array) != isinstance(dirs, array):
        raise RuntimeError("starts and dirs must both be device arrays or both be host arrays")
    s, host = _stage(starts, vec3, dev, "starts")
    d, _ = _stage(dirs, vec3, dev, "dirs")
    if len(s) != len(d):
        raise RuntimeError("starts and dirs must have the same length")
    return id_, dev, s, d, host


def mesh_query_ray_anyhit(mesh, starts, dirs, max_t: float, roots=None):
    """``result[i]`` = some triangle is hit by ray i with ``0 <= t < max_t`` (mesh.h:1893-1974).
    Device arrays in -> device ``uint8`` array out; host arrays in -> numpy ``bool`` array out."""
    id_, dev, s, d, host = _ray_pair(mesh, starts, dirs)
    out = empty(len(s), uint8, dev)
    r = _roots(roots, len(s), dev)
    ok = _lib.core().wp_b200_mesh_query_ray_anyhit(id_, _p(s), _p(d), len(s), float(max_t), _p(out),
                                                   _p(r) if r is not None else None)  # fmt: skip
    _check(ok, "mesh_query_ray_anyhit")
    return out.numpy().astype(bool) if host else out


def mesh_query_ray_count_intersections(mesh, starts, dirs, roots=None):
    """Number of triangles hit by ray i with ``t >= 0``, over the whole ray (mesh.h:1976-2032)."""
    id_, dev, s, d, host = _ray_pair(mesh, starts, dirs)
    out = empty(len(s), int32, dev)
    r = _roots(roots, len(s), dev)
    ok = _lib.core().wp_b200_mesh_query_ray_count_intersections(id_, _p(s), _p(d), len(s), _p(out),
                                                                _p(r) if r is not None else None)  # fmt: skip
    _check(ok, "mesh_query_ray_count_intersections")
    return out.numpy() if host else out


def _mesh_eval(mesh, face, u, v, velocity):
    id_, dev = _mesh_id(mesh)
    kinds = {isinstance(x, array) for x in (face, u, v)}
    if len(kinds) != 1:
        raise RuntimeError("face, u and v must all be device arrays or all be host arrays")
    f, host = _stage(face, int32, dev, "face")
    uu, _ = _stage(u, float32, dev, "u")
    vv, _ = _stage(v, float32, dev, "v")
    if not (len(f) == len(uu) == len(vv)):
        raise RuntimeError("face, u and v must have the same length")
    out = empty(len(f), vec3, dev)
    c = _lib.core()
    fn = c.wp_b200_mesh_eval_velocity if velocity else c.wp_b200_mesh_eval_position
    _check(fn(id_, _p(f), _p(uu), _p(vv), len(f), _p(out)), "mesh_eval")
    return out.numpy() if host else out


def mesh_eval_position(mesh, face, u, v):
    """``p*u + q*v + r*(1-u-v)`` on triangle ``face[i]`` of the mesh's current points (mesh.h:2767-2785)."""
    return _mesh_eval(mesh, face, u, v, False)


def mesh_eval_velocity(mesh, face, u, v):
    """Same interpolation over the mesh's velocities; zeros when it has none (mesh.h:2787-2805)."""
    return _mesh_eval(mesh, face, u, v, True)


def mesh_eval_face_normal(mesh, face):
    """``normalize(cross(q - p, r - p))`` of triangle ``face[i]`` of the mesh's current points (mesh.h:2870-2888)."""
    id_, dev = _mesh_id(mesh)
    f, host = _stage(face, int32, dev, "face")
    out = empty(len(f), vec3, dev)
    _check(_lib.core().wp_b200_mesh_eval_face_normal(id_, _p(f), len(f), _p(out)), "mesh_eval_face_normal")
    return out.numpy() if host else out


def mesh_query_furthest_point_no_sign(mesh, points, min_dist: float):
    """Farthest point of the mesh from each query (``mesh.h:678-858``; always a vertex), found when it lies strictly
    beyond ``min_dist``.  ``sign`` stays 0.  Device array in -> device arrays out; host array in -> numpy arrays out."""
    id_, dev = _mesh_id(mesh)
    pts, host = _stage(points, vec3, dev, "points")
    n = len(pts)
    out = MeshQueryPoint(empty(n, uint8, dev), zeros(n, float32, dev), empty(n, int32, dev), empty(n, float32, dev),
                         empty(n, float32, dev))  # fmt: skip
    ok = _lib.core().wp_b200_mesh_query_furthest_point_no_sign(id_, _p(pts), n, float(min_dist), _p(out.result), _p(out.face),
                                                               _p(out.u), _p(out.v))  # fmt: skip
    _check(ok, "mesh_query_furthest_point_no_sign")
    if host:
        return MeshQueryPoint(*(getattr(out, k).numpy() for k in MeshQueryPoint.__slots__))
    return out


class query_stats:
    """Context manager: count 64-byte pair fetches and 48-byte triangle fetches of the queries inside."""

    def __enter__(self):
        _lib.core().wp_b200_query_stats_enable(1)
        self.pair_fetches = self.tri_fetches = 0
        return self

    def __exit__(self, *exc):
        a, b = ctypes.c_ulonglong(), ctypes.c_ulonglong()
        _lib.core().wp_b200_query_stats_read(ctypes.byref(a), ctypes.byref(b))
        self.pair_fetches, self.tri_fetches = a.value, b.value
        _lib.core().wp_b200_query_stats_enable(0)
        return False


QUERY_ORDER_INPUT, QUERY_ORDER_MORTON, QUERY_ORDER_AUTO = 0, 1, 2


def set_query_order(mode: int) -> None:
    """How threads are assigned to the points of a batch: ``QUERY_ORDER_INPUT`` (thread i answers
    point i), ``QUERY_ORDER_MORTON`` (the batch is sorted along a space-filling curve on the device first -- the
    24-bit Hilbert curve, or the Morton curve for the signed query -- so a warp walks one neighbourhood of the
    tree), ``QUERY_ORDER_AUTO`` (default: sorted for >= 32768 points).  Answers are identical in every mode."""
    _lib.core().wp_b200_set_query_order(int(mode))


def get_query_order() -> int:
    return int(_lib.core().wp_b200_get_query_order())


def set_ray_order(mode: int) -> None:
    """``QUERY_ORDER_INPUT`` (default: thread i traces ray i -- right for primary rays, which are coherent
    as given) or ``QUERY_ORDER_MORTON`` (the batch is sorted by origin cell, then direction cell, on the
    device first -- for incoherent batches).  Answers are identical in both modes."""
    _lib.core().wp_b200_set_ray_order(1 if mode else 0)


def get_ray_order() -> int:
    return int(_lib.core().wp_b200_get_ray_order())


REFIT_AUTO, REFIT_ATOMIC, REFIT_WAVEFRONT = 0, 1, 2


def set_refit_mode(mode: int) -> None:
    """How ``refit()`` walks the tree: ``REFIT_ATOMIC`` = one atomic arrival counter per internal node (the reference's
    scheme, ``bvh.cu:42-144``); ``REFIT_WAVEFRONT`` = a visiting order planned once per build, levels inside blocks of
    1024 sorted positions in shared memory, counters only above the blocks; ``REFIT_AUTO`` (default) picks the
    wavefront from 2**21 items up.  The boxes are identical bit for bit."""
    _lib.core().wp_b200_set_refit_mode(int(mode))


def get_refit_mode() -> int:
    return int(_lib.core().wp_b200_get_refit_mode())
