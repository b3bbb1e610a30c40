"""TEST INFRASTRUCTURE ONLY -- ctypes/numpy front-end to the two CPU checkers.

* ``liblbvh_oracle.so``  : our plain-C restatement (``oracle/lbvh_oracle.c``); always buildable.
* ``_ref/libwarp_ref_cpu.so`` : the UNMODIFIED reference C++ (``warp/native/{bvh,mesh}.cpp`` +
  headers) compiled where it lies by ``oracle/Makefile``; present when built in the dev container
  and shipped (git-ignored, not gpurun-ignored) to the GPU box.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs import this module.  ``warp_b200`` never does.
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORC_PATH = os.path.join(_HERE, "liblbvh_oracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libwarp_ref_cpu.so")

HALF_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("ib", "<u4")])

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_u64p = ctypes.POINTER(ctypes.c_uint64)


def build(force: bool = False) -> None:
    """Compile the checkers (``make -C oracle``); the ``_ref`` target is a no-op without /root/reference."""
    if force or not os.path.exists(_ORC_PATH) or (
        os.path.getmtime(_ORC_PATH) < max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("lbvh_oracle.c", "bvh_query_oracle.c"))
    ):
        subprocess.run(["make", "-C", _HERE, "liblbvh_oracle.so"], check=True, capture_output=True)
    if os.path.isdir("/root/reference/warp/native") and (force or not os.path.exists(_REF_PATH)):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _f32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a if shape is None else a.reshape(shape)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


_orc = None


def orc():
    global _orc
    if _orc is None:
        build()
        _orc = ctypes.CDLL(_ORC_PATH)
        _orc.orc_morton3_1024.restype = ctypes.c_uint32
        _orc.orc_morton3_1024.argtypes = [ctypes.c_float] * 3
    return _orc


def ref_available() -> bool:
    if not os.path.exists(_REF_PATH) and os.path.isdir("/root/reference/warp/native"):
        try:
            build()
        except Exception:
            return False
    return os.path.exists(_REF_PATH)


_ref = None


def ref():
    global _ref
    if _ref is None:
        if not ref_available():
            raise RuntimeError("oracle/_ref/libwarp_ref_cpu.so is not built (needs /root/reference)")
        _ref = ctypes.CDLL(_REF_PATH)
        _ref.ref_mesh_create.restype = ctypes.c_uint64
        _ref.ref_mesh_from_tree.restype = ctypes.c_uint64
        _ref.ref_morton3_1024.restype = ctypes.c_uint32
        _ref.ref_morton3_1024.argtypes = [ctypes.c_float] * 3
        _ref.ref_get_error_string.restype = ctypes.c_char_p
    return _ref


# ----------------------------------------------------------------------------------------------
# our restatement
# ----------------------------------------------------------------------------------------------


def triangle_bounds(points, indices):
    points = _f32(points, (-1, 3))
    indices = _i32(indices).reshape(-1)
    t = indices.size // 3
    lo = np.empty((t, 3), np.float32)
    hi = np.empty((t, 3), np.float32)
    orc().orc_triangle_bounds(_p(points, _f32p), _p(indices, _i32p), ctypes.c_int(t), _p(lo, _f32p), _p(hi, _f32p))
    return lo, hi


def morton3(x, y, z) -> int:
    return int(orc().orc_morton3_1024(x, y, z))


def lbvh_build(lowers, uppers, leaf_size=1, groups=None, morton_bits=30):
    """Reference-layout LBVH over item boxes (restates bvh.cu:515-613).  Returns a dict of arrays."""
    lowers = _f32(lowers, (-1, 3))
    uppers = _f32(uppers, (-1, 3))
    n = lowers.shape[0]
    m = max(2 * n - 1, 1)
    out = {
        "n": n,
        "leaf_size": leaf_size,
        "keys": np.zeros(n, np.uint64),
        "primitive_indices": np.zeros(n, np.int32),
        "node_lowers": np.zeros(m, HALF_DTYPE),
        "node_uppers": np.zeros(m, HALF_DTYPE),
        "parents": np.zeros(m, np.int32),
        "total_lower": np.zeros(3, np.float32),
        "total_upper": np.zeros(3, np.float32),
        "inv_edges": np.zeros(3, np.float32),
    }
    root = ctypes.c_int(-1)
    g = _i32(groups) if groups is not None else None
    orc().orc_lbvh_build(
        _p(lowers, _f32p), _p(uppers, _f32p), ctypes.c_int(n), _p(g, _i32p), ctypes.c_int(leaf_size),
        ctypes.c_int(morton_bits), _p(out["keys"], _u64p), _p(out["primitive_indices"], _i32p),
        out["node_lowers"].ctypes.data_as(ctypes.c_void_p), out["node_uppers"].ctypes.data_as(ctypes.c_void_p),
        _p(out["parents"], _i32p), ctypes.byref(root),
        _p(out["total_lower"], _f32p), _p(out["total_upper"], _f32p), _p(out["inv_edges"], _f32p),
    )  # fmt: skip
    out["root"] = root.value
    return out


def lbvh_refit(tree, lowers, uppers):
    """In-place refit of ``tree`` (from :func:`lbvh_build`) to new item boxes (restates bvh.cu:42-144)."""
    lowers = _f32(lowers, (-1, 3))
    uppers = _f32(uppers, (-1, 3))
    orc().orc_lbvh_refit(
        ctypes.c_int(tree["n"]), _p(tree["parents"], _i32p), _p(tree["primitive_indices"], _i32p),
        tree["node_lowers"].ctypes.data_as(ctypes.c_void_p), tree["node_uppers"].ctypes.data_as(ctypes.c_void_p),
        _p(lowers, _f32p), _p(uppers, _f32p),
    )  # fmt: skip
    return tree


def mesh_lbvh_build(points, indices, leaf_size=4, groups=None, morton_bits=30):
    lo, hi = triangle_bounds(points, indices)
    t = lbvh_build(lo, hi, leaf_size, groups, morton_bits)
    t["item_lowers"], t["item_uppers"] = lo, hi
    return t


def _tree_args(points, indices, tree):
    points = _f32(points, (-1, 3))
    indices = _i32(indices).reshape(-1)
    return points, indices, (
        _p(points, _f32p), _p(indices, _i32p),
        tree["node_lowers"].ctypes.data_as(ctypes.c_void_p), tree["node_uppers"].ctypes.data_as(ctypes.c_void_p),
        _p(tree["primitive_indices"], _i32p), ctypes.c_int(tree["root"]),
    )  # fmt: skip


def query_point_no_sign(points, indices, tree, queries, max_dist, stats=False):
    points, indices, targs = _tree_args(points, indices, tree)
    q = _f32(queries, (-1, 3))
    n = q.shape[0]
    res = {"result": np.zeros(n, np.uint8), "face": np.zeros(n, np.int32), "u": np.zeros(n, np.float32),
           "v": np.zeros(n, np.float32)}  # fmt: skip
    st = np.zeros(2, np.uint64) if stats else None
    orc().orc_query_point_no_sign(
        *targs, _p(q, _f32p), ctypes.c_int64(n), ctypes.c_float(max_dist),
        _p(res["result"], _u8p), _p(res["face"], _i32p), _p(res["u"], _f32p), _p(res["v"], _f32p), _p(st, _u64p),
    )  # fmt: skip
    res["sign"] = np.zeros(n, np.float32)
    if stats:
        res["nodes_visited"], res["tris_tested"] = int(st[0]), int(st[1])
    return res


def query_point(points, indices, tree, queries, max_dist, stats=False):
    points, indices, targs = _tree_args(points, indices, tree)
    q = _f32(queries, (-1, 3))
    n = q.shape[0]
    res = {"result": np.zeros(n, np.uint8), "sign": np.zeros(n, np.float32), "face": np.zeros(n, np.int32),
           "u": np.zeros(n, np.float32), "v": np.zeros(n, np.float32)}  # fmt: skip
    st = np.zeros(2, np.uint64) if stats else None
    orc().orc_query_point(
        *targs, _p(q, _f32p), ctypes.c_int64(n), ctypes.c_float(max_dist),
        _p(res["result"], _u8p), _p(res["sign"], _f32p), _p(res["face"], _i32p), _p(res["u"], _f32p),
        _p(res["v"], _f32p), _p(st, _u64p),
    )  # fmt: skip
    if stats:
        res["nodes_visited"], res["tris_tested"] = int(st[0]), int(st[1])
    return res


def query_ray(points, indices, tree, starts, dirs, max_t, stats=False, roots=None):
    points, indices, targs = _tree_args(points, indices, tree)
    s = _f32(starts, (-1, 3))
    d = _f32(dirs, (-1, 3))
    n = s.shape[0]
    res = {"result": np.zeros(n, np.uint8), "sign": np.zeros(n, np.float32), "face": np.zeros(n, np.int32),
           "t": np.zeros(n, np.float32), "u": np.zeros(n, np.float32), "v": np.zeros(n, np.float32),
           "normal": np.zeros((n, 3), np.float32)}  # fmt: skip
    st = np.zeros(2, np.uint64) if stats else None
    orc().orc_query_ray(
        *targs, _p(s, _f32p), _p(d, _f32p), ctypes.c_int64(n), ctypes.c_float(max_t),
        _p(res["result"], _u8p), _p(res["sign"], _f32p), _p(res["face"], _i32p), _p(res["t"], _f32p),
        _p(res["u"], _f32p), _p(res["v"], _f32p), _p(res["normal"], _f32p), _p(st, _u64p),
        _p(None if roots is None else _i32(roots), _i32p),
    )  # fmt: skip
    if stats:
        res["nodes_visited"], res["tris_tested"] = int(st[0]), int(st[1])
    return res


def query_point_sign_parity(points, indices, tree, queries, max_dist, n_sample=1, scale=0.1, rtl=False):
    """mesh_query_point_sign_parity restatement (mesh.h:309-498, 2362-2392).  ``rtl``: draw the three direction
    offsets right to left like the g++ build of oracle/_ref (the reference's device builds draw left to right)."""
    points, indices, targs = _tree_args(points, indices, tree)
    q = _f32(queries, (-1, 3))
    n = q.shape[0]
    res = {"result": np.zeros(n, np.uint8), "sign": np.zeros(n, np.float32), "face": np.zeros(n, np.int32),
           "u": np.zeros(n, np.float32), "v": np.zeros(n, np.float32)}  # fmt: skip
    orc().orc_query_point_sign_parity(*targs, _p(q, _f32p), ctypes.c_int64(n), ctypes.c_float(max_dist), ctypes.c_int(n_sample),
                                      ctypes.c_float(scale), ctypes.c_int(1 if rtl else 0), _p(res["result"], _u8p),
                                      _p(res["sign"], _f32p), _p(res["face"], _i32p), _p(res["u"], _f32p), _p(res["v"], _f32p))  # fmt: skip
    return res


def average_edge_length(points, indices, mode=1):
    """Mesh.average_edge_length: mode 0 = the reference's CPU loop in float (mesh.cpp:140-155); mode 1 = the same
    float terms (mesh.cu:53 order) summed in double, which is what the CUDA path of this repo computes."""
    p = _f32(points, (-1, 3))
    i = _i32(indices)
    fn = orc().orc_average_edge_length
    fn.restype = ctypes.c_float
    return float(np.float32(fn(_p(p, _f32p), _p(i, _i32p), ctypes.c_int(i.size // 3), ctypes.c_int(mode))))


def query_point_sign_normal(points, indices, tree, queries, max_dist, average_edge, epsilon=1e-3):
    """mesh_query_point_sign_normal restatement (mesh.h:860-1090); ``average_edge`` = Mesh.average_edge_length."""
    points, indices, targs = _tree_args(points, indices, tree)
    q = _f32(queries, (-1, 3))
    n = q.shape[0]
    res = {"result": np.zeros(n, np.uint8), "sign": np.zeros(n, np.float32), "face": np.zeros(n, np.int32),
           "u": np.zeros(n, np.float32), "v": np.zeros(n, np.float32)}  # fmt: skip
    orc().orc_query_point_sign_normal(*targs, _p(q, _f32p), ctypes.c_int64(n), ctypes.c_float(max_dist),
                                      ctypes.c_float(average_edge), ctypes.c_float(epsilon), _p(res["result"], _u8p),
                                      _p(res["sign"], _f32p), _p(res["face"], _i32p), _p(res["u"], _f32p), _p(res["v"], _f32p))  # fmt: skip
    return res


def query_furthest_point_no_sign(points, indices, tree, queries, min_dist):
    """mesh_query_furthest_point_no_sign restatement (mesh.h:678-858)."""
    points, indices, targs = _tree_args(points, indices, tree)
    q = _f32(queries, (-1, 3))
    n = q.shape[0]
    res = {"result": np.zeros(n, np.uint8), "face": np.zeros(n, np.int32), "u": np.zeros(n, np.float32),
           "v": np.zeros(n, np.float32)}  # fmt: skip
    orc().orc_query_furthest_point_no_sign(*targs, _p(q, _f32p), ctypes.c_int64(n), ctypes.c_float(min_dist),
                                           _p(res["result"], _u8p), _p(res["face"], _i32p), _p(res["u"], _f32p),
                                           _p(res["v"], _f32p))  # fmt: skip
    return res


def mesh_face_normal(points, indices, face):
    """mesh_eval_face_normal restatement (mesh.h:2870-2888)."""
    p, i, f = _f32(points, (-1, 3)), _i32(indices), _i32(face)
    out = np.zeros((f.size, 3), np.float32)
    orc().orc_mesh_face_normal(_p(p, _f32p), _p(i, _i32p), _p(f, _i32p), ctypes.c_int64(f.size), _p(out, _f32p))
    return out


def mesh_query_sphere(points, indices, tree, centers, radii):
    """mesh_query_sphere restatement (mesh.h:2457-2737): (offsets[n+1], face indices) in iterator order; the face boxes
    of the broad phase are those of ``points`` (= mesh.lowers / uppers right after a build or refit)."""
    points, indices, targs = _tree_args(points, indices, tree)
    tlo, thi = triangle_bounds(points, indices)
    tlo, thi = _f32(tlo, (-1, 3)), _f32(thi, (-1, 3))
    c = _f32(centers, (-1, 3))
    n = c.shape[0]
    r = np.ascontiguousarray(np.broadcast_to(np.asarray(radii, np.float32), (n,)))
    offsets = np.zeros(n + 1, np.int32)
    args = (*targs, _p(tlo, _f32p), _p(thi, _f32p), _p(c, _f32p), _p(r, _f32p), ctypes.c_int64(n))
    orc().orc_mesh_query_sphere(*args, _p(offsets, _i32p), None)
    out = np.zeros(max(int(offsets[-1]), 1), np.int32)
    orc().orc_mesh_query_sphere(*args, _p(offsets, _i32p), _p(out, _i32p))
    return offsets, out[: int(offsets[-1])]


def query_ray_anyhit(points, indices, tree, starts, dirs, max_t, roots=None):
    """mesh_query_ray_anyhit restatement (mesh.h:1893-1974)."""
    points, indices, targs = _tree_args(points, indices, tree)
    s, d = _f32(starts, (-1, 3)), _f32(dirs, (-1, 3))
    out = np.zeros(s.shape[0], np.uint8)
    orc().orc_query_ray_anyhit(*targs, _p(s, _f32p), _p(d, _f32p), ctypes.c_int64(s.shape[0]), ctypes.c_float(max_t),
                               _p(out, _u8p), _p(None if roots is None else _i32(roots), _i32p))  # fmt: skip
    return out


def query_ray_count(points, indices, tree, starts, dirs, roots=None):
    """mesh_query_ray_count_intersections restatement (mesh.h:1976-2032)."""
    points, indices, targs = _tree_args(points, indices, tree)
    s, d = _f32(starts, (-1, 3)), _f32(dirs, (-1, 3))
    out = np.zeros(s.shape[0], np.int32)
    orc().orc_query_ray_count(*targs, _p(s, _f32p), _p(d, _f32p), ctypes.c_int64(s.shape[0]), _p(out, _i32p),
                              _p(None if roots is None else _i32(roots), _i32p))
    return out


def mesh_eval(attr, indices, face, u, v):
    """mesh_eval_position / mesh_eval_velocity restatement (mesh.h:2767-2805)."""
    a = _f32(attr, (-1, 3))
    idx = np.ascontiguousarray(indices, np.int32).reshape(-1)
    f = np.ascontiguousarray(face, np.int32)
    uu, vv = _f32(u), _f32(v)
    out = np.zeros((f.shape[0], 3), np.float32)
    orc().orc_mesh_eval(_p(a, _f32p), _p(idx, _i32p), _p(f, _i32p), _p(uu, _f32p), _p(vv, _f32p),
                        ctypes.c_int64(f.shape[0]), _p(out, _f32p))  # fmt: skip
    return out


def closest_point_to_triangle(a, b, c, p):
    uv = np.zeros(2, np.float32)
    orc().orc_closest_point_to_triangle(_p(_f32(a), _f32p), _p(_f32(b), _f32p), _p(_f32(c), _f32p),
                                        _p(_f32(p), _f32p), _p(uv, _f32p))  # fmt: skip
    return uv


# ----------------------------------------------------------------------------------------------
# the reference's own C++ (oracle/_ref)
# ----------------------------------------------------------------------------------------------

SAH, MEDIAN, LBVH = 0, 1, 2


class RefMesh:
    """A ``wp::Mesh`` living in host memory, queried by the reference's own header code.

    ``RefMesh(points, indices, constructor=SAH, leaf_size=4)`` runs the reference host builder
    (mesh.cpp:118-191).  ``RefMesh.from_tree(points, indices, tree)`` wraps a tree in the
    reference node layout (e.g. an LBVH from :func:`lbvh_build` or dumped from the GPU build).
    """

    def __init__(self, points, indices, constructor=SAH, leaf_size=4):
        self.points = _f32(points, (-1, 3))
        self.indices = _i32(indices).reshape(-1)
        self._own_tree = True
        self.id = ref().ref_mesh_create(
            _p(self.points, _f32p), ctypes.c_int(self.points.shape[0]), _p(self.indices, _i32p),
            ctypes.c_int(self.indices.size // 3), ctypes.c_int(constructor), ctypes.c_int(leaf_size),
        )  # fmt: skip
        if not self.id:
            raise RuntimeError("reference wp_mesh_create_host failed: " + ref().ref_get_error_string().decode())

    @classmethod
    def from_tree(cls, points, indices, tree):
        self = cls.__new__(cls)
        self.points = _f32(points, (-1, 3))
        self.indices = _i32(indices).reshape(-1)
        self._own_tree = False
        self._keep = (tree["node_lowers"], tree["node_uppers"], tree["primitive_indices"])
        self.id = ref().ref_mesh_from_tree(
            _p(self.points, _f32p), ctypes.c_int(self.points.shape[0]), _p(self.indices, _i32p),
            ctypes.c_int(self.indices.size // 3),
            tree["node_lowers"].ctypes.data_as(ctypes.c_void_p), tree["node_uppers"].ctypes.data_as(ctypes.c_void_p),
            _p(tree["primitive_indices"], _i32p), ctypes.c_int(tree["root"]),
            ctypes.c_int(tree["node_lowers"].shape[0]), ctypes.c_int(tree.get("leaf_size", 1)),
        )  # fmt: skip
        return self

    def close(self):
        if getattr(self, "id", 0):
            if self._own_tree:
                ref().ref_mesh_destroy(ctypes.c_uint64(self.id))
            else:
                ref().ref_mesh_from_tree_destroy(ctypes.c_uint64(self.id))
            self.id = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def refit(self):
        ref().ref_mesh_refit(ctypes.c_uint64(self.id))

    def tree(self):
        mx, nn, nl, root, ni = (ctypes.c_int() for _ in range(5))
        ref().ref_mesh_tree_info(ctypes.c_uint64(self.id), *(ctypes.byref(x) for x in (mx, nn, nl, root, ni)))
        t = {
            "n": ni.value, "num_nodes": nn.value, "root": root.value,
            "node_lowers": np.zeros(mx.value, HALF_DTYPE), "node_uppers": np.zeros(mx.value, HALF_DTYPE),
            "parents": np.zeros(mx.value, np.int32), "primitive_indices": np.zeros(ni.value, np.int32),
        }  # fmt: skip
        ref().ref_mesh_tree_copy(
            ctypes.c_uint64(self.id), t["node_lowers"].ctypes.data_as(ctypes.c_void_p),
            t["node_uppers"].ctypes.data_as(ctypes.c_void_p), _p(t["parents"], _i32p),
            _p(t["primitive_indices"], _i32p),
        )  # fmt: skip
        return t

    def query_point_no_sign(self, queries, max_dist, nthreads=1):
        q = _f32(queries, (-1, 3))
        n = q.shape[0]
        res = {"result": np.zeros(n, np.uint8), "face": np.zeros(n, np.int32), "u": np.zeros(n, np.float32),
               "v": np.zeros(n, np.float32), "sign": np.zeros(n, np.float32)}  # fmt: skip
        ref().ref_query_point_no_sign(
            ctypes.c_uint64(self.id), _p(q, _f32p), ctypes.c_int64(n), ctypes.c_float(max_dist),
            _p(res["result"], _u8p), _p(res["face"], _i32p), _p(res["u"], _f32p), _p(res["v"], _f32p),
            ctypes.c_int(nthreads),
        )  # fmt: skip
        return res

    def query_point(self, queries, max_dist, nthreads=1):
        q = _f32(queries, (-1, 3))
        n = q.shape[0]
        res = {"result": np.zeros(n, np.uint8), "sign": np.zeros(n, np.float32), "face": np.zeros(n, np.int32),
               "u": np.zeros(n, np.float32), "v": np.zeros(n, np.float32)}  # fmt: skip
        ref().ref_query_point(
            ctypes.c_uint64(self.id), _p(q, _f32p), ctypes.c_int64(n), ctypes.c_float(max_dist),
            _p(res["result"], _u8p), _p(res["sign"], _f32p), _p(res["face"], _i32p), _p(res["u"], _f32p),
            _p(res["v"], _f32p), ctypes.c_int(nthreads),
        )  # fmt: skip
        return res

    def query_ray(self, starts, dirs, max_t, nthreads=1, roots=None):
        s = _f32(starts, (-1, 3))
        d = _f32(dirs, (-1, 3))
        n = s.shape[0]
        res = {"result": np.zeros(n, np.uint8), "sign": np.zeros(n, np.float32), "face": np.zeros(n, np.int32),
               "t": np.zeros(n, np.float32), "u": np.zeros(n, np.float32), "v": np.zeros(n, np.float32),
               "normal": np.zeros((n, 3), np.float32)}  # fmt: skip
        ref().ref_query_ray(
            ctypes.c_uint64(self.id), _p(s, _f32p), _p(d, _f32p), ctypes.c_int64(n), ctypes.c_float(max_t),
            _p(res["result"], _u8p), _p(res["sign"], _f32p), _p(res["face"], _i32p), _p(res["t"], _f32p),
            _p(res["u"], _f32p), _p(res["v"], _f32p), _p(res["normal"], _f32p), ctypes.c_int(nthreads),
            _p(None if roots is None else _i32(roots), _i32p),
        )  # fmt: skip
        return res


    def query_point_sign_parity(self, queries, max_dist, n_sample=1, scale=0.1, nthreads=1):
        q = _f32(queries, (-1, 3))
        n = q.shape[0]
        res = {"result": np.zeros(n, np.uint8), "sign": np.zeros(n, np.float32), "face": np.zeros(n, np.int32),
               "u": np.zeros(n, np.float32), "v": np.zeros(n, np.float32)}  # fmt: skip
        ref().ref_query_point_sign_parity(ctypes.c_uint64(self.id), _p(q, _f32p), ctypes.c_int64(n), ctypes.c_float(max_dist),
                                          ctypes.c_int(n_sample), ctypes.c_float(scale), _p(res["result"], _u8p),
                                          _p(res["sign"], _f32p), _p(res["face"], _i32p), _p(res["u"], _f32p),
                                          _p(res["v"], _f32p), ctypes.c_int(nthreads))  # fmt: skip
        return res

    @property
    def average_edge_length(self):
        fn = ref().ref_mesh_get_average_edge_length
        fn.restype = ctypes.c_float
        return float(np.float32(fn(ctypes.c_uint64(self.id))))

    @average_edge_length.setter
    def average_edge_length(self, value):
        ref().ref_mesh_set_average_edge_length(ctypes.c_uint64(self.id), ctypes.c_float(value))

    def query_point_sign_normal(self, queries, max_dist, epsilon=1e-3, nthreads=1):
        q = _f32(queries, (-1, 3))
        n = q.shape[0]
        res = {"result": np.zeros(n, np.uint8), "sign": np.zeros(n, np.float32), "face": np.zeros(n, np.int32),
               "u": np.zeros(n, np.float32), "v": np.zeros(n, np.float32)}  # fmt: skip
        ref().ref_query_point_sign_normal(ctypes.c_uint64(self.id), _p(q, _f32p), ctypes.c_int64(n), ctypes.c_float(max_dist),
                                          ctypes.c_float(epsilon), _p(res["result"], _u8p), _p(res["sign"], _f32p),
                                          _p(res["face"], _i32p), _p(res["u"], _f32p), _p(res["v"], _f32p),
                                          ctypes.c_int(nthreads))  # fmt: skip
        return res

    def query_furthest_point_no_sign(self, queries, min_dist, nthreads=1):
        q = _f32(queries, (-1, 3))
        n = q.shape[0]
        res = {"result": np.zeros(n, np.uint8), "face": np.zeros(n, np.int32), "u": np.zeros(n, np.float32),
               "v": np.zeros(n, np.float32)}  # fmt: skip
        ref().ref_query_furthest_point_no_sign(ctypes.c_uint64(self.id), _p(q, _f32p), ctypes.c_int64(n),
                                               ctypes.c_float(min_dist), _p(res["result"], _u8p), _p(res["face"], _i32p),
                                               _p(res["u"], _f32p), _p(res["v"], _f32p), ctypes.c_int(nthreads))  # fmt: skip
        return res

    def eval_face_normal(self, face):
        f = _i32(face)
        out = np.zeros((f.size, 3), np.float32)
        ref().ref_mesh_eval_face_normal(ctypes.c_uint64(self.id), _p(f, _i32p), ctypes.c_int64(f.size), _p(out, _f32p))
        return out

    def query_ray_anyhit(self, starts, dirs, max_t, nthreads=1, roots=None):
        s, d = _f32(starts, (-1, 3)), _f32(dirs, (-1, 3))
        out = np.zeros(s.shape[0], np.uint8)
        ref().ref_query_ray_anyhit(ctypes.c_uint64(self.id), _p(s, _f32p), _p(d, _f32p), ctypes.c_int64(s.shape[0]),
                                   ctypes.c_float(max_t), _p(out, _u8p), ctypes.c_int(nthreads),
                                   _p(None if roots is None else _i32(roots), _i32p))  # fmt: skip
        return out

    def query_ray_count(self, starts, dirs, nthreads=1, roots=None):
        s, d = _f32(starts, (-1, 3)), _f32(dirs, (-1, 3))
        out = np.zeros(s.shape[0], np.int32)
        ref().ref_query_ray_count(ctypes.c_uint64(self.id), _p(s, _f32p), _p(d, _f32p), ctypes.c_int64(s.shape[0]),
                                  _p(out, _i32p), ctypes.c_int(nthreads), _p(None if roots is None else _i32(roots), _i32p))  # fmt: skip
        return out

    def query_aabb(self, lowers, uppers, item_bounds=None):
        """mesh_query_aabb iterator run to exhaustion per query: (offsets[n+1], indices).  ``item_bounds`` =
        (lowers, uppers) per triangle for meshes wrapped with from_tree (reference-built meshes have their own)."""
        lo, hi = _f32(lowers, (-1, 3)), _f32(uppers, (-1, 3))
        if item_bounds is not None:
            self._item_bounds = (_f32(item_bounds[0], (-1, 3)), _f32(item_bounds[1], (-1, 3)))
            ref().ref_mesh_set_bounds(ctypes.c_uint64(self.id), _p(self._item_bounds[0], _f32p), _p(self._item_bounds[1], _f32p))
        n = lo.shape[0]
        offsets = np.zeros(n + 1, np.int32)
        ref().ref_mesh_query_aabb(ctypes.c_uint64(self.id), _p(lo, _f32p), _p(hi, _f32p), ctypes.c_int64(n), _p(offsets, _i32p), None)
        indices = np.zeros(max(int(offsets[-1]), 1), np.int32)
        ref().ref_mesh_query_aabb(ctypes.c_uint64(self.id), _p(lo, _f32p), _p(hi, _f32p), ctypes.c_int64(n), _p(offsets, _i32p),
                                  _p(indices, _i32p))
        return offsets, indices[: int(offsets[-1])]

    def query_sphere(self, centers, radii, item_bounds=None):
        """mesh_query_sphere iterator run to exhaustion per query: (offsets[n+1], indices); see query_aabb."""
        c = _f32(centers, (-1, 3))
        n = c.shape[0]
        r = np.ascontiguousarray(np.broadcast_to(np.asarray(radii, np.float32), (n,)))
        if item_bounds is not None:
            self._item_bounds = (_f32(item_bounds[0], (-1, 3)), _f32(item_bounds[1], (-1, 3)))
            ref().ref_mesh_set_bounds(ctypes.c_uint64(self.id), _p(self._item_bounds[0], _f32p), _p(self._item_bounds[1], _f32p))
        offsets = np.zeros(n + 1, np.int32)
        ref().ref_mesh_query_sphere(ctypes.c_uint64(self.id), _p(c, _f32p), _p(r, _f32p), ctypes.c_int64(n), _p(offsets, _i32p), None)
        indices = np.zeros(max(int(offsets[-1]), 1), np.int32)
        ref().ref_mesh_query_sphere(ctypes.c_uint64(self.id), _p(c, _f32p), _p(r, _f32p), ctypes.c_int64(n), _p(offsets, _i32p),
                                    _p(indices, _i32p))
        return offsets, indices[: int(offsets[-1])]

    def eval(self, face, u, v, velocity=False):
        f = np.ascontiguousarray(face, np.int32)
        uu, vv = _f32(u), _f32(v)
        out = np.zeros((f.shape[0], 3), np.float32)
        ref().ref_mesh_eval(ctypes.c_uint64(self.id), ctypes.c_int(1 if velocity else 0), _p(f, _i32p), _p(uu, _f32p),
                            _p(vv, _f32p), ctypes.c_int64(f.shape[0]), _p(out, _f32p))  # fmt: skip
        return out


def ref_closest_point_to_triangle(a, b, c, p):
    uv = np.zeros(2, np.float32)
    ref().ref_closest_point_to_triangle(_p(_f32(a), _f32p), _p(_f32(b), _f32p), _p(_f32(c), _f32p),
                                        _p(_f32(p), _f32p), _p(uv, _f32p))  # fmt: skip
    return uv


def ref_morton3(x, y, z) -> int:
    return int(ref().ref_morton3_1024(x, y, z))


def ref_max_threads() -> int:
    return int(ref().ref_max_threads())


def bvh_query(tree, item_lowers, item_uppers, qa, qb, ray=False, max_dist=3.4028234663852886e38, roots=None):
    """Generic BVH query restatement (bvh.h:494-600): returns (offsets[n+1], indices) in iterator order.
    ``roots``: optional per-query start node (reference node index, -1 = tree root)."""
    lo, hi = _f32(item_lowers, (-1, 3)), _f32(item_uppers, (-1, 3))
    a, b = _f32(qa, (-1, 3)), _f32(qb, (-1, 3))
    n = a.shape[0]
    offsets = np.zeros(n + 1, np.int32)
    r = None if roots is None else _i32(roots)
    args = (tree["node_lowers"].ctypes.data_as(ctypes.c_void_p), tree["node_uppers"].ctypes.data_as(ctypes.c_void_p),
            _p(tree["primitive_indices"], _i32p), ctypes.c_int(tree["root"]), _p(lo, _f32p), _p(hi, _f32p),
            ctypes.c_int(1 if ray else 0), _p(a, _f32p), _p(b, _f32p), _p(r, _i32p), ctypes.c_int64(n),
            ctypes.c_float(max_dist))  # fmt: skip
    orc().orc_bvh_query(*args, _p(offsets, _i32p), None)
    indices = np.zeros(max(int(offsets[-1]), 1), np.int32)
    orc().orc_bvh_query(*args, _p(offsets, _i32p), _p(indices, _i32p))
    return offsets, indices[: int(offsets[-1])]


_KINDS = {"aabb": 0, "ray": 1, "sphere": 2, "capsule": 3}


def _bvh_query_kind(lib_fn, with_num_items, tree, item_lowers, item_uppers, kind, qa, qb, radii, max_dist, roots):
    lo, hi = _f32(item_lowers, (-1, 3)), _f32(item_uppers, (-1, 3))
    a = _f32(qa, (-1, 3))
    b = a if qb is None else _f32(qb, (-1, 3))
    n = a.shape[0]
    rad = None if radii is None else np.ascontiguousarray(np.broadcast_to(np.asarray(radii, np.float32), (n,)))
    offsets = np.zeros(n + 1, np.int32)
    r = None if roots is None else _i32(roots)
    args = [tree["node_lowers"].ctypes.data_as(ctypes.c_void_p), tree["node_uppers"].ctypes.data_as(ctypes.c_void_p),
            _p(tree["primitive_indices"], _i32p), ctypes.c_int(tree["root"]), _p(lo, _f32p), _p(hi, _f32p)]
    if with_num_items:
        args.append(ctypes.c_int(lo.shape[0]))
    args += [ctypes.c_int(_KINDS[kind]), _p(a, _f32p), _p(b, _f32p), _p(rad, _f32p), _p(r, _i32p), ctypes.c_int64(n),
             ctypes.c_float(max_dist)]
    lib_fn(*args, _p(offsets, _i32p), None)
    indices = np.zeros(max(int(offsets[-1]), 1), np.int32)
    lib_fn(*args, _p(offsets, _i32p), _p(indices, _i32p))
    return offsets, indices[: int(offsets[-1])]


def bvh_query_kind(tree, item_lowers, item_uppers, kind, qa, qb=None, radii=None, max_dist=3.4028234663852886e38, roots=None):
    """Generic iterator restatement for any BvhQueryKind (bvh.h:420-664): "aabb" | "ray" | "sphere" (qa = centres,
    radii) | "capsule" (qa, qb = starts, dirs; radii; closed max_dist).  Returns (offsets[n+1], indices)."""
    return _bvh_query_kind(orc().orc_bvh_query_kind, False, tree, item_lowers, item_uppers, kind, qa, qb, radii, max_dist, roots)


def ref_bvh_query_kind(tree, item_lowers, item_uppers, kind, qa, qb=None, radii=None, max_dist=3.4028234663852886e38,
                       roots=None):
    """The reference's own iterators (bvh.h:494-664, oracle/_ref) run to exhaustion per query."""
    return _bvh_query_kind(ref().ref_bvh_query_kind, True, tree, item_lowers, item_uppers, kind, qa, qb, radii, max_dist, roots)


def bvh_group_roots(tree, groups, group_ids):
    """bvh_get_group_root restatement (bvh.h:287-390): reference node index per queried group id, -1 if absent.
    ``groups`` = the per-item group array the tree was built with (None for an ungrouped tree)."""
    gid = _i32(group_ids)
    g = None if groups is None else _i32(groups)
    par = _i32(tree["parents"])
    roots = np.zeros(gid.shape[0], np.int32)
    orc().orc_bvh_group_roots(tree["node_lowers"].ctypes.data_as(ctypes.c_void_p), _p(tree["primitive_indices"], _i32p),
                              _p(par, _i32p), _p(g, _i32p), ctypes.c_int(len(tree["primitive_indices"])), _p(gid, _i32p),
                              ctypes.c_int64(gid.shape[0]), _p(roots, _i32p))  # fmt: skip
    return roots


def ref_bvh_query(tree, item_lowers, item_uppers, qa, qb, ray=False, max_dist=3.4028234663852886e38, roots=None):
    """The reference's own iterator (bvh.h:494-664, oracle/_ref) run to exhaustion per query: (offsets, indices)."""
    lo, hi = _f32(item_lowers, (-1, 3)), _f32(item_uppers, (-1, 3))
    a, b = _f32(qa, (-1, 3)), _f32(qb, (-1, 3))
    n = a.shape[0]
    offsets = np.zeros(n + 1, np.int32)
    r = None if roots is None else _i32(roots)
    args = (tree["node_lowers"].ctypes.data_as(ctypes.c_void_p), tree["node_uppers"].ctypes.data_as(ctypes.c_void_p),
            _p(tree["primitive_indices"], _i32p), ctypes.c_int(tree["root"]), _p(lo, _f32p), _p(hi, _f32p),
            ctypes.c_int(lo.shape[0]), ctypes.c_int(1 if ray else 0), _p(a, _f32p), _p(b, _f32p), _p(r, _i32p),
            ctypes.c_int64(n), ctypes.c_float(max_dist))  # fmt: skip
    ref().ref_bvh_query(*args, _p(offsets, _i32p), None)
    indices = np.zeros(max(int(offsets[-1]), 1), np.int32)
    ref().ref_bvh_query(*args, _p(offsets, _i32p), _p(indices, _i32p))
    return offsets, indices[: int(offsets[-1])]


def ref_bvh_group_roots(tree, groups, group_ids):
    """The reference's own bvh_get_group_root (bvh.h:376-390, oracle/_ref)."""
    gid = _i32(group_ids)
    g = None if groups is None else _i32(groups)
    par = _i32(tree["parents"])
    roots = np.zeros(gid.shape[0], np.int32)
    ref().ref_bvh_group_roots(tree["node_lowers"].ctypes.data_as(ctypes.c_void_p),
                              tree["node_uppers"].ctypes.data_as(ctypes.c_void_p), _p(par, _i32p),
                              _p(tree["primitive_indices"], _i32p), _p(g, _i32p), ctypes.c_int(len(tree["primitive_indices"])),
                              ctypes.c_int(tree["root"]), _p(gid, _i32p), ctypes.c_int64(gid.shape[0]), _p(roots, _i32p))  # fmt: skip
    return roots
