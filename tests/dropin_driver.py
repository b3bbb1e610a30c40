"""Drop-in proof, run as a subprocess by tests/test_gpu_dropin.py on a GPU box:

    python tests/dropin_driver.py <path to the built reference (baseline/_ref/warp_src)> <libwarp_b200.so>

Imports the UNMODIFIED reference Warp, installs the stub of INTEGRATION.md section 1 verbatim (extracted from the
markdown), and runs the LBVH cases of the reference's own tests through it -- warp/tests/geometry/test_mesh.py:111-357
(unit-cube closest point / ray golden values, leaf sizes, groups, refit, points setter) and
warp/tests/geometry/test_bvh.py:186-262, 423-511 (generic queries against brute force across refit / rebuild) -- with
unmodified @wp.kernel code that reaches the tree only through mesh.id / bvh.id.  Prints one JSON line.
"""
import json
import os
import re
import sys

ref_root, lib_path = sys.argv[1], sys.argv[2]
sys.path.insert(0, ref_root)
repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

import numpy as np  # noqa: E402
import warp as wp  # noqa: E402  (the reference)

wp.config.quiet = True
wp.init()
import warp._src.context as wctx  # noqa: E402

md = open(os.path.join(repo, "INTEGRATION.md")).read()
section = md[md.index("## 1. The stub"):md.index("## 2. ")]
stub_src = re.search(r"```python\n(.*?)```", section, re.S).group(1)
ns = {}
exec(compile(stub_src, "INTEGRATION.md#1", "exec"), ns)
b200 = ns["install_b200"](wctx.runtime, lib_path)

import ctypes  # noqa: E402

b200.wp_b200_bvh_info.argtypes = [ctypes.c_uint64, ctypes.c_void_p]
b200.wp_b200_bvh_info.restype = ctypes.c_int


def is_ours(id_):
    buf = (ctypes.c_int * 16)()
    return bool(b200.wp_b200_bvh_info(ctypes.c_uint64(id_), buf))


POINTS = np.array([(0.5, -0.5, 0.5), (-0.5, -0.5, 0.5), (0.5, 0.5, 0.5), (-0.5, 0.5, 0.5), (-0.5, -0.5, -0.5),
                   (0.5, -0.5, -0.5), (-0.5, 0.5, -0.5), (0.5, 0.5, -0.5)], dtype=np.float32)  # fmt: skip
RH = np.array([0, 3, 1, 0, 2, 3, 4, 7, 5, 4, 6, 7, 6, 2, 7, 6, 3, 2, 5, 1, 4, 5, 0, 1, 5, 2, 0, 5, 7, 2, 1, 6, 4, 1, 3, 6], dtype=np.int32)
LH = np.array([0, 1, 3, 0, 3, 2, 4, 5, 7, 4, 7, 6, 6, 7, 2, 6, 2, 3, 5, 4, 1, 5, 1, 0, 5, 0, 2, 5, 2, 7, 1, 4, 6, 1, 6, 3], dtype=np.int32)


@wp.kernel(enable_backward=False)
def k_point(mesh_id: wp.uint64, pts: wp.array(dtype=wp.vec3), max_dist: float, res: wp.array(dtype=wp.int32),
            sign: wp.array(dtype=float), face: wp.array(dtype=wp.int32), u: wp.array(dtype=float), v: wp.array(dtype=float),
            pos: wp.array(dtype=wp.vec3)):
    i = wp.tid()
    s = float(0.0)
    f = int(0)
    bu = float(0.0)
    bv = float(0.0)
    ok = wp.mesh_query_point(mesh_id, pts[i], max_dist, s, f, bu, bv)
    res[i] = wp.where(ok, 1, 0)
    sign[i] = s
    face[i] = f
    u[i] = bu
    v[i] = bv
    pos[i] = wp.mesh_eval_position(mesh_id, f, bu, bv)


@wp.kernel(enable_backward=False)
def k_ray(mesh_id: wp.uint64, starts: wp.array(dtype=wp.vec3), dirs: wp.array(dtype=wp.vec3), max_t: float,
          res: wp.array(dtype=wp.int32), t: wp.array(dtype=float), face: wp.array(dtype=wp.int32), sign: wp.array(dtype=float)):
    i = wp.tid()
    tt = float(0.0)
    bu = float(0.0)
    bv = float(0.0)
    s = float(0.0)
    n = wp.vec3()
    f = int(0)
    ok = wp.mesh_query_ray(mesh_id, starts[i], dirs[i], max_t, tt, bu, bv, s, n, f)
    res[i] = wp.where(ok, 1, 0)
    t[i] = tt
    face[i] = f
    sign[i] = s


@wp.kernel(enable_backward=False)
def k_mesh_aabb(mesh_id: wp.uint64, lo: wp.array(dtype=wp.vec3), hi: wp.array(dtype=wp.vec3), counts: wp.array(dtype=wp.int32),
                sums: wp.array(dtype=wp.int32)):
    i = wp.tid()
    q = wp.mesh_query_aabb(mesh_id, lo[i], hi[i])
    f = int(0)
    c = int(0)
    s = int(0)
    while wp.mesh_query_aabb_next(q, f):
        c += 1
        s += f
    counts[i] = c
    sums[i] = s


@wp.kernel(enable_backward=False)
def k_bvh_aabb(bvh_id: wp.uint64, lo: wp.vec3, hi: wp.vec3, hits: wp.array(dtype=wp.int32)):
    q = wp.bvh_query_aabb(bvh_id, lo, hi)
    b = int(0)
    while wp.bvh_query_next(q, b):
        hits[b] = 1


@wp.kernel(enable_backward=False)
def k_bvh_ray(bvh_id: wp.uint64, start: wp.vec3, dir: wp.vec3, hits: wp.array(dtype=wp.int32)):
    q = wp.bvh_query_ray(bvh_id, start, dir)
    b = int(0)
    while wp.bvh_query_next(q, b):
        hits[b] = 1


dev = "cuda:0"
report = {"cases": {}, "routed": {}}


def case(name):
    def deco(fn):
        try:
            fn()
            report["cases"][name] = "ok"
        except Exception as e:  # noqa: BLE001
            report["cases"][name] = f"FAIL: {type(e).__name__}: {e}"
        return fn
    return deco


def query_points(mesh, P):
    n = len(P)
    pts = wp.array(P, dtype=wp.vec3, device=dev)
    res, face = wp.zeros(n, dtype=wp.int32, device=dev), wp.zeros(n, dtype=wp.int32, device=dev)
    sign, u, v = (wp.zeros(n, dtype=float, device=dev) for _ in range(3))
    pos = wp.zeros(n, dtype=wp.vec3, device=dev)
    wp.launch(k_point, dim=n, inputs=[mesh.id, pts, 1.0e6, res, sign, face, u, v, pos], device=dev)
    return res.numpy(), sign.numpy(), face.numpy(), u.numpy(), v.numpy(), pos.numpy()


def query_rays(mesh, S, D):
    n = len(S)
    s, d = wp.array(S, dtype=wp.vec3, device=dev), wp.array(D, dtype=wp.vec3, device=dev)
    res, face = wp.zeros(n, dtype=wp.int32, device=dev), wp.zeros(n, dtype=wp.int32, device=dev)
    t, sign = wp.zeros(n, dtype=float, device=dev), wp.zeros(n, dtype=float, device=dev)
    wp.launch(k_ray, dim=n, inputs=[mesh.id, s, d, 1.0e6, res, t, face, sign], device=dev)
    return res.numpy(), t.numpy(), face.numpy(), sign.numpy()


@case("mesh_query_point golden (test_mesh.py:111-147)")
def _():
    pts = wp.array(POINTS, dtype=wp.vec3, device=dev)
    for idx, want_sign in ((RH, -1.0), (LH, 1.0)):
        mesh = wp.Mesh(points=pts, indices=wp.array(idx, dtype=int, device=dev), bvh_constructor="lbvh")
        assert is_ours(mesh.id), "mesh was not built by libwarp_b200.so"
        res, sign, face, u, v, pos = query_points(mesh, np.array([[0.1, 0.2, 0.3]], np.float32))
        assert res[0] == 1 and face[0] == 1 and np.sign(sign[0]) == want_sign, (res, face, sign)
        assert np.linalg.norm(pos[0] - np.array([0.1, 0.2, 0.5])) < 1e-6, pos
    report["routed"]["mesh"] = True


@case("mesh_query_ray golden, leaf sizes 1/2/4, default constructor (test_mesh.py:150-262)")
def _():
    pts = wp.array(POINTS, dtype=wp.vec3, device=dev)
    d = np.array([-1.2, 2.3, -3.4])
    d = (d / np.linalg.norm(d)).astype(np.float32)
    for leaf in (1, 2, 4):
        for idx, want_sign, kw in ((RH, -1.0, {"bvh_constructor": "lbvh", "bvh_leaf_size": leaf}), (LH, 1.0, {})):
            mesh = wp.Mesh(points=pts, indices=wp.array(idx, dtype=int, device=dev), **kw)
            assert is_ours(mesh.id)
            res, t, face, sign = query_rays(mesh, np.array([[0.1, 0.2, 0.3]], np.float32), d[None])
            assert res[0] == 1 and face[0] == 4 and abs(t[0] - 0.557828) < 1e-5 and np.sign(sign[0]) == want_sign, (res, t, face, sign)


@case("grouped mesh ray (test_mesh.py:265-293)")
def _():
    pts = wp.array(POINTS, dtype=wp.vec3, device=dev)
    idx = wp.array(RH, dtype=int, device=dev)
    d = np.array([-1.2, 2.3, -3.4])
    d = (d / np.linalg.norm(d)).astype(np.float32)
    g1 = np.ones(12, np.int32)
    g1[:6] = 0
    for leaf in (1, 2, 4):
        for groups in (None, np.zeros(12, np.int32), g1):
            g = None if groups is None else wp.array(groups, dtype=int, device=dev)
            mesh = wp.Mesh(points=pts, indices=idx, groups=g, bvh_constructor="lbvh", bvh_leaf_size=leaf)
            assert is_ours(mesh.id)
            res, t, face, sign = query_rays(mesh, np.array([[0.1, 0.2, 0.3]], np.float32), d[None])
            assert res[0] == 1 and abs(t[0] - 0.557828) < 1e-5, (leaf, res, t)


@case("mesh refit + points setter (test_mesh.py:296-357)")
def _():
    P = POINTS.copy()
    pts = wp.array(P, dtype=wp.vec3, device=dev)
    mesh = wp.Mesh(points=pts, indices=wp.array(RH, dtype=int, device=dev), bvh_constructor="lbvh")
    o, d = np.array([[0.0, 5.0, 0.0]], np.float32), np.array([[0.0, -1.0, 0.0]], np.float32)
    assert query_rays(mesh, o, d)[0][0] == 1
    P2 = P + np.array([10.0, 0.0, 0.0], np.float32)
    wp.copy(pts, wp.array(P2, dtype=wp.vec3, device=dev))
    mesh.refit()
    assert query_rays(mesh, o + np.array([10.0, 0, 0], np.float32), d)[0][0] == 1, "miss at the moved location after refit"
    assert query_rays(mesh, o, d)[0][0] == 0, "hit at the old location after refit"
    mesh.points = wp.array(P, dtype=wp.vec3, device=dev)  # setter: swaps the array and refits (types.py:6276-6296)
    assert query_rays(mesh, o, d)[0][0] == 1, "miss after the points setter moved the mesh back"


@case("random mesh: unmodified Warp kernels == batched B200 queries; mesh_query_aabb == brute force")
def _():
    rng = np.random.default_rng(11)
    nv, nt = 400, 900
    P = rng.random((nv, 3)).astype(np.float32)
    I = rng.integers(0, nv, (nt, 3)).astype(np.int32)
    I = I[(I[:, 0] != I[:, 1]) & (I[:, 1] != I[:, 2]) & (I[:, 0] != I[:, 2])]
    nt = len(I)
    pts = wp.array(P, dtype=wp.vec3, device=dev)
    mesh = wp.Mesh(points=pts, indices=wp.array(I.reshape(-1), dtype=int, device=dev))
    assert is_ours(mesh.id)
    Q = (rng.random((2000, 3)) * 1.4 - 0.2).astype(np.float32)
    res, sign, face, u, v, pos = query_points(mesh, Q)
    # the library's own batched query on the same tree (device pointers from Warp arrays)
    qd = wp.array(Q, dtype=wp.vec3, device=dev)
    r8 = wp.zeros(len(Q), dtype=wp.uint8, device=dev)
    f2 = wp.zeros(len(Q), dtype=wp.int32, device=dev)
    s2, u2, v2 = (wp.zeros(len(Q), dtype=float, device=dev) for _ in range(3))
    fn = b200.wp_b200_mesh_query_point
    fn.argtypes = [ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_float] + [ctypes.c_void_p] * 5
    fn.restype = ctypes.c_int
    wp.synchronize_device(dev)
    assert fn(mesh.id, qd.ptr, len(Q), 1.0e6, r8.ptr, s2.ptr, f2.ptr, u2.ptr, v2.ptr)
    wp.synchronize_device(dev)
    assert (r8.numpy() == res).all()
    same = f2.numpy() == face
    # NVRTC contracts to FMAs (--fmad=true), the library does not: faces may differ only at equal-distance ties
    A, B, C = P[I[:, 0]], P[I[:, 1]], P[I[:, 2]]
    fo = f2.numpy()
    p_ours = u2.numpy()[:, None] * A[fo] + v2.numpy()[:, None] * B[fo] + (1 - u2.numpy() - v2.numpy())[:, None] * C[fo]
    assert np.abs(np.linalg.norm(p_ours - Q, axis=1) - np.linalg.norm(pos - Q, axis=1)).max() < 1e-5
    assert same.mean() > 0.9, same.mean()
    assert (np.sign(s2.numpy()) == np.sign(sign)).mean() > 0.99
    report["face_agreement_vs_nvrtc_kernel"] = float(same.mean())
    # mesh_query_aabb through Mesh::lowers / uppers of the descriptor
    qlo = (rng.random((300, 3)) * 0.8).astype(np.float32)
    qhi = qlo + (rng.random((300, 3)) * 0.3).astype(np.float32)
    counts, sums = wp.zeros(300, dtype=wp.int32, device=dev), wp.zeros(300, dtype=wp.int32, device=dev)
    wp.launch(k_mesh_aabb, dim=300, inputs=[mesh.id, wp.array(qlo, dtype=wp.vec3, device=dev), wp.array(qhi, dtype=wp.vec3, device=dev), counts, sums], device=dev)
    tlo, thi = np.minimum(np.minimum(A, B), C), np.maximum(np.maximum(A, B), C)
    ov = ((tlo[None] <= qhi[:, None]) & (thi[None] >= qlo[:, None])).all(-1)
    assert (counts.numpy() == ov.sum(1)).all(), "mesh_query_aabb hit counts differ from brute force"
    assert (sums.numpy() == (ov * np.arange(nt)[None]).sum(1)).all(), "mesh_query_aabb hit sets differ from brute force"


@case("wp.Bvh aabb / ray vs brute force across refit and rebuild (test_bvh.py:186-262)")
def _():
    rng = np.random.default_rng(123)
    n = 100
    for leaf in (1, 2, 4):
        lowers = rng.random((n, 3)) * 5.0
        uppers = lowers + rng.random((n, 3)) * 5.0
        dl, du = wp.array(lowers, dtype=wp.vec3, device=dev), wp.array(uppers, dtype=wp.vec3, device=dev)
        bvh = wp.Bvh(dl, du, leaf_size=leaf)
        assert is_ours(bvh.id)
        report["routed"]["bvh"] = True
        qlo, qhi = np.array([2.0, 2.0, 2.0]), np.array([8.0, 8.0, 8.0])
        start, d = np.zeros(3), np.ones(3) / np.sqrt(3.0)
        for step in range(3):
            hits = wp.zeros(n, dtype=wp.int32, device=dev)
            wp.launch(k_bvh_aabb, dim=1, inputs=[bvh.id, wp.vec3(*qlo), wp.vec3(*qhi), hits], device=dev)
            lo32, hi32 = lowers.astype(np.float32), uppers.astype(np.float32)
            want = ((lo32 <= qhi) & (hi32 >= qlo)).all(1)
            assert (hits.numpy().astype(bool) == want).all(), f"aabb hits differ (leaf {leaf}, step {step})"
            hits.zero_()
            wp.launch(k_bvh_ray, dim=1, inputs=[bvh.id, wp.vec3(*start), wp.vec3(*d), hits], device=dev)
            rcp = 1.0 / d
            l1, l2 = (lo32 - start) * rcp, (hi32 - start) * rcp
            lmin, lmax = np.minimum(l1, l2).max(1), np.maximum(l1, l2).min(1)
            want = (lmax >= 0) & (lmax >= lmin)
            assert (hits.numpy().astype(bool) == want).all(), f"ray hits differ (leaf {leaf}, step {step})"
            lowers = rng.random((n, 3)) * 5.0
            uppers = lowers + rng.random((n, 3)) * 5.0
            wp.copy(dl, wp.array(lowers, dtype=wp.vec3, device=dev))
            wp.copy(du, wp.array(uppers, dtype=wp.vec3, device=dev))
            if step == 0:
                bvh.refit()
            else:
                bvh.rebuild()


@case("non-LBVH constructors stay with the reference library")
def _():
    pts = wp.array(POINTS, dtype=wp.vec3, device=dev)
    mesh = wp.Mesh(points=pts, indices=wp.array(RH, dtype=int, device=dev), bvh_constructor="sah")
    assert not is_ours(mesh.id)
    res, sign, face, u, v, pos = query_points(mesh, np.array([[0.1, 0.2, 0.3]], np.float32))
    assert res[0] == 1 and face[0] == 1
    mesh.refit()
    del mesh


wp.synchronize_device(dev)
report["ok"] = all(v == "ok" for v in report["cases"].values())
print("DROPIN_REPORT " + json.dumps(report))
sys.exit(0 if report["ok"] else 1)
