"""Runs the UNMODIFIED reference (NVIDIA/warp, built into baseline/_ref/warp_src by the recipe in
DESIGN.md) on the GPU box.  Test / baseline infrastructure only.

    python baseline/ref_cuda.py golden   -> gpurun_out/golden_ref_lbvh.npz   (commit as tests/golden/)
    python baseline/ref_cuda.py timing   -> gpurun_out/ref_cuda_timing.json  (numbers quoted in BASELINE.md)

`golden` builds `wp.Mesh(points, indices, bvh_constructor="lbvh")` with the reference's own CUDA
LBVH (warp/native/bvh.cu) for a few small meshes and dumps the device arrays reachable from the
descriptor at `mesh.id` (layout warp/native/mesh.h:18-35, bvh.h:176-207): primitive_indices,
node_lowers, node_uppers, node_parents, root -- the ground truth for "sorted order and hierarchy
topology bit-exact" -- plus the reference's answers to closest-point / ray queries on those meshes.
"""
import json
import os
import struct
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref", "warp_src"))
sys.path.insert(1, ROOT)

import warp as wp  # noqa: E402  (the reference)

from warp_b200 import meshgen as mg  # noqa: E402  (numpy-only generators, no native code)

wp.config.quiet = True
DEV = "cuda:0"
HALF = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("ib", "<u4")])


def raw(ptr, nbytes):
    return wp.array(ptr=ptr, dtype=wp.uint8, shape=(nbytes,), device=DEV).numpy().copy()


def dump_tree(mesh, n):
    d = raw(mesh.id, 328).tobytes()
    lowers, uppers, parents, counts, prim = struct.unpack_from("<5Q", d, 200)
    max_depth, max_nodes, num_nodes, num_leaf = struct.unpack_from("<4i", d, 240)
    (root_ptr,) = struct.unpack_from("<Q", d, 256)
    assert max_nodes == 2 * n - 1, (max_nodes, n)
    return {
        "node_lowers": raw(lowers, 16 * max_nodes).view(HALF),
        "node_uppers": raw(uppers, 16 * max_nodes).view(HALF),
        "parents": raw(parents, 4 * max_nodes).view(np.int32),
        "primitive_indices": raw(prim, 4 * n).view(np.int32),
        "root": np.int32(raw(root_ptr, 4).view(np.int32)[0]),
    }


@wp.kernel
def k_point(mesh: wp.uint64, pts: wp.array(dtype=wp.vec3), max_dist: float, res: wp.array(dtype=wp.int32),
            sign: wp.array(dtype=float), face: wp.array(dtype=wp.int32), u: wp.array(dtype=float), v: wp.array(dtype=float)):
    tid = wp.tid()
    q = wp.mesh_query_point(mesh, pts[tid], max_dist)
    res[tid] = wp.where(q.result, 1, 0)
    sign[tid] = q.sign
    face[tid] = q.face
    u[tid] = q.u
    v[tid] = q.v


@wp.kernel
def k_point_no_sign(mesh: wp.uint64, pts: wp.array(dtype=wp.vec3), max_dist: float, res: wp.array(dtype=wp.int32),
                    face: wp.array(dtype=wp.int32), u: wp.array(dtype=float), v: wp.array(dtype=float)):
    tid = wp.tid()
    q = wp.mesh_query_point_no_sign(mesh, pts[tid], max_dist)
    res[tid] = wp.where(q.result, 1, 0)
    face[tid] = q.face
    u[tid] = q.u
    v[tid] = q.v


@wp.kernel
def k_ray(mesh: wp.uint64, starts: wp.array(dtype=wp.vec3), dirs: wp.array(dtype=wp.vec3), max_t: float,
          res: wp.array(dtype=wp.int32), sign: wp.array(dtype=float), face: wp.array(dtype=wp.int32),
          t: wp.array(dtype=float), u: wp.array(dtype=float), v: wp.array(dtype=float), nrm: wp.array(dtype=wp.vec3)):
    tid = wp.tid()
    q = wp.mesh_query_ray(mesh, starts[tid], dirs[tid], max_t)
    res[tid] = wp.where(q.result, 1, 0)
    sign[tid] = q.sign
    face[tid] = q.face
    t[tid] = q.t
    u[tid] = q.u
    v[tid] = q.v
    nrm[tid] = q.normal


def ref_point(mesh, Q, max_dist, sign=True):
    n = len(Q)
    pts = wp.array(Q, dtype=wp.vec3, device=DEV)
    res, face = wp.zeros(n, dtype=wp.int32, device=DEV), wp.zeros(n, dtype=wp.int32, device=DEV)
    sg, u, v = (wp.zeros(n, dtype=float, device=DEV) for _ in range(3))
    if sign:
        wp.launch(k_point, dim=n, inputs=[mesh.id, pts, max_dist], outputs=[res, sg, face, u, v], device=DEV)
    else:
        wp.launch(k_point_no_sign, dim=n, inputs=[mesh.id, pts, max_dist], outputs=[res, face, u, v], device=DEV)
    return {"result": res.numpy().astype(np.uint8), "sign": sg.numpy(), "face": face.numpy(), "u": u.numpy(), "v": v.numpy()}


def ref_ray(mesh, S, D, max_t):
    n = len(S)
    s, d = wp.array(S, dtype=wp.vec3, device=DEV), wp.array(D, dtype=wp.vec3, device=DEV)
    res, face = wp.zeros(n, dtype=wp.int32, device=DEV), wp.zeros(n, dtype=wp.int32, device=DEV)
    sg, t, u, v = (wp.zeros(n, dtype=float, device=DEV) for _ in range(4))
    nrm = wp.zeros(n, dtype=wp.vec3, device=DEV)
    wp.launch(k_ray, dim=n, inputs=[mesh.id, s, d, max_t], outputs=[res, sg, face, t, u, v, nrm], device=DEV)
    return {"result": res.numpy().astype(np.uint8), "sign": sg.numpy(), "face": face.numpy(), "t": t.numpy(),
            "u": u.numpy(), "v": v.numpy(), "normal": nrm.numpy()}


def make_mesh(P, I, leaf):
    return wp.Mesh(wp.array(P, dtype=wp.vec3, device=DEV), wp.array(I, dtype=wp.int32, device=DEV),
                   bvh_constructor="lbvh", bvh_leaf_size=leaf)


def golden():
    out = {}
    cases = {
        "cube": (mg.CUBE_POINTS, mg.CUBE_INDICES_RH),
        "ico2": mg.noisy_sphere(2, 0.05, 11),
        "ico4": mg.noisy_sphere(4, 0.03, 1),
        "height33": mg.heightfield(33),
        "cloth40": mg.cloth(40, frame=3),
    }
    # degenerate: many coincident triangles -> equal keys, parity ladders, depth >= 32 rule
    Pd = np.tile(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), (200, 1))
    Pd[300:] += np.array([3, 0, 0], np.float32)
    cases["dups"] = (Pd, np.arange(600, dtype=np.int32))
    for name, (P, I) in cases.items():
        out[f"{name}_points"], out[f"{name}_indices"] = P, I
        for leaf in (1, 4):
            m = make_mesh(P, I, leaf)
            wp.synchronize()
            for k, v in dump_tree(m, len(I) // 3).items():
                out[f"{name}_leaf{leaf}_{k}"] = v
            if name in ("cube", "ico2", "height33"):
                Q = mg.box_queries(P, 256, seed=12)
                S, D = mg.random_rays(P, 256, seed=13)
                out[f"{name}_queries"], out[f"{name}_ray_starts"], out[f"{name}_ray_dirs"] = Q, S, D
                for k, v in ref_point(m, Q, 1.0e6).items():
                    out[f"{name}_leaf{leaf}_point_{k}"] = v
                for k, v in ref_ray(m, S, D, 1.0e6).items():
                    out[f"{name}_leaf{leaf}_ray_{k}"] = v
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "golden_ref_lbvh.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes", len(out), "arrays")


def timed(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    wp.synchronize()
    ts = []
    for _ in range(reps):
        wp.synchronize()
        t0 = time.perf_counter()
        fn()
        wp.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
    return float(np.median(ts)), float(np.min(ts))


def timing_c2_c3():
    """The part bench.py runs in-line (`extra.reference_cuda`): C2 build / refit / closest point and C3 rays with the
    reference's own CUDA path on the same GPU, in the same run.  Prints one JSON line."""
    res = {"warp_version": wp.__version__, "device": wp.get_device(DEV).name, "note": "unmodified reference (NVIDIA/warp "
           "built by its own build_lib.py --quick, sm_100 SASS), kernels of asv/benchmarks/spatial_query.py, default module options"}
    P, I = mg.noisy_sphere(8, 0.02, 1)
    pts = wp.array(P, dtype=wp.vec3, device=DEV)
    idx = wp.array(I, dtype=wp.int32, device=DEV)
    holder = {}

    def build():
        holder["m"] = wp.Mesh(pts, idx, bvh_constructor="lbvh")

    res["c2_triangles"] = len(I) // 3
    res["c2_build_ms_median"], res["c2_build_ms_min"] = timed(build, 10)
    m = holder["m"]
    res["c2_refit_ms_median"], res["c2_refit_ms_min"] = timed(m.refit, 20)
    nq = 1 << 24
    q = wp.array(mg.box_queries(P, nq, seed=2), dtype=wp.vec3, device=DEV)
    r, f = wp.zeros(nq, dtype=wp.int32, device=DEV), wp.zeros(nq, dtype=wp.int32, device=DEV)
    sg, u, v = (wp.zeros(nq, dtype=float, device=DEV) for _ in range(3))
    ms, mn = timed(lambda: wp.launch(k_point_no_sign, dim=nq, inputs=[m.id, q, 1.0e6], outputs=[r, f, u, v], device=DEV), 5)
    res["c2_point_no_sign_ms_median"], res["c2_point_no_sign_qps"] = ms, nq / (ms * 1e-3)
    ms, mn = timed(lambda: wp.launch(k_point, dim=nq // 8, inputs=[m.id, q, 1.0e6], outputs=[r, sg, f, u, v], device=DEV), 3)
    res["c2_point_sign_qps_on_2M_sample"] = (nq // 8) / (ms * 1e-3)
    del m, holder, q
    Ph, Ih = mg.heightfield(2237, 4)
    hm = wp.Mesh(wp.array(Ph, dtype=wp.vec3, device=DEV), wp.array(Ih, dtype=wp.int32, device=DEV), bvh_constructor="lbvh")
    S, D = mg.pinhole_rays(4096, 4096)
    n = len(S)
    s, d = wp.array(S, dtype=wp.vec3, device=DEV), wp.array(D, dtype=wp.vec3, device=DEV)
    r, f = wp.zeros(n, dtype=wp.int32, device=DEV), wp.zeros(n, dtype=wp.int32, device=DEV)
    sg, t, u, v = (wp.zeros(n, dtype=float, device=DEV) for _ in range(4))
    nrm = wp.zeros(n, dtype=wp.vec3, device=DEV)
    ms, mn = timed(lambda: wp.launch(k_ray, dim=n, inputs=[hm.id, s, d, 1.0e6], outputs=[r, sg, f, t, u, v, nrm], device=DEV), 5)
    res["c3_triangles"] = len(Ih) // 3
    res["c3_ray_ms_median"], res["c3_rays_per_s"] = ms, n / (ms * 1e-3)
    print("REF_CUDA " + json.dumps(res), flush=True)


def timing():
    res = {"warp_version": wp.__version__, "device": wp.get_device(DEV).name, "note": "reference built with "
           "build_lib.py --quick patched to emit sm_100 SASS only (see DESIGN.md); default module options"}
    P, I = mg.noisy_sphere(8, 0.02, 1)
    T = len(I) // 3
    pts = wp.array(P, dtype=wp.vec3, device=DEV)
    idx = wp.array(I, dtype=wp.int32, device=DEV)
    holder = {}

    def build():
        holder["m"] = wp.Mesh(pts, idx, bvh_constructor="lbvh")

    res["c2_triangles"] = T
    res["c2_build_ms_median"], res["c2_build_ms_min"] = timed(build, 10)
    m = holder["m"]
    res["c2_refit_ms_median"], res["c2_refit_ms_min"] = timed(m.refit, 20)
    nq = 1 << 24
    Q = mg.box_queries(P, nq, seed=2)
    q = wp.array(Q, dtype=wp.vec3, device=DEV)
    r, f = wp.zeros(nq, dtype=wp.int32, device=DEV), wp.zeros(nq, dtype=wp.int32, device=DEV)
    sg, u, v = (wp.zeros(nq, dtype=float, device=DEV) for _ in range(3))
    ms, mn = timed(lambda: wp.launch(k_point_no_sign, dim=nq, inputs=[m.id, q, 1.0e6], outputs=[r, f, u, v], device=DEV), 5)
    res["c2_point_no_sign_ms_median"], res["c2_point_no_sign_qps"] = ms, nq / (mn * 1e-3)
    ms, mn = timed(lambda: wp.launch(k_point, dim=nq // 8, inputs=[m.id, q, 1.0e6], outputs=[r, sg, f, u, v], device=DEV), 3)
    res["c2_point_sign_qps_on_2M_sample"] = (nq // 8) / (mn * 1e-3)
    del m, holder
    # C3 rays
    Ph, Ih = mg.heightfield(2237, 4)
    hm = wp.Mesh(wp.array(Ph, dtype=wp.vec3, device=DEV), wp.array(Ih, dtype=wp.int32, device=DEV), bvh_constructor="lbvh")
    S, D = mg.pinhole_rays(4096, 4096)
    n = len(S)
    s, d = wp.array(S, dtype=wp.vec3, device=DEV), wp.array(D, dtype=wp.vec3, device=DEV)
    r, f = wp.zeros(n, dtype=wp.int32, device=DEV), wp.zeros(n, dtype=wp.int32, device=DEV)
    sg, t, u, v = (wp.zeros(n, dtype=float, device=DEV) for _ in range(4))
    nrm = wp.zeros(n, dtype=wp.vec3, device=DEV)
    ms, mn = timed(lambda: wp.launch(k_ray, dim=n, inputs=[hm.id, s, d, 1.0e6], outputs=[r, sg, f, t, u, v, nrm], device=DEV), 5)
    res["c3_triangles"] = len(Ih) // 3
    res["c3_ray_ms_median"], res["c3_rays_per_s"] = ms, n / (mn * 1e-3)
    res["c3_hit_fraction"] = float(r.numpy().mean())
    del hm, s, d, r, f, sg, t, u, v, nrm
    # C4: 4 M-triangle cloth, per frame = refit + 8.4 M closest-point queries within 0.05
    n = 1415
    Pc, Ic = mg.cloth(n, 0)
    cp = wp.array(Pc, dtype=wp.vec3, device=DEV)
    cm = wp.Mesh(cp, wp.array(Ic, dtype=wp.int32, device=DEV), bvh_constructor="lbvh")
    nq = 1 << 23
    r, f = wp.zeros(nq, dtype=wp.int32, device=DEV), wp.zeros(nq, dtype=wp.int32, device=DEV)
    u, v = wp.zeros(nq, dtype=float, device=DEV), wp.zeros(nq, dtype=float, device=DEV)
    frames, prev = [], Pc
    for fr in range(1, 13):
        Pf, _ = mg.cloth(n, fr)
        rng = np.random.default_rng(5 + fr)
        Q = (prev[rng.integers(0, prev.shape[0], nq)] + rng.normal(0, 0.01, (nq, 3))).astype(np.float32)
        q = wp.array(Q, dtype=wp.vec3, device=DEV)
        cp.assign(Pf)
        wp.synchronize()
        t0 = time.perf_counter()
        cm.refit()
        wp.launch(k_point_no_sign, dim=nq, inputs=[cm.id, q, 0.05], outputs=[r, f, u, v], device=DEV)
        wp.synchronize()
        frames.append(1e3 * (time.perf_counter() - t0))
        prev = Pf
    res["c4_triangles"] = len(Ic) // 3
    res["c4_frame_ms_median"] = float(np.median(frames[2:]))
    res["c4_frame_ms_all"] = frames
    del cm, cp, q
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "ref_cuda_timing.json"), "w"), indent=1)  # partial, in case C5 fails
    # C5: 100 M-triangle heightfield, build + a 16.8 M-query sample of the 125 M-query shard
    P5, I5 = mg.heightfield(7072, 4)
    p5, i5 = wp.array(P5, dtype=wp.vec3, device=DEV), wp.array(I5, dtype=wp.int32, device=DEV)
    wp.synchronize()
    t0 = time.perf_counter()
    m5 = wp.Mesh(p5, i5, bvh_constructor="lbvh")
    wp.synchronize()
    res["c5_triangles"] = len(I5) // 3
    res["c5_build_ms_first"] = 1e3 * (time.perf_counter() - t0)
    res["c5_refit_ms_median"], _ = timed(m5.refit, 3, warm=1)
    rng = np.random.default_rng(6)
    lo, hi = P5.min(0), P5.max(0)
    c, h = 0.5 * (lo + hi), 0.6 * (hi - lo)
    nq = 1 << 24
    Q = (c + (rng.random((nq, 3), dtype=np.float32) * 2 - 1) * h).astype(np.float32)
    q = wp.array(Q, dtype=wp.vec3, device=DEV)
    r, f = wp.zeros(nq, dtype=wp.int32, device=DEV), wp.zeros(nq, dtype=wp.int32, device=DEV)
    u, v = wp.zeros(nq, dtype=float, device=DEV), wp.zeros(nq, dtype=float, device=DEV)
    ms, mn = timed(lambda: wp.launch(k_point_no_sign, dim=nq, inputs=[m5.id, q, 1.0e6], outputs=[r, f, u, v], device=DEV), 2, warm=1)
    res["c5_point_no_sign_ms_16M_sample"], res["c5_point_no_sign_qps"] = ms, nq / (mn * 1e-3)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", "ref_cuda_timing.json")
    json.dump(res, open(path, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    {"golden": golden, "timing": timing, "timing_c2_c3": timing_c2_c3}[sys.argv[1]]()
