"""Measures this repo on every BASELINE.json config (C2..C5) on one GPU and writes gpurun_out/configs_ours.json.

    python scripts/configs_run.py [c2 c3 c4 c5]

C1 is the CPU-reference config (bench.py --impl reference / cpu_baseline).  Numbers land in BASELINE.md §3.5.
"""
import json, os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import warp_b200 as wp
from warp_b200 import _lib, meshgen as mg
from bench import event_ms

core = _lib.core()
stream = core.wp_cuda_context_get_stream(None)
which = [a.lower() for a in sys.argv[1:]] or ["c2", "c3", "c4", "c5"]
res = {}


def sync():
    core.wp_cuda_context_synchronize(None)


def timed_ms(fn, reps=5, warm=1):
    for _ in range(warm):
        fn()
    sync()
    return statistics.median([event_ms(core, fn, stream) for _ in range(reps)])


def build_times(pts, idx, **kw):
    sync()
    t0 = time.perf_counter(); m = wp.Mesh(pts, idx, **kw); sync(); ctor = 1e3 * (time.perf_counter() - t0)
    ctors = []
    for _ in range(3):
        sync(); t0 = time.perf_counter(); m2 = wp.Mesh(pts, idx, **kw); sync(); ctors.append(1e3 * (time.perf_counter() - t0)); del m2
    rebuild = timed_ms(lambda: core.wp_b200_mesh_rebuild_device(m.id), 5)
    refit = timed_ms(m.refit, 5)
    return m, {"first_constructor_ms": ctor, "constructor_ms": statistics.median(ctors), "build_kernels_ms": rebuild, "refit_ms": refit}


if "c2" in which:
    P, I = mg.noisy_sphere(8, 0.02, 1)
    pts, idx = wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32)
    m, r = build_times(pts, idx)
    nq = 1 << 24
    q = wp.array(mg.box_queries(P, nq, seed=2), dtype=wp.vec3)
    out = wp.mesh_query_point_no_sign(m, q, 1e6)
    ms = timed_ms(lambda: wp.mesh_query_point_no_sign(m, q, 1e6, out=out), 5)
    r.update(triangles=len(I) // 3, no_sign_ms=ms, no_sign_qps=nq / ms * 1e3)
    ns = 1 << 21
    q2 = wp.array(mg.box_queries(P, ns, seed=2), dtype=wp.vec3)
    o2 = wp.mesh_query_point(m, q2, 1e6)
    ms = timed_ms(lambda: wp.mesh_query_point(m, q2, 1e6, out=o2), 3)
    r.update(signed_sample=ns, signed_ms=ms, signed_qps=ns / ms * 1e3)
    res["c2"] = r
    print("c2", r, flush=True)
    del m, q, out, q2, o2

if "c3" in which:
    P, I = mg.heightfield(2237, 4)
    pts, idx = wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32)
    S, D = mg.pinhole_rays(4096, 4096)
    s, d = wp.array(S, dtype=wp.vec3), wp.array(D, dtype=wp.vec3)
    for bits in (30, 63):
        m, r = build_times(pts, idx, morton_bits=bits)
        o = wp.mesh_query_ray(m, s, d, 1e6)
        ms = timed_ms(lambda: wp.mesh_query_ray(m, s, d, 1e6, out=o), 5)
        r.update(triangles=len(I) // 3, rays=len(S), ray_ms=ms, rays_per_s=len(S) / ms * 1e3, hit_fraction=float(o.result.numpy().mean()))
        res[f"c3_morton{bits}"] = r
        print("c3", bits, r, flush=True)
        del m, o
    del s, d

if "c4" in which:
    n = 1415
    P, I = mg.cloth(n, 0)
    pts, idx = wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32)
    m, r = build_times(pts, idx)
    nq = 1 << 23
    frames = []
    out = None
    prev = P
    for f in range(1, 13):
        Pf, _ = mg.cloth(n, f)
        rng = np.random.default_rng(5 + f)
        Q = (prev[rng.integers(0, prev.shape[0], nq)] + rng.normal(0, 0.01, (nq, 3))).astype(np.float32)
        q = wp.array(Q, dtype=wp.vec3)
        pts.assign(Pf)  # in-place vertex update (not timed: the simulation would do this on the device)
        if out is None:
            out = wp.mesh_query_point_no_sign(m, q, 0.05)
        sync()

        def frame():
            m.refit()
            wp.mesh_query_point_no_sign(m, q, 0.05, out=out)

        frames.append(event_ms(core, frame, stream))
        prev = Pf
        del q
    r.update(triangles=len(I) // 3, queries_per_frame=nq, frame_ms_median=statistics.median(frames[2:]), frame_ms_all=frames,
             found_fraction=float(out.result.numpy().mean()))
    res["c4"] = r
    print("c4", r, flush=True)
    del m, out

if "c5" in which:
    P, I = mg.heightfield(7072, 4)
    pts, idx = wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32)
    shard = 125_000_000
    rng = np.random.default_rng(6)
    lo, hi = P.min(0), P.max(0)
    c, h = 0.5 * (lo + hi), 0.6 * (hi - lo)
    Q = (c + (rng.random((shard, 3), dtype=np.float32) * 2 - 1) * h).astype(np.float32)
    for bits, nq in ((63, shard), (30, 1 << 24)):
        m, r = build_times(pts, idx, morton_bits=bits)
        q = wp.array(Q[:nq], dtype=wp.vec3)
        o = wp.mesh_query_point_no_sign(m, q, 1e6)
        ms = timed_ms(lambda: wp.mesh_query_point_no_sign(m, q, 1e6, out=o), 2, warm=0)
        r.update(triangles=len(I) // 3, queries=nq, no_sign_ms=ms, no_sign_qps=nq / ms * 1e3)
        res[f"c5_morton{bits}"] = r
        print("c5", bits, r, flush=True)
        del m, q, o

os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/configs_ours.json", "w"), indent=1)
