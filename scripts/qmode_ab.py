"""A/B of the memory-side variants of the unsigned closest-point kernel (query.cu QM_* bits) and of the L2 persistence
window.  One subprocess per (mode, persist) because both are read once per process.

    python scripts/qmode_ab.py [c5]      -> gpurun_out/qmode_ab.json
"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json, statistics, hashlib
sys.path.insert(0, %r)
import numpy as np
import warp_b200 as wp
from warp_b200 import _lib, meshgen as mg, workload
from bench import event_ms, kernel_timing
core = _lib.core(); stream = core.wp_cuda_context_get_stream(None)
which = sys.argv[1]
res = {}
def run(mesh, q, max_dist, reps=5):
    out = wp.mesh_query_point_no_sign(mesh, q, max_dist)
    core.wp_cuda_context_synchronize(None)
    core.wp_b200_kernel_timing_enable(1); kernel_timing(core)
    ms = [event_ms(core, lambda: wp.mesh_query_point_no_sign(mesh, q, max_dist, out=out), stream) for _ in range(reps)]
    kms, kl = kernel_timing(core); core.wp_b200_kernel_timing_enable(0)
    h = hashlib.sha1()
    for k in ("result", "face", "u", "v"):
        h.update(getattr(out, k).numpy().tobytes())
    return {"ms": statistics.median(ms), "kernel_ms": kms / max(kl, 1), "sha1": h.hexdigest()[:16]}
if which == "c2c4":
    P, I = mg.noisy_sphere(8, 0.02, 1)
    m = wp.Mesh(wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32))
    q = wp.array(mg.box_queries(P, 1 << 24, seed=2), dtype=wp.vec3)
    res["c2"] = run(m, q, 1e6)
    del m, q
    _, Ic = mg.cloth(1415, 0)
    cf = workload.ClothFrames(1415)
    cm = wp.Mesh(cf.points, wp.array(Ic, dtype=wp.int32))
    cf.advance(40); cf.update_points(); cm.refit()
    qc = wp.empty(1 << 23, wp.vec3); cf.queries(qc, 0.01)
    res["c4_frame40"] = run(cm, qc, 0.05, 3)
else:
    P, I = mg.heightfield(7072, 4)
    lo, hi = P.min(0).astype(np.float64), P.max(0).astype(np.float64)
    c, h = 0.5 * (lo + hi), 0.6 * (hi - lo)
    pts, idx = wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32)
    q = wp.empty(1 << 22, wp.vec3); workload.box_queries(q, 0, 6, c - h, c + h)
    for bits in (30, 63):
        m = wp.Mesh(pts, idx, morton_bits=bits)
        res["c5_%%d" %% bits] = run(m, q, 1e6, 2)
        del m
print("AB " + json.dumps(res))
''' % ROOT

def one(mode, persist, which):
    env = dict(os.environ, WARP_B200_QMODE=str(mode), WARP_B200_L2_PERSIST=str(persist))
    r = subprocess.run([sys.executable, "-c", WORKER, which], capture_output=True, text=True, env=env, timeout=900)
    for ln in r.stdout.splitlines():
        if ln.startswith("AB "):
            return json.loads(ln[3:])
    return {"error": r.stderr[-500:]}

which = "c5" if len(sys.argv) > 1 and sys.argv[1] == "c5" else "c2c4"
modes = [int(x) for x in os.environ.get("AB_MODES", "0,1,2,3,4,5,6,7,8,9,15").split(",")]
out = {}
for persist in (0, 1):
    for mode in modes:
        r = one(mode, persist, which)
        out[f"mode{mode}_persist{persist}"] = r
        print(mode, persist, json.dumps(r), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"qmode_ab_{which}.json"), "w"), indent=1)
