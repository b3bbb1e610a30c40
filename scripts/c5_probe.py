"""C5 instrumentation: why is the 100 M-triangle closest-point rate so far below C2's?  Counts sibling-pair and
packed-triangle fetches per query on the 30-bit parity tree and the 63-bit tree, for queries uniform in the AABB x 1.2
(the config) and for queries close to the surface.  Writes gpurun_out/c5_probe.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import warp_b200 as wp
from warp_b200 import _lib, meshgen as mg
from bench import event_ms

core = _lib.core()
stream = core.wp_cuda_context_get_stream(None)
n_side = int(os.environ.get("C5_SIDE", "7072"))
t0 = time.time()
P, I = mg.heightfield(n_side, 4)
print("gen s", time.time() - t0, flush=True)
pts, idx = wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32)
rng = np.random.default_rng(6)
lo, hi = P.min(0), P.max(0)
c, h = 0.5 * (lo + hi), 0.6 * (hi - lo)
NQ = 1 << 22
Q = (c + (rng.random((NQ, 3), dtype=np.float32) * 2 - 1) * h).astype(np.float32)
near = P[rng.integers(0, len(P), NQ)] + rng.normal(0, 0.002, (NQ, 3)).astype(np.float32)
res = {"side": n_side, "triangles": len(I) // 3}
for bits in (30, 63):
    m = wp.Mesh(pts, idx, morton_bits=bits)
    core.wp_cuda_context_synchronize(None)
    info = m.info() if hasattr(m, "info") else None
    r = {"info": str(info)}
    for name, q in (("box", Q), ("near", near.astype(np.float32))):
        qd = wp.array(q, dtype=wp.vec3)
        out = wp.mesh_query_point_no_sign(m, qd, 1e6)
        core.wp_cuda_context_synchronize(None)
        ms = min(event_ms(core, lambda: wp.mesh_query_point_no_sign(m, qd, 1e6, out=out), stream) for _ in range(2))
        ns = 1 << 20
        qs = wp.array(q[:ns], dtype=wp.vec3)
        with wp.query_stats() as st:
            wp.mesh_query_point_no_sign(m, qs, 1e6)
            core.wp_cuda_context_synchronize(None)
        r[name] = {"queries": NQ, "ms": ms, "qps": NQ / ms * 1e3, "pairs_per_query": st.pair_fetches / ns,
                   "tris_per_query": st.tri_fetches / ns}
        print(bits, name, r[name], flush=True)
        del qd, out, qs
    res[f"morton{bits}"] = r
    del m
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/c5_probe.json", "w"), indent=1)
