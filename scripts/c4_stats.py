"""C4 traversal statistics: pair / triangle fetches per query on the cloth mesh, for a few query sets."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import warp_b200 as wp
from warp_b200 import _lib, meshgen as mg
from bench import event_ms
core = _lib.core(); stream = core.wp_cuda_context_get_stream(None)
n = 1415
P, I = mg.cloth(n, 0)
m = wp.Mesh(wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32))
info = m.download_tree() if False else None
nq = 1 << 22
rng = np.random.default_rng(5)
for name, Q, md in (("jitter 0.01, max_dist 0.05", (P[rng.integers(0, len(P), nq)] + rng.normal(0, 0.01, (nq, 3))).astype(np.float32), 0.05),
                    ("jitter 0.001, max_dist 0.05", (P[rng.integers(0, len(P), nq)] + rng.normal(0, 0.001, (nq, 3))).astype(np.float32), 0.05),
                    ("jitter 0.01, max_dist 1e6", (P[rng.integers(0, len(P), nq)] + rng.normal(0, 0.01, (nq, 3))).astype(np.float32), 1e6),
                    ("xy jitter only 0.01", (P[rng.integers(0, len(P), nq)] + rng.normal(0, 0.01, (nq, 3)) * np.array([1, 1, 0])).astype(np.float32), 0.05)):
    q = wp.array(Q, dtype=wp.vec3)
    out = wp.mesh_query_point_no_sign(m, q, md)
    ms = min(event_ms(core, lambda: wp.mesh_query_point_no_sign(m, q, md, out=out), stream) for _ in range(3))
    with wp.query_stats() as st:
        wp.mesh_query_point_no_sign(m, q, md, out=out); wp.synchronize()
    print(f"{name:32s} {nq/ms/1e3:7.1f} Mq/s  pairs/q {st.pair_fetches/nq:7.1f}  tris/q {st.tri_fetches/nq:6.1f}  found {out.result.numpy().mean():.3f}")
