"""Closest-point throughput on trees with depth-rule leaves: heightfields whose 30-bit keys collide (10 M / 100 M triangles)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import warp_b200 as wp
from warp_b200 import _lib, meshgen as mg
from bench import event_ms
core = _lib.core(); stream = core.wp_cuda_context_get_stream(None)
for n, nq in [(int(a), 1 << 21) for a in sys.argv[1:]] or [(2237, 1 << 22), (7072, 1 << 21)]:
    P, I = mg.heightfield(n, 4)
    m = wp.Mesh(wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32))
    q = wp.array(mg.box_queries(P, nq, seed=6), dtype=wp.vec3)
    out = wp.mesh_query_point_no_sign(m, q, 1e6)
    ms = min(event_ms(core, lambda: wp.mesh_query_point_no_sign(m, q, 1e6, out=out), stream) for _ in range(3))
    with wp.query_stats() as st:
        wp.mesh_query_point_no_sign(m, q, 1e6, out=out); wp.synchronize()
    print(f"T={len(I)//3:>10d} {nq/ms/1e3:8.2f} Mq/s  pairs/q {st.pair_fetches/nq:7.1f}  tris/q {st.tri_fetches/nq:7.1f}  checksum {int(out.face.numpy().astype(np.int64).sum())}", flush=True)
    del m, q, out
