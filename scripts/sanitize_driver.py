"""Small driver for compute-sanitizer: touches every round-2 kernel once on small inputs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import warp_b200 as wp
from warp_b200 import _lib, meshgen as mg, workload

core = _lib.core()
P, I = mg.noisy_sphere(4, 0.02, 1)
pts, idx = wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32)
Q = mg.box_queries(P, 40000, seed=2)
for exp in (0, 1):
    core.wp_b200_set_experiment(b"small_nodes", exp)
    for bits in (30, 63):
        m = wp.Mesh(pts, idx, morton_bits=bits)
        r = wp.mesh_query_point_no_sign(m, wp.array(Q, dtype=wp.vec3), 1e6)  # ordered batch: packed records + unpack, 8-byte stack
        r2 = wp.mesh_query_point(m, wp.array(Q, dtype=wp.vec3), 1e6)
        m.refit(); m.rebuild()
core.wp_b200_set_experiment(b"small_nodes", -1)
Ph, Ih = mg.heightfield(200, 4)   # equal-key pairs, depth pass on a taller tree
mh = wp.Mesh(wp.array(Ph, dtype=wp.vec3), wp.array(Ih, dtype=wp.int32))
S, D = mg.random_rays(Ph, 20000, seed=3)
rr = wp.mesh_query_ray(mh, wp.array(S, dtype=wp.vec3), wp.array(D, dtype=wp.vec3), 1e6)
dup = np.tile(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), (700, 1))  # 700 coincident triangles: depth-rule leaves
md = wp.Mesh(wp.array(dup, dtype=wp.vec3), wp.array(np.arange(2100, dtype=np.int32), dtype=wp.int32))
rd = wp.mesh_query_point_no_sign(md, wp.array(mg.box_queries(dup, 40000, seed=4), dtype=wp.vec3), 1e6)
for ctor in ("sah", "median"):
    ms = wp.Mesh(pts, idx, bvh_constructor=ctor)
    wp.mesh_query_point_no_sign(ms, wp.array(Q, dtype=wp.vec3), 1e6)
    ms.refit()
core.wp_b200_bvh_sync_reference_layout(m.id)
cf = workload.ClothFrames(64)
qc = wp.empty(5000, wp.vec3)
cf.advance(3); cf.update_points(); cf.queries(qc, 0.01)
workload.box_queries(qc, 12345, 6, (0, 0, 0), (1, 1, 1))
out = wp.mesh_eval_face_normal(m, r.face) if hasattr(wp, "mesh_eval_face_normal") else None
wp.synchronize()
print("sanitize driver done", int(r.result.numpy().sum()), int(rr.result.numpy().sum()), int(rd.result.numpy().sum()))
