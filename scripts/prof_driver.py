"""Small driver for ncu: C2 mesh build x3 (in-place rebuild), refit x3, one query batch, one ray batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import warp_b200 as wp
from warp_b200 import meshgen as mg

nq = int(os.environ.get("PROF_NQ", 1 << 22))
P, I = mg.noisy_sphere(8, 0.02, 1)
pts = wp.array(P, dtype=wp.vec3)
mesh = wp.Mesh(pts, wp.array(I, dtype=wp.int32))
for _ in range(3):
    mesh.rebuild()
for _ in range(3):
    mesh.refit()
Q = wp.array(mg.box_queries(P, nq, seed=2), dtype=wp.vec3)
for _ in range(2):
    r = wp.mesh_query_point_no_sign(mesh, Q, 1e6)
if os.environ.get("PROF_SIGN"):
    r = wp.mesh_query_point(mesh, wp.array(mg.box_queries(P, nq // 8, seed=2), dtype=wp.vec3), 1e6)
S, D = mg.random_rays(P, nq, seed=3)
r2 = wp.mesh_query_ray(mesh, wp.array(S, dtype=wp.vec3), wp.array(D, dtype=wp.vec3), 1e6)
wp.synchronize()
print("done", int(r.result.numpy().sum()))
