"""ncu driver: C3 terrain (10 M-triangle heightfield) and the 4096 x 4096 primary rays, mesh_query_ray x 3."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import warp_b200 as wp
from warp_b200 import meshgen as mg

P, I = mg.heightfield(int(os.environ.get("PROF_HF", 2237)), 4)
mesh = wp.Mesh(wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32), morton_bits=int(os.environ.get("PROBE_BITS", "30")))
w = int(os.environ.get("PROF_W", 4096))
S, D = mg.pinhole_rays(w, w)
s, d = wp.array(S, dtype=wp.vec3), wp.array(D, dtype=wp.vec3)
for _ in range(3):
    r = wp.mesh_query_ray(mesh, s, d, 1e6)
wp.synchronize()
print("hits", float(r.result.numpy().mean()))
