"""ncu driver: refit of the C4 cloth mesh (3 998 792 triangles), which takes the wavefront path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import warp_b200 as wp
from warp_b200 import meshgen as mg

P, I = mg.cloth(1415, 0)
pts = wp.array(P, dtype=wp.vec3)
mesh = wp.Mesh(pts, wp.array(I, dtype=wp.int32))
for f in range(1, 4):
    pts.assign(mg.cloth(1415, f)[0])
    mesh.refit()
wp.synchronize()
print("done")
