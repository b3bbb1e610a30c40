"""ncu driver: rebuild x3 + refit x3 of the C3-size heightfield (9 999 392 triangles)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warp_b200 as wp
from warp_b200 import meshgen as mg
P, I = mg.heightfield(int(os.environ.get("PROF_SIDE", "2237")), 4)
m = wp.Mesh(wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32), morton_bits=int(os.environ.get("PROBE_BITS", "30")))
for _ in range(3):
    m.rebuild()
for _ in range(3):
    m.refit()
wp.synchronize()
print("done")
