import sys; sys.path.insert(0, ".")
import numpy as np, statistics
import warp_b200 as wp
from warp_b200 import _lib, meshgen as mg
from bench import event_ms
core=_lib.core(); stream=core.wp_cuda_context_get_stream(None)
P,I = mg.noisy_sphere(8,0.02,1)
m=wp.Mesh(wp.array(P,dtype=wp.vec3), wp.array(I,dtype=wp.int32))
S,D = mg.random_rays(P, 1<<22, seed=3); s=wp.array(S,dtype=wp.vec3); d=wp.array(D,dtype=wp.vec3)
out=wp.mesh_query_ray(m,s,d,1e6)
for mode in (0,1):
    wp.set_ray_order(mode)
    wp.mesh_query_ray(m,s,d,1e6,out=out)
    t=statistics.median([event_ms(core, lambda: wp.mesh_query_ray(m,s,d,1e6,out=out), stream) for _ in range(3)])
    print(f"random rays 4M, ray_order={mode}: {t:.2f} ms  {(1<<22)/t/1e3:.1f} Mrays/s")
# coherent primary rays on the C3 terrain: cost of ordering a batch that is coherent already
P, I = mg.heightfield(2237, 4)
m = wp.Mesh(wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32))
S, D = mg.pinhole_rays(4096, 4096); s = wp.array(S, dtype=wp.vec3); d = wp.array(D, dtype=wp.vec3)
out = wp.mesh_query_ray(m, s, d, 1e6)
for mode in (0, 1):
    wp.set_ray_order(mode)
    wp.mesh_query_ray(m, s, d, 1e6, out=out)
    t = statistics.median([event_ms(core, lambda: wp.mesh_query_ray(m, s, d, 1e6, out=out), stream) for _ in range(3)])
    print(f"C3 primary rays 16.8M, ray_order={mode}: {t:.2f} ms  {(1<<24)/t/1e6:.2f} Grays/s")
wp.set_ray_order(0)
