"""e2e timing of the host-buffer entry point for a given WARP_B200_HOST_CHUNK (set in the environment)."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import warp_b200 as wp
from warp_b200 import _lib, meshgen as mg
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import Pinned
core = _lib.core()
P, I = mg.noisy_sphere(8, 0.02, 1)
mesh = wp.Mesh(wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32))
nq = 1 << 24
Q = mg.box_queries(P, nq, seed=2)
hp = Pinned(core, (nq, 3), np.float32); hp.array[:] = Q
out = wp.MeshQueryPoint(*(Pinned(core, (nq,), dt).array for dt in (np.uint8, np.float32, np.int32, np.float32, np.float32)))
# raw copy bandwidth
d = wp.empty(nq * 3, wp.float32)
for name, fn in (("h2d", lambda: core.wp_memcpy_h2d(None, d.ptr, hp.array.ctypes.data, 12 * nq, None)),
                 ("d2h", lambda: core.wp_memcpy_d2h(None, hp.array.ctypes.data, d.ptr, 12 * nq, None))):
    fn(); core.wp_cuda_context_synchronize(None)
    t0 = time.perf_counter(); fn(); core.wp_cuda_context_synchronize(None); dt = time.perf_counter() - t0
    print(f"{name}: {12*nq/dt/1e9:.1f} GB/s")
wp.mesh_query_point_no_sign(mesh, hp.array, 1e6, out=out)
ts = []
for _ in range(4):
    t0 = time.perf_counter(); wp.mesh_query_point_no_sign(mesh, hp.array, 1e6, out=out); ts.append(time.perf_counter() - t0)
print(f"chunk {os.environ.get('WARP_B200_HOST_CHUNK')}: e2e {min(ts)*1e3:.2f} ms  {nq/min(ts)/1e6:.1f} Mq/s")
