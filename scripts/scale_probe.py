"""Build / refit time vs mesh size (heightfields), with the algorithmic-byte rooflines (396 / 189 B per triangle)."""
import os, sys, time, statistics, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import warp_b200 as wp
from warp_b200 import _lib, meshgen as mg
from bench import event_ms, measured_peaks
core = _lib.core()
peak, _ = measured_peaks()
stream = core.wp_cuda_context_get_stream(None)
for n in [int(a) for a in sys.argv[1:]] or [810, 2237, 4473, 7072]:
    P, I = mg.heightfield(n, 4)
    T = len(I) // 3
    pts = wp.array(P, dtype=wp.vec3); idx = wp.array(I, dtype=wp.int32)
    core.wp_cuda_context_synchronize(None)
    t0 = time.perf_counter(); m = wp.Mesh(pts, idx, morton_bits=int(os.environ.get('PROBE_BITS', '30'))); core.wp_cuda_context_synchronize(None); ctor = 1e3 * (time.perf_counter() - t0)
    fn = core.wp_b200_mesh_rebuild_device
    b = statistics.median([event_ms(core, lambda: fn(m.id), stream) for _ in range(5)])
    r = statistics.median([event_ms(core, m.refit, stream) for _ in range(5)])
    info = m.download_tree() if T < 3_000_000 else None
    print(f"T={T:>10d} ctor {ctor:8.2f} ms  rebuild {b:8.3f} ms ({396*T/b/1e6/peak:5.3f} of HBM)  refit {r:8.3f} ms ({189*T/r/1e6/peak:5.3f} of HBM)", flush=True)
    del m, pts, idx
