"""Rebuild time with / without the small-node kernel (WARP_B200_SMALL_NODES), per mesh and key width."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W = r'''
import os, sys, statistics
sys.path.insert(0, %r)
import warp_b200 as wp
from warp_b200 import _lib, meshgen as mg
from bench import event_ms
core = _lib.core(); stream = core.wp_cuda_context_get_stream(None)
for name, (P, I) in (("c2 sphere", mg.noisy_sphere(8, 0.02, 1)), ("hf 2237", mg.heightfield(2237, 4)), ("cloth 1415", mg.cloth(1415, 0))):
    pts, idx = wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32)
    for bits in (30, 63):
        m = wp.Mesh(pts, idx, morton_bits=bits)
        fn = core.wp_b200_mesh_rebuild_device
        fn(m.id); core.wp_cuda_context_synchronize(None)
        b = statistics.median([event_ms(core, lambda: fn(m.id), stream) for _ in range(9)])
        print(f"  {name:12s} bits {bits}: rebuild {b:.3f} ms", flush=True)
        del m
''' % ROOT
for flag in ("0", "1"):
    print("WARP_B200_SMALL_NODES=" + flag, flush=True)
    r = subprocess.run([sys.executable, "-c", W], env=dict(os.environ, WARP_B200_SMALL_NODES=flag), capture_output=True, text=True)
    print(r.stdout, r.stderr[-300:], flush=True)
