"""Generates tests/golden/golden_bvh_kinds.npz by running the UNMODIFIED reference C++ (oracle/_ref/libwarp_ref_cpu.so)
in the dev container: the sphere and capsule kinds of the generic wp.Bvh iterator -- bvh_query_sphere /
bvh_query_sphere_next and bvh_query_capsule / bvh_query_capsule_next (warp/native/bvh.h:462-492, 529-551, 560-664,
node tests intersect.h:158-181, 197-205) -- over LBVH trees (leaf size 1 and 4) of 2000 random boxes, and
mesh_query_sphere / mesh_query_sphere_next (mesh.h:2457-2737) on the mesh of golden_cpu.npz plus three zero-area faces.  The LBVH trees
come from the restatement, which is itself pinned on the reference's CUDA LBVH (golden_ref_lbvh.npz).

    python tests/golden/make_golden_bvh_kinds.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from helpers import random_boxes  # noqa: E402

lo, hi = random_boxes(2000, seed=41)
rng = np.random.default_rng(42)
n = 300
C = (rng.random((n, 3)) * 10).astype(np.float32)
R = (rng.random(n) * 1.5 - 0.1).astype(np.float32)  # a few negative radii: clamped to 0 (bvh.h:537, 548)
D = rng.standard_normal((n, 3)).astype(np.float32)
D[::7, 0] = 0  # directions with zero components: the robust slab path
D[::11, 1] = 0
D[::13] = (0, 0, 1)
D /= np.linalg.norm(D, axis=1, keepdims=True)
out = {"seed_boxes": np.int32(41), "centers": C, "radii": R, "dirs": D}
for leaf in (1, 4):
    tree = oracle.lbvh_build(lo, hi, leaf)
    out[f"leaf{leaf}_sphere_offsets"], out[f"leaf{leaf}_sphere_indices"] = oracle.ref_bvh_query_kind(tree, lo, hi, "sphere", C, radii=R)
    for tag, md in (("inf", 3.4028234663852886e38), ("3", 3.0)):
        out[f"leaf{leaf}_capsule{tag}_offsets"], out[f"leaf{leaf}_capsule{tag}_indices"] = oracle.ref_bvh_query_kind(
            tree, lo, hi, "capsule", C, D, radii=R, max_dist=md)
# mesh_query_sphere on the golden mesh + degenerate faces (repeated vertex, collinear, single point)
gc = np.load(os.path.join(ROOT, "tests", "golden", "golden_cpu.npz"))
P, I = gc["mesh_points"], gc["mesh_indices"].reshape(-1)
I2 = np.concatenate([I, np.array([0, 0, 5, 3, 7, 7, 10, 10, 10], np.int32)]).astype(np.int32)
MC = np.concatenate([gc["queries"][:400], P[:100]]).astype(np.float32)
MR = (rng.random(len(MC)) * 0.5 - 0.02).astype(np.float32)
out["mesh_indices"], out["mesh_centers"], out["mesh_radii"] = I2, MC, MR
tlo, thi = oracle.triangle_bounds(P, I2)
for leaf in (1, 4):
    tree = oracle.mesh_lbvh_build(P, I2, leaf)
    m = oracle.RefMesh.from_tree(P, I2, tree)
    out[f"mesh_leaf{leaf}_sphere_offsets"], out[f"mesh_leaf{leaf}_sphere_indices"] = m.query_sphere(MC, MR, item_bounds=(tlo, thi))

np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_bvh_kinds.npz"), **out)
print({k: v.shape for k, v in out.items()})
