"""Generates tests/golden/golden_cpu.npz by running the UNMODIFIED reference C++
(oracle/_ref/libwarp_ref_cpu.so, built from /root/reference/warp/native by oracle/Makefile) in the
dev container.  The fixtures pin oracle/lbvh_oracle.c on machines where the reference tree is absent.

    python tests/golden/make_golden_cpu.py

Contents (all produced by reference code, never by our restatement, except the LBVH *tree* inputs
which the reference cannot build on the host -- bvh.cpp:226-233 -- and which are stored as inputs):
  cube_*      unit-cube point / ray goldens of warp/tests/geometry/test_mesh.py:111-187 on SAH trees
  sah_*       reference SAH tree (leaf 4) of a 320-triangle noisy icosphere + 512 point / ray results
  lbvh_*      LBVH tree (oracle-built, leaf 1 and 4) of the same mesh + the reference traversal's
              answers on it for the same 512 points / rays (incl. sign)
  tri_*       2000 closest_point_to_triangle evaluations; morton_*  2000 morton3<1024> evaluations
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from warp_b200 import meshgen as mg  # noqa: E402

out = {}
d = np.array([-1.2, 2.3, -3.4], np.float32)
d /= np.linalg.norm(d)
for name, idx in (("rh", mg.CUBE_INDICES_RH), ("lh", mg.CUBE_INDICES_LH)):
    for ctor, cname in ((oracle.SAH, "sah"), (oracle.MEDIAN, "median")):
        for leaf in (1, 2, 4):
            m = oracle.RefMesh(mg.CUBE_POINTS, idx, ctor, leaf)
            p = m.query_point([[0.1, 0.2, 0.3]], 1e6)
            r = m.query_ray([[0.1, 0.2, 0.3]], [d], 1e6)
            for k, v in p.items():
                out[f"cube_{name}_{cname}_{leaf}_point_{k}"] = v
            for k, v in r.items():
                out[f"cube_{name}_{cname}_{leaf}_ray_{k}"] = v

P, I = mg.noisy_sphere(2, noise=0.05, seed=11)
Q = mg.box_queries(P, 512, seed=12)
S, D = mg.random_rays(P, 512, seed=13)
# a few axis-aligned rays / rays with zero components exercise the robust slab test
D[:8] = np.eye(3, dtype=np.float32)[np.arange(8) % 3] * np.array([1, -1], np.float32)[np.arange(8) % 2][:, None]
S[:8] = -1.5 * D[:8] + 0.05
out["mesh_points"], out["mesh_indices"], out["queries"], out["ray_starts"], out["ray_dirs"] = P, I, Q, S, D

m = oracle.RefMesh(P, I, oracle.SAH, 4)
t = m.tree()
for k in ("node_lowers", "node_uppers", "primitive_indices"):
    out[f"sah_tree_{k}"] = t[k]
out["sah_tree_root"] = np.int32(t["root"])
for nm, res in (("point", m.query_point(Q, 1e6)), ("point05", m.query_point(Q, 0.5)), ("ray", m.query_ray(S, D, 1e6))):
    for k, v in res.items():
        out[f"sah_{nm}_{k}"] = v

for leaf in (1, 4):
    tree = oracle.mesh_lbvh_build(P, I, leaf)
    for k in ("keys", "node_lowers", "node_uppers", "primitive_indices", "parents"):
        out[f"lbvh{leaf}_tree_{k}"] = tree[k]
    out[f"lbvh{leaf}_tree_root"] = np.int32(tree["root"])
    rm = oracle.RefMesh.from_tree(P, I, tree)
    for nm, res in (("point", rm.query_point(Q, 1e6)), ("point05", rm.query_point(Q, 0.5)), ("ray", rm.query_ray(S, D, 1e6))):
        for k, v in res.items():
            out[f"lbvh{leaf}_{nm}_{k}"] = v

rng = np.random.default_rng(21)
tri = rng.standard_normal((2000, 3, 3)).astype(np.float32)
pts = (rng.standard_normal((2000, 3)) * 2).astype(np.float32)
out["tri_abc"], out["tri_p"] = tri, pts
out["tri_uv"] = np.stack([oracle.ref_closest_point_to_triangle(t[0], t[1], t[2], p) for t, p in zip(tri, pts)])
xyz = rng.random((2000, 3)).astype(np.float32) * 1.2 - 0.1
out["morton_xyz"] = xyz
out["morton_code"] = np.array([oracle.ref_morton3(*map(float, v)) for v in xyz], np.uint32)

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_cpu.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes,", len(out), "arrays")
