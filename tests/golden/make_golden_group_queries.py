"""Generates tests/golden/golden_group_queries.npz by running the UNMODIFIED reference header code
(oracle/_ref/libwarp_ref_cpu.so): bvh_get_group_root (warp/native/bvh.h:376-390) and the generic iterator with a
``root`` argument (bvh.h:494-664) on grouped LBVH trees (tree arrays from the oracle builder, which is itself pinned
bit for bit on the reference's CUDA LBVH, tests/golden/golden_ref_lbvh.npz).

    python tests/golden/make_golden_group_queries.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle  # noqa: E402
from test_oracle import _group_case, _grouped_mesh_case  # noqa: E402

out = {}
for leaf in (1, 4):
    lo, hi, groups, qlo, qhi, s, d, gid = _group_case(900 + leaf, 400, 6)
    tree = oracle.lbvh_build(lo, hi, leaf, groups=groups)
    roots = oracle.ref_bvh_group_roots(tree, groups, gid)
    out[f"leaf{leaf}_roots"] = roots
    out[f"leaf{leaf}_aabb_offsets"], out[f"leaf{leaf}_aabb_indices"] = oracle.ref_bvh_query(tree, lo, hi, qlo, qhi, roots=roots)
    out[f"leaf{leaf}_ray_offsets"], out[f"leaf{leaf}_ray_indices"] = oracle.ref_bvh_query(
        tree, lo, hi, s, d, ray=True, max_dist=6.0, roots=roots)
# grouped MESH: rays restricted to a group's subtree (mesh_query_ray / _anyhit / _count_intersections with `root`)
P, I, T, groups, S, D, gid = _grouped_mesh_case()
tree = oracle.mesh_lbvh_build(P, I, 4, groups=groups)
roots = oracle.ref_bvh_group_roots(tree, groups, gid)
rm = oracle.RefMesh.from_tree(P, I, tree)
out["mesh_roots"] = roots
for k, v in rm.query_ray(S, D, 1e6, roots=roots).items():
    out[f"mesh_ray_{k}"] = v
out["mesh_anyhit"] = rm.query_ray_anyhit(S, D, 0.9, roots=roots)
out["mesh_count"] = rm.query_ray_count(S, D, roots=roots)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_group_queries.npz"), **out)
print({k: v.shape for k, v in out.items()})
