"""Generates tests/golden/golden_sign_normal.npz by running the UNMODIFIED reference C++
(oracle/_ref/libwarp_ref_cpu.so) in the dev container: mesh_query_point_sign_normal
(warp/native/mesh.h:860-1090), Mesh.average_edge_length (mesh.cpp:140-155), mesh_query_furthest_point_no_sign
(mesh.h:678-858) and mesh_eval_face_normal (mesh.h:2870-2888) on the mesh of golden_cpu.npz,
over the reference's own SAH tree and over the LBVH trees stored there.  Besides the box queries of golden_cpu.npz
the query set holds points next to vertices, edge midpoints and the vertices themselves (the welding band).

    python tests/golden/make_golden_sign_normal.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "golden_cpu.npz"))
P, I = g["mesh_points"], g["mesh_indices"]
T = I.reshape(-1, 3)
rng = np.random.default_rng(31)
near_vertex = P[rng.integers(0, len(P), 300)] + rng.normal(0, 1e-4, (300, 3))
tri = T[rng.integers(0, len(T), 300)]
edge_mid = 0.5 * (P[tri[:, 0]] + P[tri[:, 1]]) + rng.normal(0, 1e-5, (300, 3))
Q = np.concatenate([g["queries"][:600], near_vertex, edge_mid, P[:200]]).astype(np.float32)
ref_mesh = oracle.RefMesh(P, I)  # the reference's constructor computes average_edge_length
avg = ref_mesh.average_edge_length
out = {"queries": Q, "average_edge_length": np.float32(avg)}


def tree_of(prefix):
    return {k: g[f"{prefix}_tree_{k}"] for k in ("node_lowers", "node_uppers", "primitive_indices")} | {
        "root": int(g[f"{prefix}_tree_root"])}


for name in ("sah", "lbvh1", "lbvh4"):
    m = oracle.RefMesh.from_tree(P, I, tree_of(name))
    m.average_edge_length = avg
    for tag, eps, md in (("e3", 1e-3, 1e6), ("e1", 1e-1, 1e6), ("e3near", 1e-3, 0.05)):
        r = m.query_point_sign_normal(Q, md, eps)
        for k, v in r.items():
            out[f"{name}_{tag}_{k}"] = v

    for tag, md in (("far0", 0.0), ("far2", 2.0)):
        r = m.query_furthest_point_no_sign(Q, md)
        for k, v in r.items():
            out[f"{name}_{tag}_{k}"] = v
out["normal_faces"] = rng.integers(0, len(T), 256).astype(np.int32)
out["face_normals"] = ref_mesh.eval_face_normal(out["normal_faces"])

np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_sign_normal.npz"), **out)
print(avg, {k: (v.shape, v.dtype) for k, v in out.items() if k.startswith("lbvh4")})
