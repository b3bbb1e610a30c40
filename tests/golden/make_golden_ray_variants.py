"""Generates tests/golden/golden_ray_variants.npz by running the UNMODIFIED reference C++
(oracle/_ref/libwarp_ref_cpu.so) in the dev container: mesh_query_ray_anyhit,
mesh_query_ray_count_intersections, mesh_eval_position and the mesh_query_aabb iterator
(warp/native/mesh.h:1893-2032, 2767-2785, 2476-2712)
on the mesh / rays of golden_cpu.npz, over the reference's own SAH tree and over the LBVH trees stored there.

    python tests/golden/make_golden_ray_variants.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "golden_cpu.npz"))
P, I, S, D = g["mesh_points"], g["mesh_indices"], g["ray_starts"], g["ray_dirs"]
rng = np.random.default_rng(21)
# half the rays start inside the sphere (odd counts), a few are axis aligned (robust slab path)
S2 = S.copy()
S2[::2] = (0.3 * rng.standard_normal((len(S2[::2]), 3))).astype(np.float32)
out = {"starts": S2, "dirs": D}
F = rng.integers(0, len(I) // 3 if I.ndim == 1 else len(I), 512).astype(np.int32)
U = rng.random(512).astype(np.float32)
V = ((1 - U) * rng.random(512)).astype(np.float32)
out["eval_face"], out["eval_u"], out["eval_v"] = F, U, V


QLO = (rng.random((128, 3)) * 2.2 - 1.2).astype(np.float32)
QHI = (QLO + rng.random((128, 3)).astype(np.float32) * 0.6).astype(np.float32)
out["aabb_lowers"], out["aabb_uppers"] = QLO, QHI
TLO, THI = oracle.triangle_bounds(P, I)


def tree_of(prefix):
    return {k: g[f"{prefix}_tree_{k}"] for k in ("node_lowers", "node_uppers", "primitive_indices")} | {
        "root": int(g[f"{prefix}_tree_root"])}


for name in ("sah", "lbvh1", "lbvh4"):
    m = oracle.RefMesh.from_tree(P, I, tree_of(name))
    for mt in (0.5, 1.0, 1e6):
        out[f"{name}_anyhit_{mt:g}"] = m.query_ray_anyhit(S2, D, mt)
    out[f"{name}_count"] = m.query_ray_count(S2, D)
    out[f"{name}_eval_position"] = m.eval(F, U, V)
    out[f"{name}_aabb_offsets"], out[f"{name}_aabb_indices"] = m.query_aabb(QLO, QHI, item_bounds=(TLO, THI))
    # sign by ray parity; NOTE: oracle/_ref is built with g++, which draws the three direction offsets right to left
    for ns in (1, 3):
        r = m.query_point_sign_parity(g["queries"], 1e6, ns, 0.1)
        for k, v in r.items():
            out[f"{name}_parity{ns}_{k}"] = v

np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_ray_variants.npz"), **out)
print({k: (v.shape, v.dtype) for k, v in out.items()})
