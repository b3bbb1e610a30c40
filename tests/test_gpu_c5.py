"""Config C5 (BASELINE.json configs[4]): the 7072 x 7072 heightfield, 99 998 082 triangles.

The LBVH of the full-size mesh is diffed against the oracle on ALL 2N - 1 nodes (keys, sorted order, parents, leaf
marking incl. the depth >= 32 rule, boxes), the refit on every visible node, and a sample of the config's closest-point
queries bit for bit -- on the 30-bit parity tree; the 63-bit tree (quality option, no reference counterpart) must give
the same closest distance (bit-equal up to a counted handful of last-ulp cases, 1e-5 relative always) and the same face
except at exact distance ties, which are counted.
"""
import numpy as np
import pytest

from helpers import assert_results_equal, assert_tree_equal, visible_mask
from warp_b200 import meshgen as mg

SIDE = 7072


def _closest_dsq(P, I, Q, face, u, v):
    """float32 restatement of the query's own distance (mesh.h:569-573: c = u a + v b + w c; |c - p|^2)."""
    tri = P[I.reshape(-1, 3)[face]]
    u, v = u.astype(np.float32), v.astype(np.float32)
    w = (np.float32(1.0) - u - v).astype(np.float32)
    c = (u[:, None] * tri[:, 0] + v[:, None] * tri[:, 1]).astype(np.float32) + w[:, None] * tri[:, 2]
    d = (c - Q).astype(np.float32)
    return (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]).astype(np.float32)


@pytest.mark.gpu
def test_c5_full_size_tree_refit_and_queries(wp, oracle_mod):
    P, I = mg.heightfield(SIDE, 4)
    T = len(I) // 3
    assert T == 99_998_082
    pts = wp.array(P, dtype=wp.vec3)
    idx = wp.array(I, dtype=wp.int32)
    m = wp.Mesh(pts, idx)  # leaf 4, 30-bit keys: the reference's tree
    want = oracle_mod.mesh_lbvh_build(P, I, 4)
    got = m.download_tree()
    assert got["deep"] == 1 and got["height"] >= 31  # ~100 centroids per 1024^3 cell: the depth rule fires
    assert_tree_equal(got, want)
    del got

    # sampled queries of the config (uniform in the AABB x 1.2), bit for bit on the parity tree
    rng = np.random.default_rng(6)
    lo, hi = P.min(0), P.max(0)
    c, h = 0.5 * (lo + hi), 0.6 * (hi - lo)
    nq = 40_000
    Q = (c + (rng.random((nq, 3), dtype=np.float32) * 2 - 1) * h).astype(np.float32)
    ours30 = wp.mesh_query_point_no_sign(m, wp.array(Q, dtype=wp.vec3), 1e6).numpy()
    ref = oracle_mod.query_point_no_sign(P, I, want, Q, 1e6)
    assert_results_equal(ours30, ref, ("result", "face", "u", "v"))
    assert ours30["result"].all()

    # refit after a smooth deformation: every visible node's box
    P2 = P.copy()
    P2[:, 2] += (0.01 * np.sin(40.0 * P[:, 0].astype(np.float64)) * np.cos(31.0 * P[:, 1].astype(np.float64))).astype(np.float32)
    pts.assign(P2)
    m.refit()
    lo2, hi2 = oracle_mod.triangle_bounds(P2, I)
    oracle_mod.lbvh_refit(want, lo2, hi2)
    got = m.download_tree()
    vis = visible_mask(want)
    for name in ("node_lowers", "node_uppers"):
        assert np.array_equal(got[name]["ib"], want[name]["ib"])
        for f in "xyz":
            assert np.array_equal(got[name][f][vis], want[name][f][vis]), (name, f)
    del got, want, vis
    ours30b = wp.mesh_query_point_no_sign(m, wp.array(Q, dtype=wp.vec3), 1e6).numpy()
    del m

    # 63-bit tree on the deformed mesh: same closest distance everywhere, same face except at exact ties
    m63 = wp.Mesh(pts, idx, morton_bits=63)
    ours63 = wp.mesh_query_point_no_sign(m63, wp.array(Q, dtype=wp.vec3), 1e6).numpy()
    assert ours63["result"].all()
    d30 = _closest_dsq(P2, I, Q, ours30b["face"], ours30b["u"], ours30b["v"])
    d63 = _closest_dsq(P2, I, Q, ours63["face"], ours63["u"], ours63["v"])
    # A box is culled by ITS float distance, which can round a hair above the float distance of a triangle inside it, so
    # two trees may settle on triangles whose distances differ in the last bits (measured: 32 of 40 000): bit-equal
    # for >= 99.5 %, within BASELINE.json's 1e-5 relative everywhere.
    off = d30 != d63
    print(f"C5: closest distance differs in the last bits for {int(off.sum())} of {nq} queries between the two trees")
    assert off.sum() <= nq // 200 and np.allclose(d30, d63, rtol=1e-5, atol=0), f"{off.sum()} queries with a different closest distance"
    # ties are the rule on this mesh, not the exception: the closest point of a far-away query is usually a grid vertex,
    # which six triangles share (measured: ~52 % of the sample) -- where the faces differ the closest POINT must not
    diff = ours30b["face"] != ours63["face"]
    print(f"C5: {int(diff.sum())} of {nq} sampled queries resolve an exact distance tie differently on the 63-bit tree")

    def closest(a):
        tri = P2[I.reshape(-1, 3)[a["face"]]].astype(np.float64)
        u, v = a["u"].astype(np.float64)[:, None], a["v"].astype(np.float64)[:, None]
        return u * tri[:, 0] + v * tri[:, 1] + (1 - u - v) * tri[:, 2]

    # Different POINTS can sit at float-equal distances too: around the foot point the distance is flat (a grid step of
    # 1.4e-4 to the side changes d^2 ~ 0.1 by a couple of ulps), so the two trees may report neighbouring grid points.
    # What must hold is the distance (checked above) and that the points are neighbours, not far apart.
    err = np.abs(closest(ours30b) - closest(ours63)).max(axis=1)
    print(f"C5: closest point identical (< 1e-5) for {100 * float((err < 1e-5).mean()):.2f} % of the sample, max offset {float(err.max()):.2e}")
    assert err.max() < 2e-3, float(err.max())
    same = ~diff & ~off
    assert np.array_equal(ours30b["u"][same], ours63["u"][same]) and np.array_equal(ours30b["v"][same], ours63["v"][same])
