"""GPU parity tests of the batched queries: CUDA (through the C ABI) vs the oracle on the same
tree.  Every field is compared for exact equality (the library is built with -fmad=false and
visits nodes in the reference's order, so even exact distance ties resolve identically); the
tolerance form of the bar (1e-5 relative on u, v, t; ties counted) is asserted as well so the test
documents the contract BASELINE.json states."""
import os

import numpy as np
import pytest

from helpers import assert_results_equal
from warp_b200 import meshgen as mg

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "golden_cpu.npz")
POINT_FIELDS = ("result", "sign", "face", "u", "v")
RAY_FIELDS = ("result", "sign", "face", "t", "u", "v", "normal")


def gpu_mesh(wp, P, I, leaf=4):
    return wp.Mesh(wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32), bvh_constructor="lbvh", bvh_leaf_size=leaf)


def test_cube_golden_vectors(wp):
    """warp/tests/geometry/test_mesh.py:111-187 through the CUDA path, leaf sizes 1, 2, 4."""
    d = np.array([-1.2, 2.3, -3.4], np.float32)
    d /= np.linalg.norm(d)
    for idx, sgn in ((mg.CUBE_INDICES_RH, -1.0), (mg.CUBE_INDICES_LH, 1.0)):
        for leaf in (1, 2, 4):
            m = gpu_mesh(wp, mg.CUBE_POINTS, idx, leaf)
            p = wp.mesh_query_point(m, np.array([[0.1, 0.2, 0.3]], np.float32), 1e6)
            assert p.result[0] == 1 and p.face[0] == 1 and p.sign[0] == sgn
            tri = mg.CUBE_POINTS[idx.reshape(-1, 3)[1]]
            pos = p.u[0] * tri[0] + p.v[0] * tri[1] + (1 - p.u[0] - p.v[0]) * tri[2]
            assert np.allclose(pos, (0.1, 0.2, 0.5), atol=1e-6)
            r = wp.mesh_query_ray(m, np.array([[0.1, 0.2, 0.3]], np.float32), d[None], 1e6)
            assert r.result[0] == 1 and r.face[0] == 4 and abs(r.t[0] - 0.557828) < 1e-6
            assert np.sign(r.sign[0]) == sgn


def test_reference_fixture_answers(wp):
    """Answers recorded from the reference C++ on an LBVH tree (tests/golden/golden_cpu.npz)."""
    g = np.load(GOLD)
    P, I, Q, S, D = (g[k] for k in ("mesh_points", "mesh_indices", "queries", "ray_starts", "ray_dirs"))
    for leaf in (1, 4):
        m = gpu_mesh(wp, P, I, leaf)
        assert_results_equal(wp.mesh_query_point(m, Q, 1e6).numpy(), {f: g[f"lbvh{leaf}_point_{f}"] for f in POINT_FIELDS}, POINT_FIELDS)
        assert_results_equal(wp.mesh_query_point(m, Q, 0.5).numpy(), {f: g[f"lbvh{leaf}_point05_{f}"] for f in POINT_FIELDS}, POINT_FIELDS)
        assert_results_equal(wp.mesh_query_ray(m, S, D, 1e6).numpy(), {f: g[f"lbvh{leaf}_ray_{f}"] for f in RAY_FIELDS}, RAY_FIELDS)


@pytest.mark.parametrize("leaf", [1, 4, 8])
def test_point_queries_bit_exact(wp, oracle_mod, leaf):
    P, I = mg.noisy_sphere(5, 0.05, 1)
    m = gpu_mesh(wp, P, I, leaf)
    tree = oracle_mod.mesh_lbvh_build(P, I, leaf)
    Q = mg.box_queries(P, 20000, seed=2)
    for max_dist in (1e6, 0.1):
        want = oracle_mod.query_point(P, I, tree, Q, max_dist)
        got = wp.mesh_query_point(m, Q, max_dist).numpy()  # host buffers -> *_host entry point
        assert_results_equal(got, want, POINT_FIELDS)
        got_ns = wp.mesh_query_point_no_sign(m, wp.array(Q, dtype=wp.vec3), max_dist).numpy()  # device buffers
        assert_results_equal(got_ns, dict(want, sign=np.zeros_like(want["sign"])), POINT_FIELDS)
    # the contract as BASELINE.json words it: faces equal except exact ties (counted), u/v within 1e-5
    want = oracle_mod.query_point(P, I, tree, Q, 1e6)
    got = wp.mesh_query_point(m, Q, 1e6).numpy()
    ties = int((got["face"] != want["face"]).sum())
    assert ties == 0
    assert np.allclose(got["u"], want["u"], rtol=1e-5, atol=1e-7) and np.allclose(got["v"], want["v"], rtol=1e-5, atol=1e-7)


def test_point_queries_on_surface_points_have_many_ties(wp, oracle_mod):
    """Queries AT mesh vertices: distance 0 to every incident face -> exact ties; same winner as the reference order."""
    P, I = mg.icosphere(4)
    m = gpu_mesh(wp, P, I, 4)
    tree = oracle_mod.mesh_lbvh_build(P, I, 4)
    Q = P[:2000].copy()
    assert_results_equal(wp.mesh_query_point_no_sign(m, Q, 1e6).numpy(), oracle_mod.query_point_no_sign(P, I, tree, Q, 1e6), ("result", "face", "u", "v"))


@pytest.mark.parametrize("leaf", [1, 4])
def test_ray_queries_bit_exact(wp, oracle_mod, leaf):
    P, I = mg.heightfield(129)
    m = gpu_mesh(wp, P, I, leaf)
    tree = oracle_mod.mesh_lbvh_build(P, I, leaf)
    S, D = mg.pinhole_rays(160, 120)
    want = oracle_mod.query_ray(P, I, tree, S, D, 1e6)
    assert want["result"].mean() > 0.3
    assert_results_equal(wp.mesh_query_ray(m, S, D, 1e6).numpy(), want, RAY_FIELDS)
    got = wp.mesh_query_ray(m, wp.array(S, dtype=wp.vec3), wp.array(D, dtype=wp.vec3), 1.0).numpy()
    assert_results_equal(got, oracle_mod.query_ray(P, I, tree, S, D, 1.0), RAY_FIELDS)
    # grid-aligned vertical rays hit edges / vertices exactly: exercises the fp64 edge fallback (intersect.h:394-404)
    xs = np.linspace(0, 1, 129, dtype=np.float32)
    gx, gy = np.meshgrid(xs[::4], xs[::4], indexing="ij")
    S2 = np.stack([gx.ravel(), gy.ravel(), np.full(gx.size, 2.0, np.float32)], 1).astype(np.float32)
    D2 = np.tile(np.array([[0, 0, -1]], np.float32), (S2.shape[0], 1))
    want2 = oracle_mod.query_ray(P, I, tree, S2, D2, 1e6)
    assert want2["result"].all()
    assert_results_equal(wp.mesh_query_ray(m, S2, D2, 1e6).numpy(), want2, RAY_FIELDS)


def test_ray_on_shared_edge_never_leaks(wp):
    """warp/tests/geometry/test_mesh_query_ray.py:678-722: 900 rays at a 2-triangle quad all hit."""
    P = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32) - np.array([0.5, 0.5, 0], np.float32)
    I = np.array([0, 1, 2, 0, 2, 3], np.int32)
    g = np.arange(0.1, 0.4, 0.01, dtype=np.float32)
    gx, gy = np.meshgrid(g, g, indexing="ij")
    S = np.stack([gx.ravel(), gy.ravel(), np.ones(gx.size, np.float32)], 1).astype(np.float32)
    D = np.tile(np.array([[0, 0, -1]], np.float32), (S.shape[0], 1))
    S[:, :2] = S[:, [0, 0]]  # along the diagonal x == y: exactly on the shared edge
    for leaf in (1, 2, 4):
        r = wp.mesh_query_ray(gpu_mesh(wp, P, I, leaf), S, D, 1e6)
        assert r.result.all(), f"leaf {leaf}: {int((r.result == 0).sum())} rays leaked"


def test_ray_parallel_to_slab_boundaries(wp, oracle_mod):
    """warp/tests/geometry/test_mesh_query_ray.py:725-842: axis-aligned rays whose origin lies on AABB faces."""
    P, I = mg.CUBE_POINTS, mg.CUBE_INDICES_RH
    S = np.array([[0.5, 0.0, 2.0], [-0.5, 0.25, 2.0], [0.0, 0.5, -2.0], [2.0, 0.5, 0.0], [0.0, -2.0, 0.5]], np.float32)
    D = np.array([[0, 0, -1], [0, 0, -1], [0, 0, 1], [-1, 0, 0], [0, 1, 0]], np.float32)
    for leaf in (1, 2, 4):
        tree = oracle_mod.mesh_lbvh_build(P, I, leaf)
        want = oracle_mod.query_ray(P, I, tree, S, D, 1e6)
        got = wp.mesh_query_ray(gpu_mesh(wp, P, I, leaf), S, D, 1e6).numpy()
        assert got["result"].all()
        assert_results_equal(got, want, RAY_FIELDS)


def test_queries_after_refit(wp, oracle_mod):
    """warp/tests/geometry/test_mesh.py:319-357: translate +10 x, refit, hit new / miss old; then full parity."""
    P, I = mg.noisy_sphere(4, 0.02, 9)
    pts = wp.array(P, dtype=wp.vec3)
    m = wp.Mesh(pts, wp.array(I, dtype=wp.int32))
    P2 = (P + np.array([10, 0, 0], np.float32)).astype(np.float32)
    pts.assign(P2)
    m.refit()
    d = np.array([[0.0, 0.0, -1.0]], np.float32)
    assert wp.mesh_query_ray(m, np.array([[0.0, 0.0, 5.0]], np.float32), d, 1e6).result[0] == 0
    assert wp.mesh_query_ray(m, np.array([[10.0, 0.0, 5.0]], np.float32), d, 1e6).result[0] == 1
    tree = oracle_mod.mesh_lbvh_build(P, I, 4)
    lo2, hi2 = oracle_mod.triangle_bounds(P2, I)
    oracle_mod.lbvh_refit(tree, lo2, hi2)
    Q = mg.box_queries(P2, 5000, seed=4)
    assert_results_equal(wp.mesh_query_point(m, Q, 1e6).numpy(), oracle_mod.query_point(P2, I, tree, Q, 1e6), POINT_FIELDS)
    S, D = mg.random_rays(P2, 5000, seed=5)
    assert_results_equal(wp.mesh_query_ray(m, S, D, 1e6).numpy(), oracle_mod.query_ray(P2, I, tree, S, D, 1e6), RAY_FIELDS)


def test_empty_and_ragged_batches(wp, oracle_mod):
    P, I = mg.noisy_sphere(2)
    m = gpu_mesh(wp, P, I)
    r = wp.mesh_query_point_no_sign(m, np.zeros((0, 3), np.float32), 1.0)
    assert r.result.shape == (0,)
    tree = oracle_mod.mesh_lbvh_build(P, I, 4)
    for n in (1, 31, 33, 129, 1000):
        Q = mg.box_queries(P, n, seed=n)
        assert_results_equal(wp.mesh_query_point(m, Q, 1e6).numpy(), oracle_mod.query_point(P, I, tree, Q, 1e6), POINT_FIELDS)
    # max_dist so small nothing is found: all fields keep their zero defaults (mesh.h:1514-1521)
    far = np.full((5, 3), 100.0, np.float32)
    r = wp.mesh_query_point(m, far, 1.0).numpy()
    assert not r["result"].any() and not r["face"].any() and not r["u"].any() and not r["sign"].any()
    rr = wp.mesh_query_ray(m, far, np.tile(np.array([[1, 0, 0]], np.float32), (5, 1)), 1e6).numpy()
    assert not rr["result"].any() and not rr["normal"].any() and not rr["t"].any()


def test_sliver_triangles_are_skipped(wp, oracle_mod):
    """mesh.h:563: faces with |n| / sum(|e|^2) < 1e-6 never win a closest-point query."""
    P = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [0.5, 1e-8, 1]], np.float32)
    I = np.array([0, 1, 2, 3, 4, 5], np.int32)
    m = gpu_mesh(wp, P, I, 1)
    tree = oracle_mod.mesh_lbvh_build(P, I, 1)
    Q = np.array([[0.5, 0, 1.0], [0.2, 0.2, 0.9], [0.5, 0.0, 0.99]], np.float32)  # right on / next to the sliver
    want = oracle_mod.query_point_no_sign(P, I, tree, Q, 1e6)
    got = wp.mesh_query_point_no_sign(m, Q, 1e6).numpy()
    assert (got["face"] == 0).all() and got["result"].all()
    assert_results_equal(got, want, ("result", "face", "u", "v"))
    r = wp.mesh_query_ray(m, np.array([[0.5, 0.0, 2.0]], np.float32), np.array([[0, 0, -1]], np.float32), 1e6)
    assert r.result[0] == 1  # rays do not skip slivers (mesh.h:1805-1829)


def test_large_batch_properties_c2_shape(wp, oracle_mod):
    """Config C2 shapes (1.3 M triangles, 4 M of the 16 M queries): sampled oracle diff + size-independent
    properties: found everywhere (max_dist 1e6), barycentrics in range, reported point is no farther than
    any of 64 random faces, host and device entry points agree exactly."""
    P, I = mg.noisy_sphere(8, 0.02, 1)
    m = gpu_mesh(wp, P, I, 4)
    Q = mg.box_queries(P, 1 << 22, seed=2)
    got = wp.mesh_query_point_no_sign(m, Q, 1e6).numpy()
    assert got["result"].all()
    u, v = got["u"], got["v"]
    assert (u >= 0).all() and (v >= 0).all() and (u + v <= 1 + 1e-6).all()
    tri = P[I.reshape(-1, 3)[got["face"]]]
    c = u[:, None] * tri[:, 0] + v[:, None] * tri[:, 1] + (1 - u - v)[:, None] * tri[:, 2]
    d = np.linalg.norm(c - Q, axis=1)
    rng = np.random.default_rng(0)
    others = P[I.reshape(-1, 3)[rng.integers(0, len(I) // 3, 64)]].mean(axis=1)  # centroids of 64 random faces
    assert (d[:, None] <= np.linalg.norm(Q[:, None, :] - others[None], axis=2) + 1e-5).all()
    dev = wp.mesh_query_point_no_sign(m, wp.array(Q, dtype=wp.vec3), 1e6).numpy()
    assert_results_equal(dev, got, ("result", "face", "u", "v"))
    tree = oracle_mod.mesh_lbvh_build(P, I, 4)
    sel = rng.integers(0, Q.shape[0], 20000)
    want = oracle_mod.query_point_no_sign(P, I, tree, Q[sel], 1e6)
    assert_results_equal({k: got[k][sel] for k in ("result", "face", "u", "v")}, want, ("result", "face", "u", "v"))


def test_query_stats_counters(wp, oracle_mod):
    P, I = mg.noisy_sphere(4)
    m = gpu_mesh(wp, P, I, 4)
    Q = wp.array(mg.box_queries(P, 4096, seed=1), dtype=wp.vec3)
    with wp.query_stats() as st:
        wp.mesh_query_point_no_sign(m, Q, 1e6)
        wp.synchronize()
    assert st.pair_fetches > 4096 and st.tri_fetches > 4096
    tree = oracle_mod.mesh_lbvh_build(P, I, 4)
    want = oracle_mod.query_point_no_sign(P, I, tree, Q.numpy(), 1e6, stats=True)
    assert st.tri_fetches == want["tris_tested"]  # same nodes visited in the same order


def test_query_order_modes_agree(wp, oracle_mod):
    """Morton-ordered thread assignment changes who answers, never the answer."""
    P, I = mg.noisy_sphere(5, 0.05, 1)
    m = gpu_mesh(wp, P, I, 4)
    tree = oracle_mod.mesh_lbvh_build(P, I, 4)
    Q = mg.box_queries(P, 50000, seed=8)
    want = oracle_mod.query_point(P, I, tree, Q, 1e6)
    try:
        for mode in (wp.QUERY_ORDER_INPUT, wp.QUERY_ORDER_MORTON, wp.QUERY_ORDER_AUTO):
            wp.set_query_order(mode)
            assert wp.get_query_order() == mode
            assert_results_equal(wp.mesh_query_point(m, wp.array(Q, dtype=wp.vec3), 1e6).numpy(), want, POINT_FIELDS)
            assert_results_equal(wp.mesh_query_point(m, Q, 1e6).numpy(), want, POINT_FIELDS)
            # coincident query points (all keys equal) and a tiny batch
            Qs = np.repeat(Q[:3], 1000, axis=0)
            assert_results_equal(wp.mesh_query_point_no_sign(m, Qs, 1e6).numpy(), oracle_mod.query_point_no_sign(P, I, tree, Qs, 1e6), ("result", "face", "u", "v"))
    finally:
        wp.set_query_order(wp.QUERY_ORDER_AUTO)


def test_c3_rays_on_10m_heightfield(wp, oracle_mod):
    """Config C3 shapes: 10 M-triangle heightfield (deep, duplicate-heavy tree), a 1024 x 1024 pinhole image
    (1 M of the 16.8 M rays) bit-exact on a 100 k sample, plus image-level properties on all of it."""
    P, I = mg.heightfield(2237, 4)
    m = gpu_mesh(wp, P, I, 4)
    S, D = mg.pinhole_rays(1024, 1024)
    got = wp.mesh_query_ray(m, wp.array(S, dtype=wp.vec3), wp.array(D, dtype=wp.vec3), 1e6).numpy()
    hit = got["result"] == 1
    assert 0.3 < hit.mean() < 0.7
    assert (got["t"][hit] > 0).all() and (got["u"][hit] >= -1e-6).all() and (got["v"][hit] >= -1e-6).all()
    assert np.allclose(np.linalg.norm(got["normal"][hit], axis=1), 1.0, atol=1e-5)
    hp = S[hit] + got["t"][hit, None] * D[hit]
    assert (hp[:, 0] > -1e-4).all() and (hp[:, 0] < 1 + 1e-4).all() and (hp[:, 1] > -1e-4).all() and (hp[:, 1] < 1 + 1e-4).all()
    tree = oracle_mod.mesh_lbvh_build(P, I, 4)
    sel = np.random.default_rng(3).integers(0, S.shape[0], 100000)
    want = oracle_mod.query_ray(P, I, tree, S[sel], D[sel], 1e6)
    assert_results_equal({k: got[k][sel] for k in RAY_FIELDS}, want, RAY_FIELDS)
    # closest-point queries near the surface of the same deep tree, signed
    Q = (P[np.random.default_rng(4).integers(0, P.shape[0], 20000)] + np.random.default_rng(5).normal(0, 0.01, (20000, 3))).astype(np.float32)
    assert_results_equal(wp.mesh_query_point(m, Q, 0.05).numpy(), oracle_mod.query_point(P, I, tree, Q, 0.05), POINT_FIELDS)


def test_c4_cloth_refit_query_loop(wp, oracle_mod):
    """Config C4 shape: 4 M-triangle cloth, per frame = deform in place + refit + closest-point queries within
    0.05 of the previous frame's vertices.  3 frames, bit-exact on 20 k sampled queries per frame."""
    n = 1415
    P, I = mg.cloth(n, frame=0)
    pts = wp.array(P, dtype=wp.vec3)
    m = wp.Mesh(pts, wp.array(I, dtype=wp.int32))
    tree = oracle_mod.mesh_lbvh_build(P, I, 4)
    prev = P
    for frame in (1, 2, 3):
        Pf, _ = mg.cloth(n, frame=frame * 40)
        pts.assign(Pf)
        m.refit()
        lo, hi = oracle_mod.triangle_bounds(Pf, I)
        oracle_mod.lbvh_refit(tree, lo, hi)
        rng = np.random.default_rng(5 + frame)
        Q = (prev[rng.integers(0, prev.shape[0], 1 << 20)] + rng.normal(0, 0.01, (1 << 20, 3))).astype(np.float32)
        got = wp.mesh_query_point_no_sign(m, wp.array(Q, dtype=wp.vec3), 0.05).numpy()
        assert 0.5 < got["result"].mean() <= 1.0
        sel = rng.integers(0, Q.shape[0], 20000)
        want = oracle_mod.query_point_no_sign(Pf, I, tree, Q[sel], 0.05)
        assert_results_equal({k: got[k][sel] for k in ("result", "face", "u", "v")}, want, ("result", "face", "u", "v"))
        prev = Pf


def test_ray_order_modes_agree(wp, oracle_mod):
    """Origin/direction ordering of a ray batch changes who traces a ray, never the answer."""
    P, I = mg.noisy_sphere(5, 0.05, 1)
    m = gpu_mesh(wp, P, I, 4)
    tree = oracle_mod.mesh_lbvh_build(P, I, 4)
    S, D = mg.random_rays(P, 60000, seed=21)
    D[:100] = np.array([0, 0, -1], np.float32)  # zero components
    S[100:200] = S[100]                        # shared origin
    want = oracle_mod.query_ray(P, I, tree, S, D, 1e6)
    try:
        for mode in (wp.QUERY_ORDER_INPUT, wp.QUERY_ORDER_MORTON):
            wp.set_ray_order(mode)
            assert wp.get_ray_order() == mode
            assert_results_equal(wp.mesh_query_ray(m, wp.array(S, dtype=wp.vec3), wp.array(D, dtype=wp.vec3), 1e6).numpy(), want, RAY_FIELDS)
            assert_results_equal(wp.mesh_query_ray(m, S, D, 1e6).numpy(), want, RAY_FIELDS)
    finally:
        wp.set_ray_order(wp.QUERY_ORDER_INPUT)


# ------------------------------------------------------------------------------------------------
# any-hit / intersection count / eval_position, eval_velocity
# ------------------------------------------------------------------------------------------------
def test_ray_variants_reference_fixture(wp):
    """Answers recorded from the reference C++ (tests/golden/golden_ray_variants.npz)."""
    g = np.load(GOLD)
    rv = np.load(os.path.join(os.path.dirname(GOLD), "golden_ray_variants.npz"))
    P, I = g["mesh_points"], g["mesh_indices"]
    for leaf in (1, 4):
        m = gpu_mesh(wp, P, I, leaf)
        for mt in (0.5, 1.0, 1e6):
            got = wp.mesh_query_ray_anyhit(m, rv["starts"], rv["dirs"], mt)
            assert got.dtype == bool and np.array_equal(got, rv[f"lbvh{leaf}_anyhit_{mt:g}"].astype(bool))
        assert np.array_equal(wp.mesh_query_ray_count_intersections(m, rv["starts"], rv["dirs"]), rv[f"lbvh{leaf}_count"])
        pos = wp.mesh_eval_position(m, rv["eval_face"], rv["eval_u"], rv["eval_v"])
        assert np.array_equal(pos, rv[f"lbvh{leaf}_eval_position"])


@pytest.mark.parametrize("leaf", [1, 4, 8])
def test_ray_variants_bit_exact(wp, oracle_mod, leaf):
    P, I = mg.noisy_sphere(5, 0.05, 3)
    m = gpu_mesh(wp, P, I, leaf)
    tree = oracle_mod.mesh_lbvh_build(P, I, leaf)
    S, D = mg.random_rays(P, 30000, seed=4)
    S[::3] *= np.float32(0.2)  # a third of the rays start inside
    D[:60] = np.eye(3, dtype=np.float32)[np.arange(60) % 3] * np.float32(-1.0) ** (np.arange(60) // 3)[:, None].astype(np.float32)
    for mt in (0.6, 1e6):
        got = wp.mesh_query_ray_anyhit(m, S, D, mt)
        assert np.array_equal(got, oracle_mod.query_ray_anyhit(P, I, tree, S, D, mt).astype(bool))
        assert np.array_equal(got, wp.mesh_query_ray(m, S, D, mt).result.astype(bool))
    cnt = wp.mesh_query_ray_count_intersections(m, S, D)
    assert np.array_equal(cnt, oracle_mod.query_ray_count(P, I, tree, S, D))
    inside = np.linalg.norm(S, axis=1) < 0.8
    assert inside.sum() > 5000 and np.all(cnt[inside] % 2 == 1) and np.all(cnt[~inside & (np.linalg.norm(S, axis=1) > 1.3)] % 2 == 0)
    # device arrays in -> device arrays out
    ds, dd = wp.array(S, dtype=wp.vec3), wp.array(D, dtype=wp.vec3)
    assert np.array_equal(wp.mesh_query_ray_count_intersections(m, ds, dd).numpy(), cnt)
    assert np.array_equal(wp.mesh_query_ray_anyhit(m, ds, dd, 1e6).numpy().astype(bool), got)


def test_mesh_eval_follows_point_and_velocity_updates(wp, oracle_mod):
    P, I = mg.noisy_sphere(3, 0.05, 9)
    rng = np.random.default_rng(10)
    vel = rng.standard_normal(P.shape).astype(np.float32)
    nt = len(I.reshape(-1, 3))
    F = rng.integers(0, nt, 5000).astype(np.int32)
    U = rng.random(5000).astype(np.float32)
    V = ((1 - U) * rng.random(5000)).astype(np.float32)
    m = wp.Mesh(wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32), bvh_constructor="lbvh")
    assert np.array_equal(wp.mesh_eval_position(m, F, U, V), oracle_mod.mesh_eval(P, I, F, U, V))
    assert np.array_equal(wp.mesh_eval_velocity(m, F, U, V), np.zeros((5000, 3), np.float32))  # mesh.h:2791-2792
    m = wp.Mesh(wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32), velocities=wp.array(vel, dtype=wp.vec3),
                bvh_constructor="lbvh")
    assert np.array_equal(wp.mesh_eval_velocity(m, F, U, V), oracle_mod.mesh_eval(vel, I, F, U, V))
    vel2 = (vel * np.float32(-2.0)).astype(np.float32)
    m.velocities = wp.array(vel2, dtype=wp.vec3)
    assert np.array_equal(wp.mesh_eval_velocity(m, F, U, V), oracle_mod.mesh_eval(vel2, I, F, U, V))
    P2 = (P * np.float32(1.5) + np.float32(0.25)).astype(np.float32)
    m.points = wp.array(P2, dtype=wp.vec3)
    assert np.array_equal(wp.mesh_eval_position(m, F, U, V), oracle_mod.mesh_eval(P2, I, F, U, V))
    # closest point of a query evaluates back onto the surface point the query found
    Q = mg.box_queries(P2, 2000, seed=11)
    r = wp.mesh_query_point_no_sign(m, Q, 1e6)
    pos = wp.mesh_eval_position(m, r.face, r.u, r.v)
    want = oracle_mod.query_point_no_sign(P2, I, oracle_mod.mesh_lbvh_build(P2, I, 4), Q, 1e6)
    assert np.array_equal(r.face, want["face"])
    assert np.all(np.linalg.norm(pos - Q, axis=1) <= np.linalg.norm(P2, axis=1).max() * 2)


def test_ray_variants_empty_inputs(wp):
    P, I = mg.noisy_sphere(1, 0.0, 1)
    m = gpu_mesh(wp, P, I, 4)
    z = np.zeros((0, 3), np.float32)
    assert wp.mesh_query_ray_anyhit(m, z, z, 1.0).shape == (0,)
    assert wp.mesh_query_ray_count_intersections(m, z, z).shape == (0,)
    assert wp.mesh_eval_position(m, np.zeros(0, np.int32), np.zeros(0, np.float32), np.zeros(0, np.float32)).shape == (0, 3)
    with pytest.raises(RuntimeError):
        wp.mesh_query_ray_anyhit(m, np.zeros((3, 3), np.float32), np.zeros((2, 3), np.float32), 1.0)


def test_signed_query_on_grid_aligned_probes(wp, oracle_mod):
    """The sign probes are walked near child first, the reference walks them in a fixed order; both end on the same
    triangle as long as equal hit distances are broken by the reference's visiting order (query.cu,
    probe_sign_ordered).  Stress exactly that: regular meshes and query lattices aligned with them, so that the axis
    probes run along shared edges and through shared vertices (exact ties in t, the fp64 fallback of the watertight
    test).  Every field, sign included, must equal the reference-order restatement."""
    def box_grid(n):  # closed box [0,1]^3, every face an n x n grid of quads split into two triangles
        g = np.linspace(0.0, 1.0, n + 1, dtype=np.float32)
        verts, tris = [], []
        for axis in range(3):
            for side in (0.0, 1.0):
                base = len(verts)
                for a in g:
                    for b in g:
                        p = [0.0, 0.0, 0.0]
                        p[axis], p[(axis + 1) % 3], p[(axis + 2) % 3] = side, a, b
                        verts.append(p)
                for i in range(n):
                    for j in range(n):
                        v0, v1 = base + i * (n + 1) + j, base + (i + 1) * (n + 1) + j
                        quad = (v0, v1, v1 + 1, v0 + 1)
                        quad = quad if side == 1.0 else quad[::-1]  # outward orientation on both sides
                        tris += [quad[0], quad[1], quad[2], quad[0], quad[2], quad[3]]
        return np.asarray(verts, np.float32), np.asarray(tris, np.int32)

    cases = []
    P, I = box_grid(8)
    lat = np.linspace(-0.25, 1.25, 13, dtype=np.float32)  # hits the face grid lines 0, 0.125, ... exactly
    Q = np.stack(np.meshgrid(lat, lat, lat, indexing="ij"), axis=-1).reshape(-1, 3)
    cases.append((P, I, Q))
    P, I = mg.cloth(65, 3)  # open height field on a 65 x 65 vertex grid over [0,1]^2
    gx = np.linspace(0.0, 1.0, 65, dtype=np.float32)[::4]
    Q = np.stack(np.meshgrid(gx, gx, np.linspace(-0.1, 0.1, 9, dtype=np.float32), indexing="ij"), axis=-1).reshape(-1, 3)
    cases.append((P, I, np.concatenate([Q, P[::7]]).astype(np.float32)))
    for P, I, Q in cases:
        for leaf in (1, 4):
            m = gpu_mesh(wp, P, I, leaf)
            tree = oracle_mod.mesh_lbvh_build(P, I, leaf)
            want = oracle_mod.query_point(P, I, tree, Q, 1e6)
            got = wp.mesh_query_point(m, Q, 1e6).numpy()
            assert_results_equal(got, want, POINT_FIELDS)
            assert (want["sign"] < 0).sum() > 10 and (want["sign"] > 0).sum() > 10


def test_sign_parity(wp, oracle_mod):
    """mesh_query_point_sign_parity: bit-exact against the restatement in the order of the reference's device builds
    (offsets drawn x, y, z), on a closed mesh and on one with every other face removed."""
    P, I = mg.noisy_sphere(4, 0.05, 51)
    I_open = I.reshape(-1, 3)[::2].reshape(-1).copy()
    Q = mg.box_queries(P, 20000, seed=52)
    for idx in (I, I_open):
        m = gpu_mesh(wp, P, idx, 4)
        tree = oracle_mod.mesh_lbvh_build(P, idx, 4)
        for ns, sc, md in ((1, 0.1, 1e6), (3, 0.1, 1e6), (4, 0.5, 1e6), (0, 0.1, 1e6), (1, 0.1, 0.05)):
            want = oracle_mod.query_point_sign_parity(P, idx, tree, Q, md, ns, sc, rtl=False)
            got = wp.mesh_query_point_sign_parity(m, Q, md, ns, sc)
            assert_results_equal(got.numpy(), want, POINT_FIELDS)
    # closed mesh: same inside / outside answer as the three-axis-probe sign of mesh_query_point
    m = gpu_mesh(wp, P, I, 4)
    a = wp.mesh_query_point_sign_parity(m, wp.array(Q, dtype=wp.vec3), 1e6).numpy()
    b = wp.mesh_query_point(m, Q, 1e6).numpy()
    assert np.array_equal(a["sign"], b["sign"]) and np.array_equal(a["face"], b["face"])
    assert 0.05 < (a["sign"] < 0).mean() < 0.6


def test_sign_normal(wp, oracle_mod):
    """mesh_query_point_sign_normal: result / face / u / v bit-exact against the restatement (pinned on the reference
    C++); the sign may differ only where CUDA's acosf and glibc's round a corner angle differently AND the accumulated
    normal is perpendicular to the offset -- counted, and bounded at 0.1 %.  average_edge_length: float terms of
    mesh.cu:53 summed in double."""
    P, I = mg.noisy_sphere(4, 0.05, 53)
    rng = np.random.default_rng(54)
    T = I.reshape(-1, 3)
    tri = T[rng.integers(0, len(T), 3000)]
    Q = np.concatenate([mg.box_queries(P, 20000, seed=55), P[rng.integers(0, len(P), 3000)] + rng.normal(0, 1e-4, (3000, 3)),
                        0.5 * (P[tri[:, 0]] + P[tri[:, 1]]), P[:2000]]).astype(np.float32)
    m = gpu_mesh(wp, P, I, 4)
    tree = oracle_mod.mesh_lbvh_build(P, I, 4)
    avg = wp.mesh_average_edge_length(m)
    assert np.float32(avg) == np.float32(oracle_mod.average_edge_length(P, I, mode=1))
    fields = [f for f in POINT_FIELDS if f != "sign"]
    for eps, md in ((1e-3, 1e6), (0.1, 1e6), (0.0, 1e6), (1e-3, 0.05)):
        want = oracle_mod.query_point_sign_normal(P, I, tree, Q, md, avg, eps)
        got = wp.mesh_query_point_sign_normal(m, Q, md, eps).numpy()
        assert_results_equal(got, want, fields)
        print("sign_normal eps=%g max_dist=%g: %d sign differences vs the oracle" % (eps, md, int((got["sign"] != want["sign"]).sum())))
        assert (got["sign"] != want["sign"]).mean() <= 1e-3
        assert np.array_equal(got["sign"] == 0, want["result"] == 0)
    # closed mesh: the normal sign is the ray-vote sign of mesh_query_point except next to the surface
    # (test_mesh_query_point.py:320-570 asks the same of all sign methods)
    votes = wp.mesh_query_point(m, Q[:20000], 1e6).numpy()
    normal = wp.mesh_query_point_sign_normal(m, Q[:20000], 1e6).numpy()
    assert (votes["face"] == normal["face"]).mean() > 0.99  # (faces may differ at shared edges: distances vs squared distances)
    print("sign_normal vs ray votes: %d of 20000 differ" % int((votes["sign"] != normal["sign"]).sum()))
    assert (votes["sign"] == normal["sign"]).mean() > 0.995
    # Warp kernels read wp::Mesh::average_edge_length through the id (mesh.h:889): the field (offset 320 of the 328-byte
    # descriptor) is refreshed with the reference-layout arrays
    import ctypes

    from warp_b200 import _lib

    m.download_tree()  # -> wp_b200_bvh_sync_reference_layout
    wp.synchronize()
    field = np.zeros(1, np.float32)
    assert _lib.core().wp_memcpy_d2h(None, ctypes.c_void_p(field.ctypes.data), ctypes.c_void_p(m.id + 320), 4, None)
    wp.synchronize()
    assert field[0] == np.float32(avg)
    # device arrays in -> device arrays out, unordered batch path (< 32768 queries) and the ordered one agree
    a = wp.mesh_query_point_sign_normal(m, wp.array(Q[:1000], dtype=wp.vec3), 1e6).numpy()
    b = wp.mesh_query_point_sign_normal(m, np.tile(Q[:1000], (40, 1)), 1e6).numpy()
    assert all(np.array_equal(a[k], b[k][:1000]) for k in POINT_FIELDS)
    # the average follows the points: refit after scaling the mesh by 2
    pts = wp.array(P, dtype=wp.vec3)
    m2 = wp.Mesh(pts, wp.array(I, dtype=wp.int32), bvh_constructor="lbvh")
    pts.assign((2.0 * P).astype(np.float32))
    m2.refit()
    assert abs(wp.mesh_average_edge_length(m2) - 2.0 * avg) <= 1e-5 * avg
    assert wp.mesh_query_point_sign_normal(m2, np.zeros((0, 3), np.float32), 1.0).numpy()["face"].shape == (0,)


def test_furthest_point_and_face_normal(wp, oracle_mod):
    """mesh_query_furthest_point_no_sign and mesh_eval_face_normal: bit-exact against the restatement (pinned on the
    reference C++), leaf sizes 1 / 4, several min_dist, empty batch, after a refit."""
    P, I = mg.noisy_sphere(4, 0.05, 57)
    Q = mg.box_queries(P, 20000, seed=58)
    fields = ("result", "face", "u", "v")
    for leaf in (1, 4):
        m = gpu_mesh(wp, P, I, leaf)
        tree = oracle_mod.mesh_lbvh_build(P, I, leaf)
        for md in (0.0, 1.5, 2.0, 1e6):
            want = oracle_mod.query_furthest_point_no_sign(P, I, tree, Q, md)
            got = wp.mesh_query_furthest_point_no_sign(m, Q, md).numpy()
            assert_results_equal(got, want, fields)
            assert not got["sign"].any()
    dev = wp.mesh_query_furthest_point_no_sign(m, wp.array(Q, dtype=wp.vec3), 0.0).numpy()
    assert_results_equal(dev, oracle_mod.query_furthest_point_no_sign(P, I, tree, Q, 0.0), fields)
    assert wp.mesh_query_furthest_point_no_sign(m, np.zeros((0, 3), np.float32), 0.0).numpy()["face"].shape == (0,)
    f = np.arange(len(I) // 3, dtype=np.int32)
    assert np.array_equal(wp.mesh_eval_face_normal(m, f), oracle_mod.mesh_face_normal(P, I, f))
    # both follow the points: refit after a deformation
    P2 = mg.renoise_sphere(P, 0.05, 59)
    pts = wp.array(P, dtype=wp.vec3)
    m2 = wp.Mesh(pts, wp.array(I, dtype=wp.int32), bvh_constructor="lbvh")
    pts.assign(P2)
    m2.refit()
    lo2, hi2 = oracle_mod.triangle_bounds(P2, I)
    oracle_mod.lbvh_refit(tree, lo2, hi2)
    assert_results_equal(wp.mesh_query_furthest_point_no_sign(m2, Q, 0.0).numpy(),
                         oracle_mod.query_furthest_point_no_sign(P2, I, tree, Q, 0.0), fields)
    assert np.array_equal(wp.mesh_eval_face_normal(m2, wp.array(f, dtype=wp.int32)).numpy(), oracle_mod.mesh_face_normal(P2, I, f))


def test_rooted_mesh_rays(wp, oracle_mod):
    """mesh_query_ray / _anyhit / _count_intersections restricted to a group's subtree (`root` argument of the
    reference): fixture from the reference C++, then the oracle; group roots of a grouped MESH."""
    from test_oracle import _grouped_mesh_case

    g = np.load(os.path.join(os.path.dirname(GOLD), "golden_group_queries.npz"))
    P, I, T, groups, S, D, gid = _grouped_mesh_case()
    m = wp.Mesh(wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32), groups=wp.array(groups, dtype=wp.int32),
                bvh_constructor="lbvh", bvh_leaf_size=4)
    roots = wp.bvh_get_group_root(m, gid)
    assert np.array_equal(roots, g["mesh_roots"])
    Sd, Dd, Rd = wp.array(S, dtype=wp.vec3), wp.array(D, dtype=wp.vec3), wp.array(roots, dtype=wp.int32)
    got = wp.mesh_query_ray(m, Sd, Dd, 1e6, roots=Rd).numpy()
    assert_results_equal(got, {k: g[f"mesh_ray_{k}"] for k in RAY_FIELDS}, RAY_FIELDS)
    assert np.array_equal(wp.mesh_query_ray_anyhit(m, Sd, Dd, 0.9, roots=Rd).numpy(), g["mesh_anyhit"])
    assert np.array_equal(wp.mesh_query_ray_count_intersections(m, Sd, Dd, roots=roots).numpy(), g["mesh_count"])
    hit = (got["result"] == 1) & (roots != -1)
    assert hit.sum() > 100 and np.array_equal(groups[got["face"][hit]], gid[hit])
    # leaf size 1 and a bigger batch against the oracle
    m1 = wp.Mesh(wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32), groups=wp.array(groups, dtype=wp.int32),
                 bvh_constructor="lbvh", bvh_leaf_size=1)
    tree = oracle_mod.mesh_lbvh_build(P, I, 1, groups=groups)
    S2, D2 = mg.random_rays(P, 20000, seed=64)
    gid2 = np.random.default_rng(65).integers(0, 5, 20000).astype(np.int32)
    r2 = wp.bvh_get_group_root(m1, gid2)
    assert np.array_equal(r2, oracle_mod.bvh_group_roots(tree, groups, gid2))
    got2 = wp.mesh_query_ray(m1, wp.array(S2, dtype=wp.vec3), wp.array(D2, dtype=wp.vec3), 1e6, roots=r2).numpy()
    assert_results_equal(got2, oracle_mod.query_ray(P, I, tree, S2, D2, 1e6, roots=r2), RAY_FIELDS)
