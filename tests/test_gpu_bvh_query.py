"""GPU parity of the batched generic wp.Bvh queries against the restatement of the reference iterator
(exact hit lists IN ITERATOR ORDER) and against numpy brute force (exact sets), following the procedure of
warp/tests/geometry/test_bvh.py:186-262: build, query, refit, query, rebuild, query."""
import numpy as np
import pytest

from helpers import random_boxes
from test_oracle import _brute_aabb, _brute_ray

pytestmark = pytest.mark.gpu


def _check(wp, oracle_mod, bvh, tree, lo, hi, qlo, qhi, s, d):
    off, idx = wp.bvh_query_aabb(bvh, qlo, qhi).numpy()
    woff, widx = oracle_mod.bvh_query(tree, lo, hi, qlo, qhi)
    assert np.array_equal(off, woff) and np.array_equal(idx, widx)
    for i, want in enumerate(_brute_aabb(lo, hi, qlo, qhi)):
        assert sorted(idx[off[i] : off[i + 1]].tolist()) == want.tolist()
    for md in (3.4028234663852886e38, 4.0):
        off, idx = wp.bvh_query_ray(bvh, wp.array(s, dtype=wp.vec3), wp.array(d, dtype=wp.vec3), md).numpy()
        woff, widx = oracle_mod.bvh_query(tree, lo, hi, s, d, ray=True, max_dist=md)
        assert np.array_equal(off, woff) and np.array_equal(idx, widx)
        for i, want in enumerate(_brute_ray(lo, hi, s, d, np.float32(md))):
            assert sorted(idx[off[i] : off[i + 1]].tolist()) == want.tolist()


@pytest.mark.parametrize("leaf", [1, 2, 4])
def test_bvh_queries_build_refit_rebuild(wp, oracle_mod, leaf):
    lo, hi = random_boxes(100, seed=123)
    lo_d, hi_d = wp.array(lo, dtype=wp.vec3), wp.array(hi, dtype=wp.vec3)
    bvh = wp.Bvh(lo_d, hi_d, constructor="lbvh", leaf_size=leaf)
    tree = oracle_mod.lbvh_build(lo, hi, leaf)
    rng = np.random.default_rng(5)
    qlo = (rng.random((257, 3)) * 10).astype(np.float32)
    qhi = (qlo + rng.random((257, 3)).astype(np.float32) * 3).astype(np.float32)
    s = (rng.random((257, 3)) * 10).astype(np.float32)
    d = rng.standard_normal((257, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    _check(wp, oracle_mod, bvh, tree, lo, hi, qlo, qhi, s, d)
    lo2, hi2 = random_boxes(100, seed=124)
    lo_d.assign(lo2), hi_d.assign(hi2)
    bvh.refit()
    oracle_mod.lbvh_refit(tree, lo2, hi2)
    _check(wp, oracle_mod, bvh, tree, lo2, hi2, qlo, qhi, s, d)
    bvh.rebuild()
    _check(wp, oracle_mod, bvh, oracle_mod.lbvh_build(lo2, hi2, leaf), lo2, hi2, qlo, qhi, s, d)


def test_bvh_queries_large_and_edge_cases(wp, oracle_mod):
    lo, hi = random_boxes(200000, seed=9, extent=50.0)
    bvh = wp.Bvh(wp.array(lo, dtype=wp.vec3), wp.array(hi, dtype=wp.vec3), leaf_size=4)
    tree = oracle_mod.lbvh_build(lo, hi, 4)
    rng = np.random.default_rng(6)
    qlo = (rng.random((50000, 3)) * 50).astype(np.float32)
    qhi = (qlo + rng.random((50000, 3)).astype(np.float32) * 2).astype(np.float32)
    res = wp.bvh_query_aabb(bvh, qlo, qhi)
    off, idx = res.numpy()
    woff, widx = oracle_mod.bvh_query(tree, lo, hi, qlo, qhi)
    assert res.total == int(woff[-1]) > 50000
    assert np.array_equal(off, woff) and np.array_equal(idx, widx)
    # no hits at all, empty batch, single-item tree
    far = np.full((7, 3), 1e6, np.float32)
    r = wp.bvh_query_aabb(bvh, far, far + 1)
    assert r.total == 0 and not r.numpy()[0].any()
    assert wp.bvh_query_aabb(bvh, np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32)).total == 0
    one = wp.Bvh(wp.array(lo[:1], dtype=wp.vec3), wp.array(hi[:1], dtype=wp.vec3))
    r = wp.bvh_query_aabb(one, lo[:1] - 1, hi[:1] + 1)
    assert r.lists()[0].tolist() == [0]
    with pytest.raises(TypeError):
        wp.bvh_query_aabb(object(), far, far)


def test_mesh_query_aabb(wp, oracle_mod):
    """wp.mesh_query_aabb batched: reference fixture (iterator order), oracle on a bigger mesh, refit follows the points."""
    import os
    from warp_b200 import meshgen as mg

    gdir = os.path.join(os.path.dirname(__file__), "golden")
    g, rv = np.load(os.path.join(gdir, "golden_cpu.npz")), np.load(os.path.join(gdir, "golden_ray_variants.npz"))
    P, I = g["mesh_points"], g["mesh_indices"]
    for leaf in (1, 4):
        m = wp.Mesh(wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32), bvh_constructor="lbvh", bvh_leaf_size=leaf)
        off, idx = wp.mesh_query_aabb(m, rv["aabb_lowers"], rv["aabb_uppers"]).numpy()
        assert np.array_equal(off, rv[f"lbvh{leaf}_aabb_offsets"]) and np.array_equal(idx, rv[f"lbvh{leaf}_aabb_indices"])

    P, I = mg.noisy_sphere(5, 0.05, 21)
    pts = wp.array(P, dtype=wp.vec3)
    m = wp.Mesh(pts, wp.array(I, dtype=wp.int32), bvh_constructor="lbvh", bvh_leaf_size=4)
    tree = oracle_mod.mesh_lbvh_build(P, I, 4)
    rng = np.random.default_rng(22)
    qlo = (rng.random((3000, 3)) * 2.4 - 1.3).astype(np.float32)
    qhi = (qlo + rng.random((3000, 3)).astype(np.float32) * 0.3).astype(np.float32)
    tlo, thi = oracle_mod.triangle_bounds(P, I)
    off, idx = wp.mesh_query_aabb(m, wp.array(qlo, dtype=wp.vec3), wp.array(qhi, dtype=wp.vec3)).numpy()
    woff, widx = oracle_mod.bvh_query(tree, tlo, thi, qlo, qhi)
    assert off[-1] > 10000 and np.array_equal(off, woff) and np.array_equal(idx, widx)
    for i, want in enumerate(_brute_aabb(tlo, thi, qlo[:200], qhi[:200])):
        assert sorted(idx[off[i] : off[i + 1]].tolist()) == want.tolist()
    # move the points, refit: the item boxes the query tests are the refreshed ones
    P2 = (P * np.float32(1.25)).astype(np.float32)
    pts.assign(P2)
    m.refit()
    tlo2, thi2 = oracle_mod.triangle_bounds(P2, I)
    off2, idx2 = wp.mesh_query_aabb(m, qlo, qhi).numpy()
    for i, want in enumerate(_brute_aabb(tlo2, thi2, qlo[:300], qhi[:300])):
        assert sorted(idx2[off2[i] : off2[i + 1]].tolist()) == want.tolist()
    with pytest.raises(TypeError):
        wp.mesh_query_aabb(wp.Bvh(wp.array(tlo, dtype=wp.vec3), wp.array(thi, dtype=wp.vec3)), qlo, qhi)


@pytest.mark.parametrize("leaf", [1, 4])
def test_group_roots_and_rooted_queries(wp, oracle_mod, leaf):
    """wp.bvh_get_group_root + root= traversal (warp/tests/geometry/test_grouped_bvh.py): fixture from the
    reference's header code, then the oracle on a larger grouped tree, before and after a refit."""
    import os
    from test_oracle import _group_case

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_group_queries.npz"))
    lo, hi, groups, qlo, qhi, s, d, gid = _group_case(900 + leaf, 400, 6)
    bvh = wp.Bvh(wp.array(lo, dtype=wp.vec3), wp.array(hi, dtype=wp.vec3), groups=wp.array(groups, dtype=wp.int32), leaf_size=leaf)
    roots = wp.bvh_get_group_root(bvh, gid)
    assert np.array_equal(roots, g[f"leaf{leaf}_roots"])
    off, idx = wp.bvh_query_aabb(bvh, qlo, qhi, roots=roots).numpy()
    assert np.array_equal(off, g[f"leaf{leaf}_aabb_offsets"]) and np.array_equal(idx, g[f"leaf{leaf}_aabb_indices"])
    off, idx = wp.bvh_query_ray(bvh, s, d, 6.0, roots=wp.array(roots, dtype=wp.int32)).numpy()
    assert np.array_equal(off, g[f"leaf{leaf}_ray_offsets"]) and np.array_equal(idx, g[f"leaf{leaf}_ray_indices"])

    lo, hi, groups, qlo, qhi, s, d, gid = _group_case(77 + leaf, 20000, 37)
    lo_d, hi_d = wp.array(lo, dtype=wp.vec3), wp.array(hi, dtype=wp.vec3)
    bvh = wp.Bvh(lo_d, hi_d, groups=wp.array(groups, dtype=wp.int32), leaf_size=leaf)
    tree = oracle_mod.lbvh_build(lo, hi, leaf, groups=groups)
    roots = wp.bvh_get_group_root(bvh, wp.array(gid, dtype=wp.int32)).numpy()
    assert np.array_equal(roots, oracle_mod.bvh_group_roots(tree, groups, gid))
    for r in (roots, np.full(64, -1, np.int32)):
        off, idx = wp.bvh_query_aabb(bvh, qlo, qhi, roots=r).numpy()
        woff, widx = oracle_mod.bvh_query(tree, lo, hi, qlo, qhi, roots=r)
        assert np.array_equal(off, woff) and np.array_equal(idx, widx)
    lo2, hi2 = (lo + np.float32(0.5)).astype(np.float32), (hi + np.float32(0.75)).astype(np.float32)
    lo_d.assign(lo2), hi_d.assign(hi2)
    bvh.refit()
    oracle_mod.lbvh_refit(tree, lo2, hi2)
    off, idx = wp.bvh_query_ray(bvh, s, d, 8.0, roots=roots).numpy()
    woff, widx = oracle_mod.bvh_query(tree, lo2, hi2, s, d, ray=True, max_dist=8.0, roots=roots)
    assert np.array_equal(off, woff) and np.array_equal(idx, widx)
    # an ungrouped tree: every item is in group 0
    plain = wp.Bvh(lo_d, hi_d, leaf_size=leaf)
    info_root = plain.download_tree()["root"]
    assert wp.bvh_get_group_root(plain, np.array([0, 1, -3], np.int32)).tolist() == [info_root, -1, -1]


@pytest.mark.parametrize("leaf", [1, 2, 4])
def test_bvh_sphere_and_capsule_queries(wp, oracle_mod, leaf):
    """bvh_query_sphere / bvh_query_capsule: exact hit lists in iterator order against the restatement (pinned on the
    reference C++), build -> refit -> rebuild like test_bvh.py:186-262; per-query and scalar radii, negative radii,
    axis-aligned directions, closed max_dist, roots, empty batch."""
    from test_oracle import _brute_sphere

    lo, hi = random_boxes(3000, seed=131)
    lo_d, hi_d = wp.array(lo, dtype=wp.vec3), wp.array(hi, dtype=wp.vec3)
    bvh = wp.Bvh(lo_d, hi_d, constructor="lbvh", leaf_size=leaf)
    tree = oracle_mod.lbvh_build(lo, hi, leaf)
    rng = np.random.default_rng(132)
    n = 1500
    C = (rng.random((n, 3)) * 10).astype(np.float32)
    R = (rng.random(n) * 1.5 - 0.1).astype(np.float32)
    D = rng.standard_normal((n, 3)).astype(np.float32)
    D[::7, 0] = 0
    D[::11, 1] = 0
    D[::13] = (0, 0, 1)
    D /= np.linalg.norm(D, axis=1, keepdims=True)

    def check(tree, lo, hi):
        off, idx = wp.bvh_query_sphere(bvh, C, R).numpy()
        woff, widx = oracle_mod.bvh_query_kind(tree, lo, hi, "sphere", C, radii=R)
        assert np.array_equal(off, woff) and np.array_equal(idx, widx)
        if leaf > 1:
            for i, want in enumerate(_brute_sphere(lo, hi, C[:200], R[:200])):
                assert sorted(idx[off[i] : off[i + 1]].tolist()) == want.tolist()
        off, idx = wp.bvh_query_sphere(bvh, wp.array(C, dtype=wp.vec3), 0.75).numpy()
        woff, widx = oracle_mod.bvh_query_kind(tree, lo, hi, "sphere", C, radii=0.75)
        assert np.array_equal(off, woff) and np.array_equal(idx, widx)
        for md in (3.4028234663852886e38, 3.0):
            off, idx = wp.bvh_query_capsule(bvh, C, D, R, md).numpy()
            woff, widx = oracle_mod.bvh_query_kind(tree, lo, hi, "capsule", C, D, radii=R, max_dist=md)
            assert np.array_equal(off, woff) and np.array_equal(idx, widx)
        roots = np.full(n, -1, np.int32)
        roots[::2] = tree["root"]
        off, idx = wp.bvh_query_capsule(bvh, C, D, 0.3, 2.0, roots=roots).numpy()
        woff, widx = oracle_mod.bvh_query_kind(tree, lo, hi, "capsule", C, D, radii=0.3, max_dist=2.0, roots=roots)
        assert np.array_equal(off, woff) and np.array_equal(idx, widx)

    check(tree, lo, hi)
    lo2, hi2 = random_boxes(3000, seed=133)
    lo_d.assign(lo2), hi_d.assign(hi2)
    bvh.refit()
    oracle_mod.lbvh_refit(tree, lo2, hi2)
    check(tree, lo2, hi2)
    bvh.rebuild()
    check(oracle_mod.lbvh_build(lo2, hi2, leaf), lo2, hi2)
    z = np.zeros((0, 3), np.float32)
    assert wp.bvh_query_sphere(bvh, z, 1.0).total == 0 and wp.bvh_query_capsule(bvh, z, z, 1.0).total == 0
    with pytest.raises(RuntimeError):
        wp.bvh_query_sphere(bvh, C, R[:5])


@pytest.mark.parametrize("leaf", [1, 4])
def test_mesh_query_sphere(wp, oracle_mod, leaf):
    """mesh_query_sphere: exact face lists in iterator order against the restatement (pinned on the reference C++),
    incl. zero-area faces, scalar and per-query radii, after a refit, empty batch; fixture from the reference."""
    import os

    from warp_b200 import meshgen as mg

    P, I = mg.noisy_sphere(4, 0.05, 141)
    I = np.concatenate([I, np.array([0, 0, 5, 3, 7, 7, 10, 10, 10], np.int32)]).astype(np.int32)
    rng = np.random.default_rng(142)
    C = np.concatenate([mg.box_queries(P, 6000, seed=143), P[:500]]).astype(np.float32)
    R = (rng.random(len(C)) * 0.3 - 0.01).astype(np.float32)
    pts = wp.array(P, dtype=wp.vec3)
    m = wp.Mesh(pts, wp.array(I, dtype=wp.int32), bvh_constructor="lbvh", bvh_leaf_size=leaf)
    tree = oracle_mod.mesh_lbvh_build(P, I, leaf)
    for radii in (R, 0.1):
        off, idx = wp.mesh_query_sphere(m, C, radii).numpy()
        woff, widx = oracle_mod.mesh_query_sphere(P, I, tree, C, radii)
        assert int(woff[-1]) > 1000 and np.array_equal(off, woff) and np.array_equal(idx, widx)
    P2 = mg.renoise_sphere(P, 0.05, 144)
    pts.assign(P2)
    m.refit()
    lo2, hi2 = oracle_mod.triangle_bounds(P2, I)
    oracle_mod.lbvh_refit(tree, lo2, hi2)
    off, idx = wp.mesh_query_sphere(m, wp.array(C, dtype=wp.vec3), wp.array(R, dtype=wp.float32)).numpy()
    woff, widx = oracle_mod.mesh_query_sphere(P2, I, tree, C, R)
    assert np.array_equal(off, woff) and np.array_equal(idx, widx)
    assert wp.mesh_query_sphere(m, np.zeros((0, 3), np.float32), 1.0).total == 0
    # the reference's own answers (fixture)
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_bvh_kinds.npz"))
    gc = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_cpu.npz"))
    gm = wp.Mesh(wp.array(gc["mesh_points"], dtype=wp.vec3), wp.array(g["mesh_indices"], dtype=wp.int32),
                 bvh_constructor="lbvh", bvh_leaf_size=leaf)
    off, idx = wp.mesh_query_sphere(gm, g["mesh_centers"], g["mesh_radii"]).numpy()
    assert np.array_equal(off, g[f"mesh_leaf{leaf}_sphere_offsets"]) and np.array_equal(idx, g[f"mesh_leaf{leaf}_sphere_indices"])
