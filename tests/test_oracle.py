"""CPU tests: the oracle restatement (oracle/lbvh_oracle.c) against the reference's golden vectors,
the committed fixtures produced by the reference C++ (tests/golden/golden_cpu.npz), and -- when
oracle/_ref is present -- the reference C++ itself on fresh random inputs."""
import os

import numpy as np
import pytest

from helpers import assert_results_equal, random_boxes, visible_nodes
from warp_b200 import meshgen as mg

GOLD = os.path.join(os.path.dirname(__file__), "golden", "golden_cpu.npz")
POINT_FIELDS = ("result", "sign", "face", "u", "v")
RAY_FIELDS = ("result", "sign", "face", "t", "u", "v", "normal")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_cube_golden_vectors(oracle_mod):
    """warp/tests/geometry/test_mesh.py:111-187: face 1 / pos (0.1,0.2,0.5) / sign -+1; ray t 0.557828, face 4."""
    o = oracle_mod
    d = np.array([-1.2, 2.3, -3.4], np.float32)
    d /= np.linalg.norm(d)
    for idx, sgn in ((mg.CUBE_INDICES_RH, -1.0), (mg.CUBE_INDICES_LH, 1.0)):
        for leaf in (1, 2, 4):
            tree = o.mesh_lbvh_build(mg.CUBE_POINTS, idx, leaf)
            p = o.query_point(mg.CUBE_POINTS, idx, tree, [[0.1, 0.2, 0.3]], 1e6)
            assert p["result"][0] == 1 and p["face"][0] == 1 and p["sign"][0] == sgn
            tri = mg.CUBE_POINTS[idx.reshape(-1, 3)[p["face"][0]]]
            u, v = p["u"][0], p["v"][0]
            pos = u * tri[0] + v * tri[1] + (1 - u - v) * tri[2]
            assert np.allclose(pos, (0.1, 0.2, 0.5), atol=1e-6)
            r = o.query_ray(mg.CUBE_POINTS, idx, tree, [[0.1, 0.2, 0.3]], [d], 1e6)
            assert r["result"][0] == 1 and r["face"][0] == 4
            assert abs(r["t"][0] - 0.557828) < 1e-6
            assert np.sign(r["sign"][0]) == sgn
            tri = mg.CUBE_POINTS[idx.reshape(-1, 3)[4]]
            u, v = r["u"][0], r["v"][0]
            pos = u * tri[0] + v * tri[1] + (1 - u - v) * tri[2]
            assert np.allclose(pos, (-0.0565217, 0.5, -0.143478), atol=1e-6)


def test_cube_matches_reference_fixture(oracle_mod, gold):
    """LBVH-tree answers of the oracle == reference SAH/median-tree answers (tree independent fields)."""
    o = oracle_mod
    d = np.array([-1.2, 2.3, -3.4], np.float32)
    d /= np.linalg.norm(d)
    for name, idx in (("rh", mg.CUBE_INDICES_RH), ("lh", mg.CUBE_INDICES_LH)):
        tree = o.mesh_lbvh_build(mg.CUBE_POINTS, idx, 4)
        p = o.query_point(mg.CUBE_POINTS, idx, tree, [[0.1, 0.2, 0.3]], 1e6)
        r = o.query_ray(mg.CUBE_POINTS, idx, tree, [[0.1, 0.2, 0.3]], [d], 1e6)
        for cname in ("sah", "median"):
            for leaf in (1, 2, 4):
                for f in POINT_FIELDS:
                    assert np.array_equal(p[f], gold[f"cube_{name}_{cname}_{leaf}_point_{f}"]), (name, cname, leaf, f)
                for f in RAY_FIELDS:
                    assert np.array_equal(r[f], gold[f"cube_{name}_{cname}_{leaf}_ray_{f}"]), (name, cname, leaf, f)


def test_traversal_matches_reference_fixture(oracle_mod, gold):
    """Our traversal restatement run on reference-built (SAH) and LBVH trees == the reference's traversal."""
    o = oracle_mod
    P, I, Q, S, D = (gold[k] for k in ("mesh_points", "mesh_indices", "queries", "ray_starts", "ray_dirs"))
    for prefix in ("sah", "lbvh1", "lbvh4"):
        tree = {k: gold[f"{prefix}_tree_{k}"] for k in ("node_lowers", "node_uppers", "primitive_indices")}
        tree["root"] = int(gold[f"{prefix}_tree_root"])
        assert_results_equal(o.query_point(P, I, tree, Q, 1e6), {f: gold[f"{prefix}_point_{f}"] for f in POINT_FIELDS}, POINT_FIELDS)
        assert_results_equal(o.query_point(P, I, tree, Q, 0.5), {f: gold[f"{prefix}_point05_{f}"] for f in POINT_FIELDS}, POINT_FIELDS)
        assert_results_equal(o.query_ray(P, I, tree, S, D, 1e6), {f: gold[f"{prefix}_ray_{f}"] for f in RAY_FIELDS}, RAY_FIELDS)
        ns = o.query_point_no_sign(P, I, tree, Q, 1e6)
        for f in ("result", "face", "u", "v"):
            assert np.array_equal(ns[f], gold[f"{prefix}_point_{f}"])
        assert not ns["sign"].any()


def test_lbvh_build_matches_fixture(oracle_mod, gold):
    """Rebuilding the fixture LBVH reproduces the stored keys / order / topology (regression pin)."""
    o = oracle_mod
    P, I = gold["mesh_points"], gold["mesh_indices"]
    for leaf in (1, 4):
        t = o.mesh_lbvh_build(P, I, leaf)
        for k in ("keys", "primitive_indices", "parents", "node_lowers", "node_uppers"):
            assert np.array_equal(t[k], gold[f"lbvh{leaf}_tree_{k}"]), k
        assert t["root"] == int(gold[f"lbvh{leaf}_tree_root"])


def test_primitives_match_reference_fixture(oracle_mod, gold):
    o = oracle_mod
    uv = np.stack([o.closest_point_to_triangle(t[0], t[1], t[2], p) for t, p in zip(gold["tri_abc"], gold["tri_p"])])
    assert np.array_equal(uv, gold["tri_uv"])
    codes = np.array([o.morton3(*map(float, v)) for v in gold["morton_xyz"]], np.uint32)
    assert np.array_equal(codes, gold["morton_code"])


def _check_tree_invariants(tree, lowers, uppers, leaf_size):
    n = tree["n"]
    lo, hi, par = tree["node_lowers"], tree["node_uppers"], tree["parents"]
    assert np.all(np.diff(tree["keys"].astype(np.int64)) >= 0), "keys sorted"
    assert sorted(tree["primitive_indices"].tolist()) == list(range(n)), "permutation"
    # stable: equal keys keep ascending item order
    same = tree["keys"][1:] == tree["keys"][:-1]
    assert np.all(tree["primitive_indices"][1:][same] > tree["primitive_indices"][:-1][same])
    assert par[tree["root"]] == -1
    vis = visible_nodes(tree)
    covered = []
    for c in vis:
        if lo["ib"][c] >> 31:
            s, e = int(lo["ib"][c] & 0x7FFFFFFF), int(hi["ib"][c] & 0x7FFFFFFF)
            covered.extend(range(s, e))
            items = tree["primitive_indices"][s:e]
            assert np.array_equal([lo["x"][c], lo["y"][c], lo["z"][c]], lowers[items].min(axis=0))
            assert np.array_equal([hi["x"][c], hi["y"][c], hi["z"][c]], uppers[items].max(axis=0))
        else:
            l, r = int(lo["ib"][c] & 0x7FFFFFFF), int(hi["ib"][c] & 0x7FFFFFFF)
            assert par[l] == c and par[r] == c
            for f, agg, half in (("x", min, lo), ("y", min, lo), ("z", min, lo), ("x", max, hi), ("y", max, hi), ("z", max, hi)):
                assert half[f][c] == agg(half[f][l], half[f][r])
    assert sorted(covered) == list(range(n)), "visible leaves partition the items"


@pytest.mark.parametrize("n,leaf", [(1, 1), (2, 1), (2, 2), (3, 4), (100, 1), (100, 4), (1000, 8)])
def test_lbvh_invariants_random_boxes(oracle_mod, n, leaf):
    lo, hi = random_boxes(n, seed=123 + n)
    t = oracle_mod.lbvh_build(lo, hi, leaf)
    _check_tree_invariants(t, lo, hi, leaf)


def test_lbvh_depth_rule_on_duplicates(oracle_mod):
    """Many identical boxes -> equal keys -> parity-driven ladders deeper than 32 -> depth-rule leaves."""
    n = 600
    lo = np.zeros((n, 3), np.float32)
    lo[: n // 2] += 1.0
    hi = lo + 0.5
    t = oracle_mod.lbvh_build(lo, hi, 1)
    _check_tree_invariants(t, lo, hi, 1)
    sizes = [(int(t["node_uppers"]["ib"][c] & 0x7FFFFFFF) - int(t["node_lowers"]["ib"][c] & 0x7FFFFFFF))
             for c in visible_nodes(t) if t["node_lowers"]["ib"][c] >> 31]
    assert max(sizes) > 1, "expected at least one depth-forced leaf holding several items"


def test_refit_matches_rebuild_bounds(oracle_mod):
    o = oracle_mod
    P, I = mg.noisy_sphere(3, seed=5)
    t = o.mesh_lbvh_build(P, I, 4)
    P2 = (P + np.float32(10.0) * np.array([1, 0, 0], np.float32)).astype(np.float32)
    lo2, hi2 = o.triangle_bounds(P2, I)
    o.lbvh_refit(t, lo2, hi2)
    _check_tree_invariants(t, lo2, hi2, 4)
    # ray that hit the old position now misses, ray at the new position hits (test_mesh.py:319-357)
    d = np.array([[0.0, 0.0, -1.0]], np.float32)
    assert o.query_ray(P2, I, t, [[0.0, 0.0, 5.0]], d, 1e6)["result"][0] == 0
    assert o.query_ray(P2, I, t, [[10.0, 0.0, 5.0]], d, 1e6)["result"][0] == 1


def test_live_reference_agreement(oracle_mod):
    """Fresh random inputs through the reference C++ (skipped where oracle/_ref was not shipped)."""
    o = oracle_mod
    if not o.ref_available():
        pytest.skip("oracle/_ref/libwarp_ref_cpu.so not present")
    P, I = mg.noisy_sphere(3, noise=0.1, seed=77)
    Q = mg.box_queries(P, 3000, seed=78)
    S, D = mg.random_rays(P, 3000, seed=79)
    for leaf in (1, 4, 8):
        tree = o.mesh_lbvh_build(P, I, leaf)
        rm = o.RefMesh.from_tree(P, I, tree)
        assert_results_equal(o.query_point(P, I, tree, Q, 1e6), rm.query_point(Q, 1e6), POINT_FIELDS)
        assert_results_equal(o.query_point(P, I, tree, Q, 0.05), rm.query_point(Q, 0.05), POINT_FIELDS)
        assert_results_equal(o.query_ray(P, I, tree, S, D, 1e6), rm.query_ray(S, D, 1e6), RAY_FIELDS)
    # reference refit of a reference-built SAH tree produces exact unions of the moved item boxes
    rm = o.RefMesh(P, I, o.SAH, 4)
    P2 = (P * np.float32(1.3)).astype(np.float32)
    rm.points[:] = P2
    rm.refit()
    lo2, hi2 = o.triangle_bounds(P2, I)
    t2 = rm.tree()
    vis = visible_nodes(t2)
    for c in vis[:: max(1, len(vis) // 200)]:
        if t2["node_lowers"]["ib"][c] >> 31:
            s, e = int(t2["node_lowers"]["ib"][c] & 0x7FFFFFFF), int(t2["node_uppers"]["ib"][c] & 0x7FFFFFFF)
            items = t2["primitive_indices"][s:e]
            assert np.array_equal([t2["node_lowers"][f][c] for f in "xyz"], lo2[items].min(axis=0))


# ------------------------------------------------------------------------------------------------
# the reference's own CUDA LBVH (bvh.cu run on a B200; fixtures made by baseline/ref_cuda.py golden)
# ------------------------------------------------------------------------------------------------
REF_GOLD = os.path.join(os.path.dirname(__file__), "golden", "golden_ref_lbvh.npz")
REF_CASES = ("cube", "ico2", "ico4", "height33", "cloth40", "dups")


@pytest.fixture(scope="module")
def ref_gold():
    return np.load(REF_GOLD)


@pytest.mark.parametrize("name", REF_CASES)
@pytest.mark.parametrize("leaf", [1, 4])
def test_lbvh_restatement_equals_reference_cuda_lbvh(oracle_mod, ref_gold, name, leaf):
    """Sorted primitive order, parents, root and both half-node arrays (all 2N-1 nodes, muted ones
    included) of the oracle are bit-identical to what the reference's bvh.cu produced on the GPU."""
    g = ref_gold
    t = oracle_mod.mesh_lbvh_build(g[f"{name}_points"], g[f"{name}_indices"], leaf)
    assert t["root"] == int(g[f"{name}_leaf{leaf}_root"])
    assert np.array_equal(t["primitive_indices"], g[f"{name}_leaf{leaf}_primitive_indices"])
    assert np.array_equal(t["parents"], g[f"{name}_leaf{leaf}_parents"])
    for k in ("node_lowers", "node_uppers"):
        for f in ("x", "y", "z", "ib"):
            assert np.array_equal(t[k][f], g[f"{name}_leaf{leaf}_{k}"][f]), (k, f)


@pytest.mark.parametrize("name", ["cube", "ico2", "height33"])
def test_queries_vs_reference_cuda_kernels(oracle_mod, ref_gold, name):
    """Against the reference's NVRTC-compiled kernels (--fmad=true) the contract is the tolerance one:
    result flags equal; faces equal except where both answers are the same geometric point (a shared
    edge / vertex: an exact tie that FMA rounding breaks differently) -- counted; u, v, t within 1e-5."""
    g = ref_gold
    P, I = g[f"{name}_points"], g[f"{name}_indices"]
    tri = I.reshape(-1, 3)
    Q, S, D = g[f"{name}_queries"], g[f"{name}_ray_starts"], g[f"{name}_ray_dirs"]

    def pos(face, u, v):
        T = P[tri[face]]
        return u[:, None] * T[:, 0] + v[:, None] * T[:, 1] + (1 - u - v)[:, None] * T[:, 2]

    for leaf in (1, 4):
        t = oracle_mod.mesh_lbvh_build(P, I, leaf)
        a = oracle_mod.query_point(P, I, t, Q, 1e6)
        r = {k: g[f"{name}_leaf{leaf}_point_{k}"] for k in POINT_FIELDS}
        assert np.array_equal(a["result"], r["result"]) and np.array_equal(a["sign"], r["sign"])
        same = a["face"] == r["face"]
        assert np.allclose(a["u"][same], r["u"][same], rtol=0, atol=1e-5)
        assert np.allclose(a["v"][same], r["v"][same], rtol=0, atol=1e-5)
        ties = int((~same).sum())
        # measured on these fixtures (256 queries each): cube 0 / 0, ico2 19 / 19, height33 15 / 21 for leaf 1 / 4 -- the
        # counts are a property of the fixtures and of the two builds' rounding, so they are pinned, with a margin of 2
        measured = {("cube", 1): 0, ("cube", 4): 0, ("ico2", 1): 19, ("ico2", 4): 19, ("height33", 1): 15, ("height33", 4): 21}
        assert ties <= measured[(name, leaf)] + 2, (name, leaf, ties)
        pa, pr = pos(a["face"], a["u"], a["v"]), pos(r["face"], r["u"], r["v"])
        assert np.abs(pa - pr).max() < 1e-5, "differing faces must still be the same closest point"
        b = oracle_mod.query_ray(P, I, t, S, D, 1e6)
        rr = {k: g[f"{name}_leaf{leaf}_ray_{k}"] for k in RAY_FIELDS}
        assert np.array_equal(b["result"], rr["result"]) and np.array_equal(b["face"], rr["face"])
        assert np.allclose(b["t"], rr["t"], rtol=1e-5, atol=0), "t within 1e-5 relative"
        for k in ("u", "v"):  # barycentrics live in [0, 1]: 1e-5 of full scale
            assert np.allclose(b[k], rr[k], rtol=0, atol=1e-5), k
        assert np.allclose(b["normal"], rr["normal"], atol=1e-6)
        assert np.array_equal(np.sign(b["sign"]), np.sign(rr["sign"]))


def _brute_aabb(lo, hi, qlo, qhi):
    return [np.flatnonzero(~((qlo[i] > hi).any(1) | (qhi[i] < lo).any(1))) for i in range(len(qlo))]


def _brute_ray(lo, hi, s, d, max_dist):
    out = []
    for i in range(len(s)):
        with np.errstate(divide="ignore", invalid="ignore"):
            rcp = np.float32(1.0) / d[i]
            l1, l2 = (lo - s[i]) * rcp, (hi - s[i]) * rcp
        lmin, lmax = np.minimum(l1, l2).max(1), np.maximum(l1, l2).min(1)
        out.append(np.flatnonzero((lmax >= 0) & (lmax >= lmin) & ~(lmin >= max_dist)))
    return out


@pytest.mark.parametrize("leaf", [1, 2, 4])
def test_generic_bvh_query_restatement_vs_brute_force(oracle_mod, leaf):
    """warp/tests/geometry/test_bvh.py:186-262 criterion: exact overlap-set equality with numpy brute force."""
    lo, hi = random_boxes(100, seed=123)
    tree = oracle_mod.lbvh_build(lo, hi, leaf)
    rng = np.random.default_rng(5)
    qlo = (rng.random((64, 3)) * 10).astype(np.float32)
    qhi = (qlo + rng.random((64, 3)).astype(np.float32) * 3).astype(np.float32)
    off, idx = oracle_mod.bvh_query(tree, lo, hi, qlo, qhi)
    for i, want in enumerate(_brute_aabb(lo, hi, qlo, qhi)):
        assert sorted(idx[off[i] : off[i + 1]].tolist()) == want.tolist()
    s = (rng.random((64, 3)) * 10).astype(np.float32)
    d = rng.standard_normal((64, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    for md in (3.4e38, 4.0):
        off, idx = oracle_mod.bvh_query(tree, lo, hi, s, d, ray=True, max_dist=md)
        for i, want in enumerate(_brute_ray(lo, hi, s, d, np.float32(md))):
            assert sorted(idx[off[i] : off[i + 1]].tolist()) == want.tolist()


# ------------------------------------------------------------------------------------------------
# any-hit / intersection count / eval_position
# ------------------------------------------------------------------------------------------------
RAYV_GOLD = os.path.join(os.path.dirname(__file__), "golden", "golden_ray_variants.npz")


def test_ray_variants_match_reference_fixture(oracle_mod, gold):
    o = oracle_mod
    rv = np.load(RAYV_GOLD)
    P, I = gold["mesh_points"], gold["mesh_indices"]
    S, D = rv["starts"], rv["dirs"]
    for name in ("sah", "lbvh1", "lbvh4"):
        tree = {k: gold[f"{name}_tree_{k}"] for k in ("node_lowers", "node_uppers", "primitive_indices")}
        tree["root"] = int(gold[f"{name}_tree_root"])
        for mt in (0.5, 1.0, 1e6):
            assert np.array_equal(o.query_ray_anyhit(P, I, tree, S, D, mt), rv[f"{name}_anyhit_{mt:g}"])
        cnt = o.query_ray_count(P, I, tree, S, D)
        assert np.array_equal(cnt, rv[f"{name}_count"])
        assert np.array_equal(o.mesh_eval(P, I, rv["eval_face"], rv["eval_u"], rv["eval_v"]), rv[f"{name}_eval_position"])
    # closed surface: rays starting inside (even rows) cross it an odd number of times, and the counts do
    # not depend on the tree
    assert np.all(rv["sah_count"][::2][np.linalg.norm(S[::2], axis=1) < 0.8] % 2 == 1)
    assert np.array_equal(rv["sah_count"], rv["lbvh1_count"])
    assert rv["sah_anyhit_1e+06"].sum() > 256 and rv["sah_anyhit_0.5"].sum() < rv["sah_anyhit_1e+06"].sum()


def test_ray_variants_live_reference(oracle_mod):
    o = oracle_mod
    if not o.ref_available():
        pytest.skip("oracle/_ref/libwarp_ref_cpu.so not present")
    P, I = mg.noisy_sphere(3, noise=0.1, seed=5)
    S, D = mg.random_rays(P, 4000, seed=6)
    S[::3] *= np.float32(0.2)
    D[:30] = np.eye(3, dtype=np.float32)[np.arange(30) % 3]
    for leaf in (1, 4):
        tree = o.mesh_lbvh_build(P, I, leaf)
        rm = o.RefMesh.from_tree(P, I, tree)
        for mt in (0.7, 1e6):
            a = o.query_ray_anyhit(P, I, tree, S, D, mt)
            assert np.array_equal(a, rm.query_ray_anyhit(S, D, mt))
            assert np.array_equal(a, o.query_ray(P, I, tree, S, D, mt)["result"])
        assert np.array_equal(o.query_ray_count(P, I, tree, S, D), rm.query_ray_count(S, D))


def test_mesh_query_aabb_restatement_matches_reference_fixture(oracle_mod, gold):
    """mesh_query_aabb + mesh_query_aabb_next (mesh.h:2476-2712) == the generic iterator restatement over
    per-triangle boxes: hit lists equal IN ORDER on the reference's SAH tree and on LBVH trees."""
    o = oracle_mod
    rv = np.load(RAYV_GOLD)
    P, I = gold["mesh_points"], gold["mesh_indices"]
    tlo, thi = o.triangle_bounds(P, I)
    for name in ("sah", "lbvh1", "lbvh4"):
        tree = {k: gold[f"{name}_tree_{k}"] for k in ("node_lowers", "node_uppers", "primitive_indices")}
        tree["root"] = int(gold[f"{name}_tree_root"])
        off, idx = o.bvh_query(tree, tlo, thi, rv["aabb_lowers"], rv["aabb_uppers"])
        assert np.array_equal(off, rv[f"{name}_aabb_offsets"]) and np.array_equal(idx, rv[f"{name}_aabb_indices"])
        assert off[-1] > 500
    for i, want in enumerate(_brute_aabb(tlo, thi, rv["aabb_lowers"], rv["aabb_uppers"])):
        o_, x_ = rv["lbvh4_aabb_offsets"], rv["lbvh4_aabb_indices"]
        assert sorted(x_[o_[i] : o_[i + 1]].tolist()) == want.tolist()


# ------------------------------------------------------------------------------------------------
# generic iterator + group roots pinned on the reference's own header code (oracle/_ref) and on a fixture
# ------------------------------------------------------------------------------------------------
GROUP_GOLD = os.path.join(os.path.dirname(__file__), "golden", "golden_group_queries.npz")


def _group_case(seed, n, ngroups):
    rng = np.random.default_rng(seed)
    lo, hi = random_boxes(n, seed=seed)
    groups = rng.integers(0, ngroups, n).astype(np.int32)
    qlo = (rng.random((64, 3)) * 10).astype(np.float32)
    qhi = (qlo + rng.random((64, 3)).astype(np.float32) * 4).astype(np.float32)
    s = (rng.random((64, 3)) * 10).astype(np.float32)
    d = rng.standard_normal((64, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    gid = rng.integers(-1, ngroups + 2, 64).astype(np.int32)  # some absent groups
    return lo, hi, groups, qlo, qhi, s, d, gid


def test_group_roots_and_rooted_queries_match_reference_fixture(oracle_mod):
    o = oracle_mod
    g = np.load(GROUP_GOLD)
    for leaf in (1, 4):
        lo, hi, groups, qlo, qhi, s, d, gid = _group_case(900 + leaf, 400, 6)
        tree = o.lbvh_build(lo, hi, leaf, groups=groups)
        roots = o.bvh_group_roots(tree, groups, gid)
        assert np.array_equal(roots, g[f"leaf{leaf}_roots"])
        assert (roots == -1).sum() > 0 and (roots >= 0).sum() > 30
        off, idx = o.bvh_query(tree, lo, hi, qlo, qhi, roots=roots)
        assert np.array_equal(off, g[f"leaf{leaf}_aabb_offsets"]) and np.array_equal(idx, g[f"leaf{leaf}_aabb_indices"])
        off, idx = o.bvh_query(tree, lo, hi, s, d, ray=True, max_dist=6.0, roots=roots)
        assert np.array_equal(off, g[f"leaf{leaf}_ray_offsets"]) and np.array_equal(idx, g[f"leaf{leaf}_ray_indices"])
        # a rooted query reports exactly the whole-tree hits that belong to the group (absent group -> whole tree)
        woff, widx = o.bvh_query(tree, lo, hi, qlo, qhi)
        off, idx = o.bvh_query(tree, lo, hi, qlo, qhi, roots=roots)
        for i in range(64):
            whole = widx[woff[i] : woff[i + 1]]
            want = whole if roots[i] == -1 else whole[groups[whole] == gid[i]]
            assert sorted(idx[off[i] : off[i + 1]].tolist()) == sorted(want.tolist())


def test_generic_iterator_live_reference(oracle_mod):
    o = oracle_mod
    if not o.ref_available():
        pytest.skip("oracle/_ref/libwarp_ref_cpu.so not present")
    for seed, n, ng, leaf in ((1, 300, 5, 1), (2, 1000, 9, 4), (3, 64, 64, 2), (4, 500, 1, 8)):
        lo, hi, groups, qlo, qhi, s, d, gid = _group_case(seed, n, ng)
        for grp in (groups, None):
            tree = o.lbvh_build(lo, hi, leaf, groups=grp)
            roots = o.bvh_group_roots(tree, grp, gid)
            assert np.array_equal(roots, o.ref_bvh_group_roots(tree, grp, gid))
            for r in (None, roots):
                a, b = o.bvh_query(tree, lo, hi, qlo, qhi, roots=r), o.ref_bvh_query(tree, lo, hi, qlo, qhi, roots=r)
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
                a = o.bvh_query(tree, lo, hi, s, d, ray=True, max_dist=5.0, roots=r)
                b = o.ref_bvh_query(tree, lo, hi, s, d, ray=True, max_dist=5.0, roots=r)
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_sign_parity_restatement(oracle_mod, gold):
    """mesh_query_point_sign_parity: fixture from the reference C++ (g++ build: offsets drawn right to left), and the
    left-to-right order of the reference's device builds gives the same signs on a closed mesh."""
    o = oracle_mod
    rv = np.load(RAYV_GOLD)
    P, I, Q = gold["mesh_points"], gold["mesh_indices"], gold["queries"]
    for name in ("sah", "lbvh1", "lbvh4"):
        tree = {k: gold[f"{name}_tree_{k}"] for k in ("node_lowers", "node_uppers", "primitive_indices")}
        tree["root"] = int(gold[f"{name}_tree_root"])
        for ns in (1, 3):
            want = {k: rv[f"{name}_parity{ns}_{k}"] for k in POINT_FIELDS}
            assert_results_equal(o.query_point_sign_parity(P, I, tree, Q, 1e6, ns, 0.1, rtl=True), want, POINT_FIELDS)
            ltr = o.query_point_sign_parity(P, I, tree, Q, 1e6, ns, 0.1, rtl=False)
            assert np.array_equal(ltr["sign"], want["sign"]) and np.array_equal(ltr["face"], want["face"])
        # the parity sign agrees with the three-axis-probe sign of mesh_query_point on this closed mesh
        assert np.array_equal(rv[f"{name}_parity1_sign"], gold[f"{name}_point_sign"] if f"{name}_point_sign" in gold else rv[f"{name}_parity1_sign"])
    assert (rv["lbvh4_parity1_sign"] < 0).sum() > 20 and (rv["lbvh4_parity1_sign"] > 0).sum() > 20


def test_sign_parity_live_reference(oracle_mod):
    o = oracle_mod
    if not o.ref_available():
        pytest.skip("oracle/_ref/libwarp_ref_cpu.so not present")
    P, I = mg.noisy_sphere(3, noise=0.1, seed=15)
    I_open = I.reshape(-1, 3)[::2].reshape(-1).copy()  # every other face removed: parity depends on the direction
    Q = mg.box_queries(P, 2000, seed=16)
    for idx in (I, I_open):
        tree = o.mesh_lbvh_build(P, idx, 4)
        rm = o.RefMesh.from_tree(P, idx, tree)
        for ns, sc in ((1, 0.1), (4, 0.5), (0, 0.1)):
            assert_results_equal(o.query_point_sign_parity(P, idx, tree, Q, 1e6, ns, sc, rtl=True),
                                 rm.query_point_sign_parity(Q, 1e6, ns, sc), POINT_FIELDS)
        assert_results_equal(o.query_point_sign_parity(P, idx, tree, Q, 0.05, 1, 0.1, rtl=True),
                             rm.query_point_sign_parity(Q, 0.05, 1, 0.1), POINT_FIELDS)


SIGNN_GOLD = os.path.join(os.path.dirname(__file__), "golden", "golden_sign_normal.npz")


def test_sign_normal_restatement(oracle_mod, gold):
    """mesh_query_point_sign_normal + Mesh.average_edge_length: fixture produced by the reference C++
    (tests/golden/make_golden_sign_normal.py), SAH and LBVH trees, welding bands 1e-3 / 1e-1, near / far max_dist."""
    o = oracle_mod
    sn = np.load(SIGNN_GOLD)
    P, I, Q = gold["mesh_points"], gold["mesh_indices"], sn["queries"]
    avg = float(sn["average_edge_length"])
    assert np.float32(o.average_edge_length(P, I, mode=0)) == np.float32(avg)  # the reference's float loop, bit for bit
    assert abs(o.average_edge_length(P, I, mode=1) - avg) <= 1e-6 * avg       # double accumulation (the CUDA path's)
    for name in ("sah", "lbvh1", "lbvh4"):
        tree = {k: gold[f"{name}_tree_{k}"] for k in ("node_lowers", "node_uppers", "primitive_indices")}
        tree["root"] = int(gold[f"{name}_tree_root"])
        for tag, eps, md in (("e3", 1e-3, 1e6), ("e1", 1e-1, 1e6), ("e3near", 1e-3, 0.05)):
            want = {k: sn[f"{name}_{tag}_{k}"] for k in POINT_FIELDS}
            assert_results_equal(o.query_point_sign_normal(P, I, tree, Q, md, avg, eps), want, POINT_FIELDS)
    s = sn["lbvh4_e3_sign"]
    assert (s < 0).sum() > 50 and (s > 0).sum() > 50
    # closed mesh: away from the welding band the normal sign is the ray-vote sign of mesh_query_point
    box = o.query_point(P, I, tree, Q[:600], 1e6)
    assert (box["sign"] == sn["lbvh4_e3_sign"][:600]).mean() > 0.99


def test_sign_normal_live_reference(oracle_mod):
    o = oracle_mod
    if not o.ref_available():
        pytest.skip("oracle/_ref/libwarp_ref_cpu.so not present")
    P, I = mg.noisy_sphere(3, noise=0.1, seed=17)
    rng = np.random.default_rng(18)
    T = I.reshape(-1, 3)
    tri = T[rng.integers(0, len(T), 500)]
    Q = np.concatenate([mg.box_queries(P, 2000, seed=19), P[rng.integers(0, len(P), 500)] + rng.normal(0, 1e-4, (500, 3)),
                        0.5 * (P[tri[:, 0]] + P[tri[:, 2]]), P[:300]]).astype(np.float32)
    rm = o.RefMesh(P, I)
    avg = rm.average_edge_length
    assert np.float32(avg) == np.float32(o.average_edge_length(P, I, mode=0))
    tree = o.mesh_lbvh_build(P, I, 4)
    rl = o.RefMesh.from_tree(P, I, tree)
    rl.average_edge_length = avg
    for eps, md in ((1e-3, 1e6), (0.2, 1e6), (0.0, 1e6), (1e-3, 0.05)):
        assert_results_equal(o.query_point_sign_normal(P, I, tree, Q, md, avg, eps), rl.query_point_sign_normal(Q, md, eps),
                             POINT_FIELDS)


FAR_FIELDS = ("result", "face", "u", "v")


def test_furthest_point_and_face_normal_restatement(oracle_mod, gold):
    """mesh_query_furthest_point_no_sign / mesh_eval_face_normal: fixture produced by the reference C++
    (tests/golden/make_golden_sign_normal.py), SAH and LBVH trees; then a brute-force check and the live reference."""
    o = oracle_mod
    sn = np.load(SIGNN_GOLD)
    P, I, Q = gold["mesh_points"], gold["mesh_indices"], sn["queries"]
    for name in ("sah", "lbvh1", "lbvh4"):
        tree = {k: gold[f"{name}_tree_{k}"] for k in ("node_lowers", "node_uppers", "primitive_indices")}
        tree["root"] = int(gold[f"{name}_tree_root"])
        for tag, md in (("far0", 0.0), ("far2", 2.0)):
            want = {k: sn[f"{name}_{tag}_{k}"] for k in FAR_FIELDS}
            assert_results_equal(o.query_furthest_point_no_sign(P, I, tree, Q, md), want, FAR_FIELDS)
    assert np.array_equal(o.mesh_face_normal(P, I, sn["normal_faces"]), sn["face_normals"])
    # brute force: the farthest point of a mesh is its farthest vertex
    got = o.query_furthest_point_no_sign(P, I, tree, Q[:200], 0.0)
    T = I.reshape(-1, 3)
    bary = np.stack([got["u"], got["v"], 1 - got["u"] - got["v"]], axis=1)
    pos = (P[T[got["face"]]] * bary[:, :, None]).sum(axis=1)
    d_got = np.linalg.norm(pos - Q[:200], axis=1)
    d_max = np.sqrt(((P[None, :, :] - Q[:200, None, :]) ** 2).sum(-1)).max(axis=1)
    assert np.allclose(d_got, d_max, rtol=1e-6)
    if o.ref_available():
        P2, I2 = mg.noisy_sphere(3, noise=0.1, seed=23)
        Q2 = mg.box_queries(P2, 3000, seed=24)
        tree2 = o.mesh_lbvh_build(P2, I2, 4)
        rl = o.RefMesh.from_tree(P2, I2, tree2)
        for md in (0.0, 1.5, 1e6):
            assert_results_equal(o.query_furthest_point_no_sign(P2, I2, tree2, Q2, md), rl.query_furthest_point_no_sign(Q2, md),
                                 FAR_FIELDS)
        f = np.arange(len(I2) // 3, dtype=np.int32)
        assert np.array_equal(o.mesh_face_normal(P2, I2, f), rl.eval_face_normal(f))


def _brute_sphere(lo, hi, centers, radii):
    r = np.maximum(radii, 0).astype(np.float32)
    out = []
    for c, rr in zip(centers, r):
        d = np.maximum(np.maximum(lo - c, c - hi), np.float32(0))
        out.append(np.flatnonzero((d * d).sum(axis=1, dtype=np.float32) <= rr * rr))
    return out


def test_bvh_sphere_capsule_restatement(oracle_mod):
    """Sphere / capsule kinds of the generic iterator: fixture from the reference C++
    (tests/golden/make_golden_bvh_kinds.py), exact lists in iterator order; sphere sets against brute force; a
    zero-radius capsule closed at max_dist contains the plain ray's half-open list; then the live reference."""
    o = oracle_mod
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_bvh_kinds.npz"))
    lo, hi = random_boxes(2000, seed=int(g["seed_boxes"]))
    C, R, D = g["centers"], g["radii"], g["dirs"]
    for leaf in (1, 4):
        tree = o.lbvh_build(lo, hi, leaf)
        off, idx = o.bvh_query_kind(tree, lo, hi, "sphere", C, radii=R)
        assert np.array_equal(off, g[f"leaf{leaf}_sphere_offsets"]) and np.array_equal(idx, g[f"leaf{leaf}_sphere_indices"])
        if leaf == 4:  # (single-item leaves are reported on the node test alone, which is the same test here)
            for i, want in enumerate(_brute_sphere(lo, hi, C, R)):
                assert sorted(idx[off[i] : off[i + 1]].tolist()) == want.tolist()
        for tag, md in (("inf", 3.4028234663852886e38), ("3", 3.0)):
            off, idx = o.bvh_query_kind(tree, lo, hi, "capsule", C, D, radii=R, max_dist=md)
            assert np.array_equal(off, g[f"leaf{leaf}_capsule{tag}_offsets"])
            assert np.array_equal(idx, g[f"leaf{leaf}_capsule{tag}_indices"])
        roff, ridx = o.bvh_query(tree, lo, hi, C, D, ray=True, max_dist=3.0)
        coff, cidx = o.bvh_query_kind(tree, lo, hi, "capsule", C, D, radii=0.0, max_dist=3.0)
        for i in range(len(C)):
            assert set(ridx[roff[i] : roff[i + 1]].tolist()) <= set(cidx[coff[i] : coff[i + 1]].tolist())
    if o.ref_available():
        lo2, hi2 = random_boxes(5000, seed=43)
        tree = o.lbvh_build(lo2, hi2, 2)
        roots = np.full(len(C), -1, np.int32)
        roots[::2] = tree["root"]
        for kind, kw in (("sphere", dict(qa=C, radii=R)), ("capsule", dict(qa=C, qb=D, radii=0.3, max_dist=2.0)),
                         ("capsule", dict(qa=C, qb=D, radii=R, roots=roots))):
            a, b = o.bvh_query_kind(tree, lo2, hi2, kind, **kw), o.ref_bvh_query_kind(tree, lo2, hi2, kind, **kw)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_mesh_query_sphere_restatement(oracle_mod, gold):
    """mesh_query_sphere: fixture from the reference C++ (tests/golden/make_golden_bvh_kinds.py; the mesh carries three
    zero-area faces), exact lists in iterator order, and the sets against a numpy closest-point check."""
    o = oracle_mod
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_bvh_kinds.npz"))
    P, I, C, R = gold["mesh_points"], g["mesh_indices"], g["mesh_centers"], g["mesh_radii"]
    for leaf in (1, 4):
        tree = o.mesh_lbvh_build(P, I, leaf)
        off, idx = o.mesh_query_sphere(P, I, tree, C, R)
        assert np.array_equal(off, g[f"mesh_leaf{leaf}_sphere_offsets"]) and np.array_equal(idx, g[f"mesh_leaf{leaf}_sphere_indices"])
    # same faces as "closest point of the face within the radius" evaluated face by face (non-degenerate faces)
    T = I.reshape(-1, 3)[:-3]
    tree = o.mesh_lbvh_build(P, T.reshape(-1), 4)
    off, idx = o.mesh_query_sphere(P, T.reshape(-1), tree, C[:60], R[:60])
    for i in range(60):
        got = sorted(idx[off[i] : off[i + 1]].tolist())
        rr = max(float(R[i]), 0.0)
        want = []
        for f, (a, b, c) in enumerate(T):
            uv = o.closest_point_to_triangle(P[a], P[b], P[c], C[i])
            cp = P[a] * uv[0] + P[b] * uv[1] + P[c] * (np.float32(1) - uv[0] - uv[1])
            d = cp - C[i]
            if np.float32(d @ d) <= np.float32(rr * rr) * np.float32(1 + 1e-5):
                want.append(f)
        assert set(got) <= set(want) and len(want) - len(got) <= 1  # (one borderline face at most)


def _topology_from_key_deltas(keys, prim, width):
    """Parents of every node of the reference LBVH WITHOUT walking it bottom-up: the tree is the Cartesian tree of the
    key-delta array (delta[i] = common-prefix length of keys i and i+1, bvh.cu:218-226).  The node split after sorted
    position s covers (L, R] with L, R the nearest splits on either side whose delta is SMALLER -- two independent
    searches per node; two equal deltas below the key width never compete, because sorted keys put a smaller delta
    between them.  Only runs of EQUAL keys (delta == width) depend on the parity tie-break of bvh.cu:325-329 and are
    replayed sequentially, run by run.  Groundwork for a parallel hierarchy kernel."""
    n = len(keys)
    k = [int(v) for v in keys]
    delta = [width if k[i] == k[i + 1] else width - (k[i] ^ k[i + 1]).bit_length() for i in range(n - 1)]
    par = [int(v) % 2 for v in prim]

    def goes_right(l, r):  # bvh.cu:300-334, ungrouped
        if l == 0:
            return True
        if r == n - 1:
            return False
        if delta[r] != delta[l - 1]:
            return delta[r] > delta[l - 1]
        return (par[l - 1] ^ par[r]) != 0

    rng_l, rng_r = [0] * (n - 1), [0] * (n - 1)
    # nearest smaller delta to the left / right (plain stack sweeps here; a kernel would search per node)
    stack = []
    for s in range(n - 1):
        while stack and delta[stack[-1]] >= delta[s]:
            stack.pop()
        rng_l[s] = stack[-1] + 1 if stack else 0
        stack.append(s)
    stack = []
    for s in range(n - 2, -1, -1):
        while stack and delta[stack[-1]] >= delta[s]:
            stack.pop()
        rng_r[s] = stack[-1] if stack else n - 1
        stack.append(s)
    # runs of equal keys: replay the bottom-up process inside the run [p, q]
    s = 0
    while s < n - 1:
        if delta[s] != width:
            s += 1
            continue
        p = s
        while s < n - 1 and delta[s] == width:
            s += 1
        q = s  # keys p..q are equal, splits p..q-1
        parked = []  # nodes waiting for their right sibling: (l, r), each the left child of split r
        for i in range(p, q + 1):
            l, r = i, i
            while not (l == p and r == q):
                if goes_right(l, r):
                    parked.append((l, r))
                    break
                pl, pr = parked.pop()  # left sibling: ends at l - 1
                assert pr == l - 1
                rng_l[pr], rng_r[pr] = pl, r
                l = pl
        assert not parked
    parents = [-1] * (2 * n - 1)
    for i in range(n):
        parents[i] = n + (i if goes_right(i, i) else i - 1)
    for s in range(n - 1):
        l, r = rng_l[s], rng_r[s]
        if l == 0 and r == n - 1:
            root = n + s
        else:
            parents[n + s] = n + (r if goes_right(l, r) else l - 1)
    return np.asarray(parents, np.int32), root


def test_topology_is_the_cartesian_tree_of_key_deltas(oracle_mod):
    o = oracle_mod
    rng = np.random.default_rng(77)
    cases = []
    P, I = mg.noisy_sphere(4, 0.05, 71)
    lo, hi = o.triangle_bounds(P, I)
    cases.append((lo, hi, 30))
    cases.append((lo, hi, 63))
    c = (rng.random((3000, 3)) * 4).astype(np.float32)  # clustered boxes: many equal 30-bit keys
    c = np.repeat(c[:300], 10, axis=0) + (rng.random((3000, 3)) * 1e-4).astype(np.float32)
    cases.append((c, c + np.float32(0.01), 30))
    same = np.tile(np.array([[1.0, 2.0, 3.0]], np.float32), (257, 1))  # one run of 257 equal keys
    cases.append((np.concatenate([same, c[:100]]), np.concatenate([same + 1, c[:100] + 1]), 30))
    for lo, hi, bits in cases:
        tree = o.lbvh_build(lo, hi, 1, morton_bits=bits)
        width = 32 if bits == 30 else 64
        parents, root = _topology_from_key_deltas(tree["keys"], tree["primitive_indices"], width)
        assert root == tree["root"]
        want = tree["parents"].copy()
        want[root] = -1
        assert np.array_equal(parents, want)
    # grouped trees (key = group << 32 | code, bvh.cu:205-209): the stay-inside-the-group rule of bvh.cu:305-321 never
    # overrides the delta comparison -- the same-group neighbour always shares the longer prefix -- so the grouped
    # tree is the Cartesian tree of the 64-bit deltas too
    lo, hi, _ = cases[2]
    for groups in (rng.integers(0, 7, len(lo)).astype(np.int32), (np.arange(len(lo)) // 450).astype(np.int32)):
        tree = o.lbvh_build(lo, hi, 1, groups=groups)
        parents, root = _topology_from_key_deltas(tree["keys"], tree["primitive_indices"], 64)
        want = tree["parents"].copy()
        want[tree["root"]] = -1
        assert root == tree["root"] and np.array_equal(parents, want)


def _grouped_mesh_case():
    P, I = mg.noisy_sphere(3, noise=0.05, seed=61)
    T = len(I) // 3
    cz = P[I.reshape(-1, 3)].mean(axis=1)[:, 2]
    groups = np.clip(((cz + 1.1) / 2.2 * 5).astype(np.int32), 0, 4)  # five z bands
    S, D = mg.random_rays(P, 1500, seed=62)
    S[::3] *= np.float32(0.2)
    gid = np.random.default_rng(63).integers(-1, 6, 1500).astype(np.int32)
    return P, I, T, groups, S, D, gid


def test_rooted_mesh_rays_live_reference(oracle_mod):
    """mesh_query_ray / _anyhit / _count_intersections with a `root` argument (a group's subtree) against the
    reference C++; a rooted ray sees exactly the faces of its group."""
    o = oracle_mod
    if not o.ref_available():
        pytest.skip("oracle/_ref/libwarp_ref_cpu.so not present")
    P, I, T, groups, S, D, gid = _grouped_mesh_case()
    for leaf in (1, 4):
        tree = o.mesh_lbvh_build(P, I, leaf, groups=groups)
        roots = o.bvh_group_roots(tree, groups, gid)
        assert np.array_equal(roots, o.ref_bvh_group_roots(tree, groups, gid))
        rm = o.RefMesh.from_tree(P, I, tree)
        got = o.query_ray(P, I, tree, S, D, 1e6, roots=roots)
        assert_results_equal(got, rm.query_ray(S, D, 1e6, roots=roots), RAY_FIELDS)
        assert np.array_equal(o.query_ray_anyhit(P, I, tree, S, D, 0.9, roots=roots), rm.query_ray_anyhit(S, D, 0.9, roots=roots))
        assert np.array_equal(o.query_ray_count(P, I, tree, S, D, roots=roots), rm.query_ray_count(S, D, roots=roots))
        hit = (got["result"] == 1) & (roots != -1)
        assert hit.sum() > 100 and np.array_equal(groups[got["face"][hit]], gid[hit])
