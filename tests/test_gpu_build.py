"""GPU parity tests of LBVH build / refit / rebuild: the CUDA path (through the C ABI) against the
oracle restatement of warp/native/bvh.cu on the same inputs.  Bit-exact: Morton keys, sorted
primitive order, hierarchy topology (children, parents, root, packed-leaf ranges) and node boxes."""
import os

import numpy as np
import pytest

from helpers import assert_tree_equal, random_boxes, visible_nodes
from warp_b200 import meshgen as mg

pytestmark = pytest.mark.gpu


def gpu_mesh(wp, P, I, leaf=None, **kw):
    pts = wp.array(P, dtype=wp.vec3, device="cuda:0")
    idx = wp.array(I, dtype=wp.int32, device="cuda:0")
    return wp.Mesh(pts, idx, bvh_constructor="lbvh", bvh_leaf_size=leaf, **kw)


MESHES = {
    "cube": lambda: (mg.CUBE_POINTS, mg.CUBE_INDICES_RH),
    "ico2_noisy": lambda: mg.noisy_sphere(2, 0.05, 11),
    "ico5_noisy": lambda: mg.noisy_sphere(5, 0.02, 1),
    "height65": lambda: mg.heightfield(65),
    "cloth130": lambda: mg.cloth(130, frame=3),
}


@pytest.mark.parametrize("name", list(MESHES))
@pytest.mark.parametrize("leaf", [1, 2, 4, 8])
def test_mesh_build_bit_exact(wp, oracle_mod, name, leaf):
    P, I = MESHES[name]()
    m = gpu_mesh(wp, P, I, leaf)
    got = m.download_tree()
    want = oracle_mod.mesh_lbvh_build(P, I, leaf)
    assert_tree_equal(got, want)
    for k in ("total_lower", "total_upper", "inv_edges"):
        assert np.array_equal(got[k], want[k]), k


def test_default_leaf_size_is_4(wp, oracle_mod):
    P, I = mg.noisy_sphere(3)
    m = gpu_mesh(wp, P, I)  # bvh_leaf_size=None -> 4 (types.py:6181-6182)
    assert m.bvh_leaf_size == 4
    assert_tree_equal(m.download_tree(), oracle_mod.mesh_lbvh_build(P, I, 4))


@pytest.mark.parametrize("n", [1, 2, 3, 5, 33, 255, 256, 257, 4095, 4096, 4097, 10000, 70001])
def test_bvh_build_sizes(wp, oracle_mod, n):
    """Edge sizes around the sort tile (4096) and block (256) boundaries; wp.Bvh default leaf_size 1."""
    lo, hi = random_boxes(n, seed=n)
    b = wp.Bvh(wp.array(lo, dtype=wp.vec3), wp.array(hi, dtype=wp.vec3))
    assert_tree_equal(b.download_tree(), oracle_mod.lbvh_build(lo, hi, 1))


@pytest.mark.parametrize("n,leaf", [(1, 1), (1, 4), (2, 2), (3, 4), (4, 8)])
def test_root_is_a_packed_leaf(wp, oracle_mod, n, leaf):
    """warp/tests/geometry/test_bvh.py:423-511: single-item trees and leaf_size >= N."""
    lo, hi = random_boxes(n, seed=7)
    lo_d, hi_d = wp.array(lo, dtype=wp.vec3), wp.array(hi, dtype=wp.vec3)
    b = wp.Bvh(lo_d, hi_d, leaf_size=leaf)
    want = oracle_mod.lbvh_build(lo, hi, leaf)
    assert_tree_equal(b.download_tree(), want)
    lo2, hi2 = lo + 3.0, hi + 4.0
    lo_d.assign(lo2), hi_d.assign(hi2)
    b.refit()
    got = b.download_tree()
    oracle_mod.lbvh_refit(want, lo2, hi2)
    r = want["root"]
    for f in "xyz":
        assert got["node_lowers"][f][r] == want["node_lowers"][f][r]
        assert got["node_uppers"][f][r] == want["node_uppers"][f][r]


def test_duplicate_keys_depth_rule(wp, oracle_mod):
    """Thousands of coincident boxes: equal keys, parity tie-breaks, depth >= 32 packed leaves."""
    for n, leaf in ((600, 1), (5000, 4)):
        lo = np.zeros((n, 3), np.float32)
        lo[: n // 2] += 1.0
        lo[n // 3 : n // 2, 1] += 2.0
        hi = lo + 0.5
        b = wp.Bvh(wp.array(lo, dtype=wp.vec3), wp.array(hi, dtype=wp.vec3), leaf_size=leaf)
        got = b.download_tree()
        assert got["deep"] == 1
        assert_tree_equal(got, oracle_mod.lbvh_build(lo, hi, leaf))


def test_sort_cross_check_against_cub(wp, oracle_mod):
    """The hand-written onesweep and the library sort (WARP_B200_SORT=cub) give the same tree."""
    P, I = mg.noisy_sphere(6, 0.02, 3)  # 81 920 triangles = 20 sort tiles
    a = gpu_mesh(wp, P, I, 4).download_tree()
    os.environ["WARP_B200_SORT"] = "cub"
    try:
        b = gpu_mesh(wp, P, I, 4).download_tree()
    finally:
        del os.environ["WARP_B200_SORT"]
    assert_tree_equal(a, b)
    assert_tree_equal(a, oracle_mod.mesh_lbvh_build(P, I, 4))


def test_large_mesh_build_properties(wp, oracle_mod):
    """1.3 M triangles (config C2 size): full diff against the oracle + permutation / order properties."""
    P, I = mg.noisy_sphere(8, 0.02, 1)
    m = gpu_mesh(wp, P, I, 4)
    got = m.download_tree()
    assert np.all(np.diff(got["keys"].astype(np.int64)) >= 0)
    assert np.array_equal(np.sort(got["primitive_indices"]), np.arange(len(I) // 3))
    assert_tree_equal(got, oracle_mod.mesh_lbvh_build(P, I, 4))


def _refit_and_compare(wp, oracle_mod, m, pts_dev, P2, I, want):
    pts_dev.assign(P2)
    m.refit()
    got = m.download_tree()
    lo2, hi2 = oracle_mod.triangle_bounds(P2, I)
    oracle_mod.lbvh_refit(want, lo2, hi2)
    vis = visible_nodes(want)
    assert np.array_equal(got["parents"], want["parents"])
    for name in ("node_lowers", "node_uppers"):
        assert np.array_equal(got[name]["ib"], want[name]["ib"])
        for f in "xyz":
            assert np.array_equal(got[name][f][vis], want[name][f][vis]), (name, f)


@pytest.mark.parametrize("leaf", [1, 4])
def test_mesh_refit_bit_exact_on_visible_nodes(wp, oracle_mod, leaf):
    """Refit after deformation: every node a query can reach has the oracle's box (muted nodes under
    packed leaves keep stale boxes in the reference too, bvh.cu:100-118, and are never read)."""
    P, I = mg.noisy_sphere(5, 0.02, 1)
    pts = wp.array(P, dtype=wp.vec3)
    m = wp.Mesh(pts, wp.array(I, dtype=wp.int32), bvh_leaf_size=leaf)
    want = oracle_mod.mesh_lbvh_build(P, I, leaf)
    _refit_and_compare(wp, oracle_mod, m, pts, mg.renoise_sphere(P, 0.05, 3), I, want)
    _refit_and_compare(wp, oracle_mod, m, pts, mg.renoise_sphere(P, 0.01, 4), I, want)  # second refit: counter parity


def test_points_setter_triggers_refit(wp, oracle_mod):
    """warp/tests/geometry/test_mesh_query_point.py:891-972: assigning mesh.points refits."""
    P, I = mg.noisy_sphere(4)
    m = gpu_mesh(wp, P, I, 4)
    want = oracle_mod.mesh_lbvh_build(P, I, 4)
    P2 = (P + np.array([10, 0, 0], np.float32)).astype(np.float32)
    m.points = wp.array(P2, dtype=wp.vec3)
    got = m.download_tree()
    lo2, hi2 = oracle_mod.triangle_bounds(P2, I)
    oracle_mod.lbvh_refit(want, lo2, hi2)
    r = want["root"]
    assert got["node_lowers"]["x"][r] == want["node_lowers"]["x"][r]
    with pytest.raises(RuntimeError, match="same shape"):
        m.points = wp.array(P2[:-1], dtype=wp.vec3)


def test_bvh_refit_then_rebuild(wp, oracle_mod):
    """warp/tests/geometry/test_bvh.py:186-262 procedure (100 boxes, rng 123) on the build products."""
    lo, hi = random_boxes(100, seed=123)
    lo_d, hi_d = wp.array(lo, dtype=wp.vec3), wp.array(hi, dtype=wp.vec3)
    b = wp.Bvh(lo_d, hi_d, constructor="lbvh", leaf_size=2)
    want = oracle_mod.lbvh_build(lo, hi, 2)
    assert_tree_equal(b.download_tree(), want)
    lo2, hi2 = random_boxes(100, seed=124)
    lo_d.assign(lo2), hi_d.assign(hi2)
    b.refit()
    got = b.download_tree()
    oracle_mod.lbvh_refit(want, lo2, hi2)
    vis = visible_nodes(want)
    for f in "xyz":
        assert np.array_equal(got["node_lowers"][f][vis], want["node_lowers"][f][vis])
        assert np.array_equal(got["node_uppers"][f][vis], want["node_uppers"][f][vis])
    b.rebuild()
    assert_tree_equal(b.download_tree(), oracle_mod.lbvh_build(lo2, hi2, 2))
    b.rebuild()  # twice: buffers and counters are reused
    assert_tree_equal(b.download_tree(), oracle_mod.lbvh_build(lo2, hi2, 2))


def test_native_failure_raises_with_error_string(wp):
    """warp/tests/geometry/test_mesh.py:443-485: id 0 -> RuntimeError carrying wp_get_error_string()."""
    P, I = mg.CUBE_POINTS, mg.CUBE_INDICES_RH
    with pytest.raises(RuntimeError, match="Failed to create mesh: .*constructor"):
        gpu_mesh_cubql = wp.Mesh(wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32), bvh_constructor="cubql")  # noqa: F841
    lo, hi = random_boxes(4)
    with pytest.raises(RuntimeError, match="Failed to create BVH"):
        wp.Bvh(wp.array(lo, dtype=wp.vec3), wp.array(hi, dtype=wp.vec3), constructor="cubql")
    with pytest.raises(RuntimeError, match="Failed to create BVH: .*grouped"):  # host constructors take ungrouped items only
        wp.Bvh(wp.array(lo, dtype=wp.vec3), wp.array(hi, dtype=wp.vec3), constructor="median", groups=wp.array(np.zeros(4, np.int32), dtype=wp.int32))


REF_GOLD = os.path.join(os.path.dirname(__file__), "golden", "golden_ref_lbvh.npz")


@pytest.mark.parametrize("name", ["cube", "ico2", "ico4", "height33", "cloth40", "dups"])
@pytest.mark.parametrize("leaf", [1, 4])
def test_build_equals_reference_cuda_lbvh_dump(wp, name, leaf):
    """Directly against arrays dumped from the reference's own bvh.cu on a B200 (baseline/ref_cuda.py golden)."""
    g = np.load(REF_GOLD)
    m = gpu_mesh(wp, g[f"{name}_points"], g[f"{name}_indices"], leaf)
    t = m.download_tree()
    assert t["root"] == int(g[f"{name}_leaf{leaf}_root"])
    assert np.array_equal(t["primitive_indices"], g[f"{name}_leaf{leaf}_primitive_indices"])
    assert np.array_equal(t["parents"], g[f"{name}_leaf{leaf}_parents"])
    for k in ("node_lowers", "node_uppers"):
        for f in ("x", "y", "z", "ib"):
            assert np.array_equal(t[k][f], g[f"{name}_leaf{leaf}_{k}"][f]), (k, f)


def test_c3_heightfield_10m_build_bit_exact(wp, oracle_mod):
    """Config C3 mesh (9 999 392 triangles): 68 % duplicate 30-bit keys, depth 44 -> the depth >= 32 rule and
    the parity tie-break decide a large part of the topology.  Full diff of all 2N-1 nodes."""
    P, I = mg.heightfield(2237, 4)
    m = gpu_mesh(wp, P, I, 4)
    got = m.download_tree()
    assert got["deep"] == 1 and got["height"] >= 32
    want = oracle_mod.mesh_lbvh_build(P, I, 4)
    assert_tree_equal(got, want)
    # refit after a vertical deformation: visible boxes equal the oracle's
    P2 = P.copy()
    P2[:, 2] += 0.01 * np.sin(40.0 * P2[:, 0]).astype(np.float32)
    m.points.assign(P2)
    m.refit()
    got = m.download_tree()
    lo2, hi2 = oracle_mod.triangle_bounds(P2, I)
    oracle_mod.lbvh_refit(want, lo2, hi2)
    lo, hi = want["node_lowers"], want["node_uppers"]
    # walk the visible tree iteratively (20 M nodes: vectorised frontier expansion)
    frontier = np.array([want["root"]])
    vis = []
    while frontier.size:
        vis.append(frontier)
        inner = frontier[(lo["ib"][frontier] >> 31) == 0]
        frontier = np.concatenate([lo["ib"][inner] & 0x7FFFFFFF, hi["ib"][inner] & 0x7FFFFFFF]).astype(np.int64)
    vis = np.concatenate(vis)
    for name in ("node_lowers", "node_uppers"):
        for f in "xyz":
            assert np.array_equal(got[name][f][vis], want[name][f][vis]), (name, f)


def test_grouped_bvh_build_bit_exact(wp, oracle_mod):
    """Grouped trees (key = group << 32 | code, group-aware parent choice, no packed leaf across groups:
    bvh.cu:205-209, 296-321, 431-437) against the oracle; procedure of warp/tests/geometry/test_grouped_bvh.py."""
    rng = np.random.default_rng(7)
    for n, ngroups, leaf in ((64, 4, 1), (1000, 7, 4), (5000, 3, 2), (300, 300, 4)):
        lo, hi = random_boxes(n, seed=n + 1)
        groups = rng.integers(0, ngroups, n).astype(np.int32)
        b = wp.Bvh(wp.array(lo, dtype=wp.vec3), wp.array(hi, dtype=wp.vec3), groups=wp.array(groups, dtype=wp.int32), leaf_size=leaf)
        got = b.download_tree()
        assert got["key_bits"] == 64
        want = oracle_mod.lbvh_build(lo, hi, leaf, groups=groups)
        assert_tree_equal(got, want)
        # every visible leaf holds items of a single group
        for c in visible_nodes(got):
            if got["node_lowers"]["ib"][c] >> 31:
                s, e = int(got["node_lowers"]["ib"][c] & 0x7FFFFFFF), int(got["node_uppers"]["ib"][c] & 0x7FFFFFFF)
                assert len(set(groups[got["primitive_indices"][s:e]].tolist())) == 1
        lo2, hi2 = lo + 1.0, hi + 2.0
        b.lowers.assign(lo2), b.uppers.assign(hi2)
        b.refit()
        got = b.download_tree()
        oracle_mod.lbvh_refit(want, lo2, hi2)
        vis = visible_nodes(want)
        for f in "xyz":
            assert np.array_equal(got["node_lowers"][f][vis], want["node_lowers"][f][vis])
            assert np.array_equal(got["node_uppers"][f][vis], want["node_uppers"][f][vis])


def test_grouped_mesh(wp, oracle_mod):
    P, I = mg.noisy_sphere(4, 0.03, 2)
    T = len(I) // 3
    groups = (np.arange(T) * 5 // T).astype(np.int32)
    m = wp.Mesh(wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32), groups=wp.array(groups, dtype=wp.int32))
    want = oracle_mod.mesh_lbvh_build(P, I, 4, groups=groups)
    assert_tree_equal(m.download_tree(), want)
    Q = mg.box_queries(P, 5000, seed=3)
    got = wp.mesh_query_point(m, Q, 1e6).numpy()
    ref = oracle_mod.query_point(P, I, want, Q, 1e6)
    for k in ("result", "sign", "face", "u", "v"):
        assert np.array_equal(got[k], ref[k]), k


@pytest.mark.parametrize("name", ["ico5_noisy", "height65"])
def test_morton63_quality_mode(wp, oracle_mod, name):
    """63-bit Morton option (NOT a reference parity mode): same pipeline, wider keys; checked against the
    oracle run with the same key width, and queries on it against the oracle traversal of that tree."""
    P, I = MESHES[name]()
    m = wp.Mesh(wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32), morton_bits=63)
    got = m.download_tree()
    assert got["key_bits"] == 64
    want = oracle_mod.mesh_lbvh_build(P, I, 4, morton_bits=63)
    assert_tree_equal(got, want)
    Q = mg.box_queries(P, 5000, seed=3)
    a, b = wp.mesh_query_point_no_sign(m, Q, 1e6).numpy(), oracle_mod.query_point_no_sign(P, I, want, Q, 1e6)
    for k in ("result", "face", "u", "v"):
        assert np.array_equal(a[k], b[k]), k
    m30 = wp.Mesh(wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32))
    assert m30.download_tree()["key_bits"] == 32  # the option does not leak into later trees


def test_morton63_removes_duplicate_keys_on_10m_heightfield(wp):
    P, I = mg.heightfield(2237, 4)
    m = wp.Mesh(wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32), morton_bits=63)
    t = m.download_tree()
    keys = t["keys"]
    # what is left are the two triangles of a quad whose AABBs coincide (same box centre by construction): pairs, not runs
    dup, run3 = (keys[1:] == keys[:-1]).mean(), (keys[2:] == keys[:-2]).mean()
    assert dup < 0.3 and run3 < 0.05, (dup, run3)
    t30 = wp.Mesh(m.points, m.indices).download_tree()
    k30 = t30["keys"]
    assert (k30[1:] == k30[:-1]).mean() > 0.6 and t30["height"] > t["height"]  # 30-bit keys: long runs, deeper tree
    assert np.array_equal(np.sort(t["primitive_indices"]), np.arange(len(I) // 3))


# ------------------------------------------------------------------------------------------------
# refit variants: atomic arrival counters vs the planned wavefront (levels in shared memory)
# ------------------------------------------------------------------------------------------------
@pytest.fixture
def refit_mode(wp):
    from warp_b200 import _lib

    core = _lib.core()

    def set_mode(mode):
        core.wp_b200_set_refit_mode(mode)

    yield set_mode
    core.wp_b200_set_refit_mode(0)


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("leaf", [1, 4, 8])
def test_refit_variants_mesh(wp, oracle_mod, refit_mode, mode, leaf):
    """Both refit variants give the oracle's boxes on every visible node, refit after refit (the counters are
    never cleared), and after an in-place rebuild (the wavefront plan is rebuilt)."""
    refit_mode(mode)
    P, I = mg.noisy_sphere(5, 0.02, 11)  # 20 480 triangles = 20 wavefront blocks
    pts = wp.array(P, dtype=wp.vec3)
    m = wp.Mesh(pts, wp.array(I, dtype=wp.int32), bvh_leaf_size=leaf)
    want = oracle_mod.mesh_lbvh_build(P, I, leaf)
    for k in range(3):
        _refit_and_compare(wp, oracle_mod, m, pts, mg.renoise_sphere(P, 0.03 * (k + 1), 20 + k), I, want)
    P3 = mg.renoise_sphere(P, 0.2, 30)
    pts.assign(P3)
    m.rebuild()
    want = oracle_mod.mesh_lbvh_build(P3, I, leaf)
    assert_tree_equal(m.download_tree(), want)
    _refit_and_compare(wp, oracle_mod, m, pts, mg.renoise_sphere(P, 0.05, 31), I, want)


@pytest.mark.parametrize("mode", [1, 2])
def test_refit_variants_deep_grouped_and_tiny(wp, oracle_mod, refit_mode, mode):
    refit_mode(mode)
    rng = np.random.default_rng(5)
    # coincident boxes: depth-rule leaves with hundreds of items, muted subtrees that are not flagged themselves
    n = 5000
    lo = np.zeros((n, 3), np.float32)
    lo[: n // 2] += 1.0
    lo[n // 3 : n // 2, 1] += 2.0
    hi = lo + 0.5
    cases = [(lo, hi, None, 1), (lo, hi, None, 4)]
    # grouped tree, items of a group scattered; tiny trees (root is a packed leaf / a single block)
    glo, ghi = random_boxes(7000, seed=6)
    cases.append((glo, ghi, rng.integers(0, 9, 7000).astype(np.int32), 4))
    for m_ in (1, 2, 3, 5, 9):
        tlo, thi = random_boxes(m_, seed=m_)
        cases.append((tlo, thi, None, 4))
    for lo, hi, groups, leaf in cases:
        lo_d, hi_d = wp.array(lo, dtype=wp.vec3), wp.array(hi, dtype=wp.vec3)
        b = wp.Bvh(lo_d, hi_d, leaf_size=leaf, groups=None if groups is None else wp.array(groups, dtype=wp.int32))
        want = oracle_mod.lbvh_build(lo, hi, leaf, groups=groups)
        for k in range(2):
            d = rng.standard_normal(lo.shape).astype(np.float32) * np.float32(0.3)
            lo2, hi2 = (lo + d).astype(np.float32), (hi + d + np.float32(0.1 * k)).astype(np.float32)
            lo_d.assign(lo2), hi_d.assign(hi2)
            b.refit()
            oracle_mod.lbvh_refit(want, lo2, hi2)
            got = b.download_tree()
            vis = visible_nodes(want)
            for name in ("node_lowers", "node_uppers"):
                assert np.array_equal(got[name]["ib"], want[name]["ib"])
                for f in "xyz":
                    assert np.array_equal(got[name][f][vis], want[name][f][vis]), (len(lo), leaf, name, f)


@pytest.mark.parametrize("n", [2, 3, 1023, 1024, 1025, 2047, 2048, 2049, 4097, 6144])
def test_block_boundary_sizes_build_and_refit(wp, oracle_mod, refit_mode, n):
    """Item counts around the block sizes of the hierarchy kernel (2048) and of the wavefront refit (1024):
    full tree diff after the build, visible boxes after a refit in both modes."""
    rng = np.random.default_rng(n)
    lo, hi = random_boxes(n, seed=n)
    for leaf in (1, 4):
        lo_d, hi_d = wp.array(lo, dtype=wp.vec3), wp.array(hi, dtype=wp.vec3)
        b = wp.Bvh(lo_d, hi_d, leaf_size=leaf)
        want = oracle_mod.lbvh_build(lo, hi, leaf)
        assert_tree_equal(b.download_tree(), want)
        for mode in (1, 2):
            refit_mode(mode)
            d = rng.standard_normal(lo.shape).astype(np.float32)
            lo2, hi2 = (lo + d).astype(np.float32), (hi + d).astype(np.float32)
            lo_d.assign(lo2), hi_d.assign(hi2)
            b.refit()
            oracle_mod.lbvh_refit(want, lo2, hi2)
            got = b.download_tree()
            vis = visible_nodes(want)
            for name in ("node_lowers", "node_uppers"):
                for f in "xyz":
                    assert np.array_equal(got[name][f][vis], want[name][f][vis]), (n, leaf, mode, name, f)


def test_experiment_parallel_topology(wp, oracle_mod):
    """EXPERIMENT: parents of all internal nodes recomputed from the sorted keys by the
    dependency-free k_topology kernel equal the parents the builder produced bottom-up -- 30-bit, 63-bit and grouped
    keys, clustered duplicates -- and a run of equal keys longer than the prototype replays is reported, not mangled."""
    import ctypes

    from warp_b200 import _lib

    fn = _lib.core().wp_b200_experiment_parallel_topology

    def check(tree_obj, n, expect_ok=True):
        got = wp.empty(n - 1, wp.int32, tree_obj.device)
        us = fn(tree_obj.id, ctypes.c_void_p(got.ptr), 3)
        if not expect_ok:
            assert us == -2.0
            return us
        assert us > 0, us
        want = tree_obj.download_tree()["parents"][n:].astype(np.int32)
        assert np.array_equal(got.numpy(), want)
        return us

    P, I = mg.noisy_sphere(6, 0.02, 1)
    for bits in (30, 63):
        m = wp.Mesh(wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32), morton_bits=bits)
        check(m, len(I) // 3)
    rng = np.random.default_rng(5)
    c = (rng.random((300, 3)) * 4).astype(np.float32)
    c = (np.repeat(c, 10, axis=0) + (rng.random((3000, 3)) * 1e-4).astype(np.float32)).astype(np.float32)
    lo, hi = wp.array(c, dtype=wp.vec3), wp.array(c + np.float32(0.01), dtype=wp.vec3)
    check(wp.Bvh(lo, hi, leaf_size=1), 3000)
    groups = wp.array((np.arange(3000) // 450).astype(np.int32), dtype=wp.int32)
    check(wp.Bvh(lo, hi, groups=groups, leaf_size=2), 3000)
    same = np.tile(np.array([[1.0, 2.0, 3.0]], np.float32), (257, 1))
    lo2, hi2 = wp.array(np.concatenate([same, c[:100]]), dtype=wp.vec3), wp.array(np.concatenate([same + 1, c[:100] + 1]), dtype=wp.vec3)
    check(wp.Bvh(lo2, hi2), 357, expect_ok=False)
    # timing at C2 size (1.31 M triangles) and on the 10 M-triangle heightfield, beside the merge kernel's 135 / 730 us
    P, I = mg.noisy_sphere(8, 0.02, 1)
    m = wp.Mesh(wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32))
    us_c2 = check(m, len(I) // 3)
    P, I = mg.heightfield(2237, 4)
    m = wp.Mesh(wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32))
    us_10m = check(m, len(I) // 3)
    print(f"k_topology: {us_c2:.1f} us at 1.31 M triangles, {us_10m:.1f} us at 10 M")


@pytest.mark.parametrize("n", [(1 << 24) - 1, 1 << 24, (1 << 24) + 1])
def test_refit_wavefront_at_the_sort_tile_switch(wp, oracle_mod, n):
    """n = 2^24 is where the build sort switches to 4096-key tiles while the refit plan (n - 1 keys) still sorts with
    2048-key tiles: the plan's look-back words must fit the arena (they used to run into `heights`).  The planned
    wavefront refit (per-object option, no process-wide toggle) must give the boxes of the atomic refit on every
    node, and at n = 2^24 the oracle's on every visible node."""
    rng = np.random.default_rng(n & 0xFFFF)
    lo = (rng.random((n, 3), dtype=np.float32) * np.float32(50.0)).astype(np.float32)
    hi = (lo + rng.random((n, 3), dtype=np.float32) * np.float32(0.2)).astype(np.float32)
    lo_d, hi_d = wp.array(lo, dtype=wp.vec3), wp.array(hi, dtype=wp.vec3)
    b = wp.Bvh(lo_d, hi_d, leaf_size=4)
    assert b.get_option("refit_mode") == -1
    d = (rng.standard_normal((n, 3)).astype(np.float32) * np.float32(0.1)).astype(np.float32)
    lo2, hi2 = (lo + d).astype(np.float32), (hi + d).astype(np.float32)
    lo_d.assign(lo2), hi_d.assign(hi2)
    got = {}
    for mode in (2, 1):
        b.set_option("refit_mode", mode)
        assert b.get_option("refit_mode") == mode
        b.refit()
        got[mode] = b.download_tree()
    for name in ("node_lowers", "node_uppers"):
        for f in ("ib", "x", "y", "z"):
            assert np.array_equal(got[2][name][f], got[1][name][f]), (n, name, f)
    assert np.array_equal(got[2]["parents"], got[1]["parents"])
    if n == 1 << 24:
        want = oracle_mod.lbvh_build(lo, hi, 4)
        assert np.array_equal(got[2]["parents"], want["parents"])
        oracle_mod.lbvh_refit(want, lo2, hi2)
        # visible nodes = not below a packed leaf: a node is muted iff its parent is a leaf or muted (top-down order
        # is not available cheaply at this size, so compare the leaf flags and the boxes of every node whose parent
        # is an inner node)
        par = want["parents"]
        inner_parent = np.ones(len(par), bool)
        has_parent = par >= 0
        inner_parent[has_parent] = (want["node_lowers"]["ib"][par[has_parent]] >> 31) == 0
        muted = np.zeros(len(par), bool)
        for _ in range(64):  # propagate "muted" down the (at most 64 deep below a leaf) chains
            new = muted.copy()
            new[has_parent] |= ~inner_parent[has_parent] | muted[par[has_parent]]
            if np.array_equal(new, muted):
                break
            muted = new
        vis = ~muted
        for name in ("node_lowers", "node_uppers"):
            assert np.array_equal(got[2][name]["ib"], want[name]["ib"])
            for f in "xyz":
                assert np.array_equal(got[2][name][f][vis], want[name][f][vis]), (name, f)


def test_small_node_pass_builds_the_same_tree(wp, oracle_mod):
    """The Karras-style small-node pass (k_small_nodes: ranges of small distinct-key nodes found independently from the
    sorted keys; off by default) followed by the merge from the finished subtrees gives the oracle's
    tree on all 2N - 1 nodes: mesh with distinct keys, mesh full of equal-key pairs, 63-bit keys, groups, leaf sizes."""
    from warp_b200 import _lib

    core = _lib.core()
    assert core.wp_b200_set_experiment(b"small_nodes", 1)
    try:
        P, I = mg.noisy_sphere(6, 0.02, 3)
        for leaf in (1, 4, 8):
            m = wp.Mesh(wp.array(P, dtype=wp.vec3), wp.array(I, dtype=wp.int32), bvh_leaf_size=leaf)
            assert_tree_equal(m.download_tree(), oracle_mod.mesh_lbvh_build(P, I, leaf))
        Ph, Ih = mg.heightfield(300, 4)  # equal-key pairs everywhere (the two triangles of a quad)
        for bits in (30, 63):
            m = wp.Mesh(wp.array(Ph, dtype=wp.vec3), wp.array(Ih, dtype=wp.int32), morton_bits=bits)
            assert_tree_equal(m.download_tree(), oracle_mod.mesh_lbvh_build(Ph, Ih, 4, morton_bits=bits))
        for n in (65, 1023, 1024, 1025, 2049, 5000):
            lo, hi = random_boxes(n, seed=n)
            groups = (np.arange(n) % 7).astype(np.int32)
            b = wp.Bvh(wp.array(lo, dtype=wp.vec3), wp.array(hi, dtype=wp.vec3), leaf_size=2)
            assert_tree_equal(b.download_tree(), oracle_mod.lbvh_build(lo, hi, 2))
            g = wp.Bvh(wp.array(lo, dtype=wp.vec3), wp.array(hi, dtype=wp.vec3), leaf_size=2, groups=wp.array(groups, dtype=wp.int32))
            assert_tree_equal(g.download_tree(), oracle_mod.lbvh_build(lo, hi, 2, groups=groups))
            # refit after the build (atomic and wavefront): visible boxes
            d = np.float32(0.25)
            lo2, hi2 = (lo + d).astype(np.float32), (hi + d).astype(np.float32)
            want = oracle_mod.lbvh_build(lo, hi, 2)
            oracle_mod.lbvh_refit(want, lo2, hi2)
            b.lowers.assign(lo2), b.uppers.assign(hi2)
            for mode in (1, 2):
                b.set_option("refit_mode", mode)
                b.refit()
                got = b.download_tree()
                vis = visible_nodes(want)
                for name in ("node_lowers", "node_uppers"):
                    for f in "xyz":
                        assert np.array_equal(got[name][f][vis], want[name][f][vis]), (n, mode, name, f)
    finally:
        core.wp_b200_set_experiment(b"small_nodes", -1)
