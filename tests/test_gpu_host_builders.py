"""The host constructors ("sah", "median": SURVEY.md 8f rank 4; warp/native/bvh.cpp:216-572, bvh.cu:625-652) uploaded into
the sibling-pair layout: the item order must be the reference's own (same partitions, same order inside the leaves), and
every query must return what the reference's traversal returns on ITS tree of the same constructor -- bit for bit."""
import numpy as np
import pytest

from helpers import assert_results_equal, random_boxes
from warp_b200 import meshgen as mg

CTOR = {"sah": 0, "median": 1}


def _meshes():
    yield "sphere", mg.noisy_sphere(4, 0.03, 5)
    yield "heightfield", mg.heightfield(48, 4)
    # duplicates and zero-extent axes: coincident triangles (the SAH plane degenerates, the split falls back to the middle)
    P = np.tile(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), (70, 1))
    P[90:] += np.array([2, 0, 0], np.float32)
    yield "coincident", (P, np.arange(len(P), dtype=np.int32))


@pytest.mark.gpu
@pytest.mark.parametrize("ctor", ["sah", "median"])
@pytest.mark.parametrize("leaf", [1, 4])
def test_host_constructor_meshes_match_the_reference(wp, oracle_mod, ctor, leaf):
    if not oracle_mod.ref_available():
        pytest.skip("oracle/_ref (the reference's own host builders) is not built on this box")
    rng = np.random.default_rng(3)
    for name, (P, I) in _meshes():
        pts = wp.array(P, dtype=wp.vec3)
        m = wp.Mesh(pts, wp.array(I, dtype=wp.int32), bvh_constructor=ctor, bvh_leaf_size=leaf)
        ref = oracle_mod.RefMesh(P, I, CTOR[ctor], leaf)
        assert np.array_equal(m.download_primitive_indices(), ref.tree()["primitive_indices"]), (name, "item order")
        Q = mg.box_queries(P, 6000, seed=8)
        S, D = mg.random_rays(P, 6000, seed=9)
        for max_dist in (1.0e6, 0.05):
            assert_results_equal(wp.mesh_query_point_no_sign(m, Q, max_dist).numpy(), ref.query_point_no_sign(Q, max_dist),
                                 ("result", "face", "u", "v"))
        assert_results_equal(wp.mesh_query_point(m, Q, 1.0e6).numpy(), ref.query_point(Q, 1.0e6), ("result", "sign", "face", "u", "v"))
        assert_results_equal(wp.mesh_query_ray(m, S, D, 1.0e6).numpy(), ref.query_ray(S, D, 1.0e6),
                             ("result", "sign", "face", "t", "u", "v", "normal"))
        # refit (always the counter climb on these trees) after moving the vertices
        P2 = (P + rng.normal(0, 0.01, P.shape)).astype(np.float32)
        pts.assign(P2)
        m.refit()
        ref.points[:] = P2
        ref.refit()
        assert_results_equal(wp.mesh_query_point_no_sign(m, Q, 1.0e6).numpy(), ref.query_point_no_sign(Q, 1.0e6),
                             ("result", "face", "u", "v"))
        assert_results_equal(wp.mesh_query_ray(m, S, D, 1.0e6).numpy(), ref.query_ray(S, D, 1.0e6),
                             ("result", "sign", "face", "t", "u", "v", "normal"))
        # an in-place rebuild is an LBVH, whatever built the tree first (bvh.cu:819-843)
        m.rebuild()
        want = oracle_mod.mesh_lbvh_build(P2, I, leaf)
        assert np.array_equal(m.download_tree()["primitive_indices"], want["primitive_indices"])
        pts.assign(P)
        with pytest.raises(RuntimeError, match="mirror is not available"):
            wp.Mesh(pts, m.indices, bvh_constructor=ctor, bvh_leaf_size=leaf).download_tree()


@pytest.mark.gpu
@pytest.mark.parametrize("ctor", ["sah", "median"])
def test_host_constructor_bvh_queries_match_brute_force(wp, ctor):
    """wp.Bvh with a host constructor: the generic iterators return exactly the overlapping items (as sets -- the
    iterator order follows the tree), also after refit and after the in-place rebuild."""
    for n, leaf in ((1, 1), (2, 1), (97, 1), (3000, 2), (3000, 8)):
        lo, hi = random_boxes(n, seed=n + leaf)
        lo_d, hi_d = wp.array(lo, dtype=wp.vec3), wp.array(hi, dtype=wp.vec3)
        b = wp.Bvh(lo_d, hi_d, constructor=ctor, leaf_size=leaf)
        qlo, qhi = random_boxes(200, seed=77, size=2.0)
        for step in range(3):
            lists = wp.bvh_query_aabb(b, qlo, qhi).lists()
            ov = ((lo[None] <= qhi[:, None]) & (hi[None] >= qlo[:, None])).all(-1)
            for q in range(200):
                assert set(lists[q].tolist()) == set(np.flatnonzero(ov[q]).tolist()), (n, leaf, step, q)
            lo, hi = random_boxes(n, seed=1000 + n + step)
            lo_d.assign(lo), hi_d.assign(hi)
            if step == 0:
                b.refit()
            else:
                b.rebuild()
