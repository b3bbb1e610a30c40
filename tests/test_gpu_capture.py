"""CUDA-graph capture of the per-step loop (refit / in-place rebuild + device-buffer queries): one graph launch
per frame gives the same answers as the eager calls and as the oracle (BASELINE config 4;
the reference's capture tests: warp/tests/geometry/test_bvh.py rebuild-under-capture, test_mesh.py)."""
import numpy as np
import pytest

from helpers import assert_results_equal
from warp_b200 import meshgen as mg
from warp_b200.queries import MeshQueryPoint, MeshQueryRay

pytestmark = pytest.mark.gpu

POINT_FIELDS = ("result", "face", "u", "v")
RAY_FIELDS = ("result", "sign", "face", "t", "u", "v", "normal")


def _outs(wp, n):
    e = wp.empty
    return (MeshQueryPoint(e(n, wp.uint8), wp.zeros(n, wp.float32), e(n, wp.int32), e(n, wp.float32), e(n, wp.float32)),
            MeshQueryRay(e(n, wp.uint8), e(n, wp.float32), e(n, wp.int32), e(n, wp.float32), e(n, wp.float32),
                         e(n, wp.float32), e(n, wp.vec3)))  # fmt: skip


def test_refit_and_queries_in_one_graph(wp, oracle_mod):
    P, I = mg.noisy_sphere(4, 0.02, 31)
    n = 40000  # above the Morton-ordering threshold: the captured graph contains the ordering sort too
    pts = wp.array(P, dtype=wp.vec3)
    m = wp.Mesh(pts, wp.array(I, dtype=wp.int32), bvh_constructor="lbvh")
    Q = mg.box_queries(P, n, seed=32)
    S, D = mg.random_rays(P, n, seed=33)
    Qd, Sd, Dd = (wp.array(a, dtype=wp.vec3) for a in (Q, S, D))
    pout, rout = _outs(wp, n)
    # one eager pass first: grow-only scratch gets allocated outside the capture
    m.refit()
    wp.mesh_query_point_no_sign(m, Qd, 1e6, out=pout)
    wp.mesh_query_ray(m, Sd, Dd, 1e6, out=rout)
    with wp.ScopedCapture() as cap:
        m.refit()
        wp.mesh_query_point_no_sign(m, Qd, 1e6, out=pout)
        wp.mesh_query_ray(m, Sd, Dd, 1e6, out=rout)
    assert cap.graph is not None
    tree = oracle_mod.mesh_lbvh_build(P, I, 4)
    for frame in range(3):
        P2 = (P * np.float32(1.0 + 0.1 * frame) + np.float32(0.05 * frame)).astype(np.float32)
        pts.assign(P2)
        wp.capture_launch(cap.graph)
        lo2, hi2 = oracle_mod.triangle_bounds(P2, I)
        oracle_mod.lbvh_refit(tree, lo2, hi2)
        assert_results_equal(pout.numpy(), oracle_mod.query_point_no_sign(P2, I, tree, Q, 1e6), POINT_FIELDS)
        assert_results_equal(rout.numpy(), oracle_mod.query_ray(P2, I, tree, S, D, 1e6), RAY_FIELDS)


def test_rebuild_under_capture(wp, oracle_mod):
    P, I = mg.noisy_sphere(3, 0.05, 41)
    pts = wp.array(P, dtype=wp.vec3)
    m = wp.Mesh(pts, wp.array(I, dtype=wp.int32), bvh_constructor="lbvh")
    n = 5000
    Q = mg.box_queries(P, n, seed=42)
    Qd = wp.array(Q, dtype=wp.vec3)
    pout, _ = _outs(wp, n)
    stream = wp.Stream()
    with wp.ScopedStream(stream):
        wp.capture_begin()
        m.rebuild()
        wp.mesh_query_point_no_sign(m, Qd, 1e6, out=pout)
        graph = wp.capture_end()
    rng = np.random.default_rng(43)
    for _ in range(2):
        P2 = (P + rng.standard_normal(P.shape).astype(np.float32) * np.float32(0.2)).astype(np.float32)
        pts.assign(P2)
        wp.capture_launch(graph, stream)
        stream.synchronize()
        tree = oracle_mod.mesh_lbvh_build(P2, I, 4)
        assert_results_equal(pout.numpy(), oracle_mod.query_point_no_sign(P2, I, tree, Q, 1e6), POINT_FIELDS)
        t = m.download_tree()
        assert np.array_equal(t["primitive_indices"], tree["primitive_indices"])


def test_capture_on_default_stream_is_refused(wp):
    with pytest.raises(RuntimeError):
        wp.capture_begin()
