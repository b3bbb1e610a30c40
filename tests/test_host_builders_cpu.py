"""CPU-only pin of the host constructors (csrc/host_build.cu, restating warp/native/bvh.cpp:216-572): the item order and
the leaf boundaries of the sah / median trees equal those of the reference's own host builder (oracle/_ref, the unmodified
reference C++), for several meshes and leaf sizes.  Needs no GPU: wp_b200_host_build_order is host code."""
import ctypes

import numpy as np
import pytest

from warp_b200 import _lib, meshgen as mg


def _ours(lo, hi, leaf, ctor):
    lo, hi = np.ascontiguousarray(lo, np.float32), np.ascontiguousarray(hi, np.float32)
    n = len(lo)
    order, starts = np.zeros(n, np.int32), np.zeros(n, np.uint8)
    depth = _lib.core().wp_b200_host_build_order(lo.ctypes.data, hi.ctypes.data, n, leaf, ctor, order.ctypes.data, starts.ctypes.data)
    assert depth >= 0
    return order, starts, depth


def _ref_leaf_starts(tree):
    lo, hi = tree["node_lowers"], tree["node_uppers"]
    starts = np.zeros(tree["n"], np.uint8)
    for c in range(tree["num_nodes"]):
        if lo["ib"][c] >> 31:
            starts[lo["ib"][c] & 0x7FFFFFFF] = 1
    return starts


@pytest.mark.parametrize("ctor", [0, 1])
@pytest.mark.parametrize("leaf", [1, 4, 8])
def test_item_order_and_leaves_match_the_reference_host_builder(oracle_mod, ctor, leaf):
    if not oracle_mod.ref_available():
        pytest.skip("oracle/_ref is not built")
    cases = [mg.noisy_sphere(3, 0.05, 2), mg.heightfield(40, 4), mg.cloth(33, 2), (mg.CUBE_POINTS, mg.CUBE_INDICES_RH)]
    P = np.tile(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32), (50, 1))  # coincident triangles: degenerate planes
    P[60:] += np.array([3, 0, 0], np.float32)
    cases.append((P, np.arange(len(P), dtype=np.int32)))
    for P, I in cases:
        lo, hi = oracle_mod.triangle_bounds(P, I)
        order, starts, depth = _ours(lo, hi, leaf, ctor)
        ref = oracle_mod.RefMesh(P, I, ctor, leaf).tree()
        assert np.array_equal(order, ref["primitive_indices"]), "item order differs from the reference host builder"
        assert np.array_equal(starts, _ref_leaf_starts(ref)), "leaf boundaries differ"
        assert sorted(order.tolist()) == list(range(len(order)))


def test_bad_arguments_are_refused():
    z = np.zeros(3, np.float32)
    out = np.zeros(1, np.int32)
    f = _lib.core().wp_b200_host_build_order
    assert f(z.ctypes.data, z.ctypes.data, 0, 4, 0, out.ctypes.data, None) == -1
    assert f(z.ctypes.data, z.ctypes.data, 1, 0, 0, out.ctypes.data, None) == -1
    assert f(z.ctypes.data, z.ctypes.data, 1, 4, 2, out.ctypes.data, None) == -1
    assert f(z.ctypes.data, z.ctypes.data, 1, 4, 1, out.ctypes.data, None) == 0 and out[0] == 0
