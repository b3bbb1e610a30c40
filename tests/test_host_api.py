"""CPU tests of the host-side mirror of the reference's Mesh / Bvh validation
(warp/_src/types.py:5891-5945, 6132-6191; reference tests warp/tests/geometry/test_mesh.py:386-440)."""
import numpy as np
import pytest

import warp_b200 as wp
from warp_b200.types import Device


class FakeArray:
    """Duck-typed stand-in so validation can be exercised without a device."""

    def __init__(self, n, dtype, device="cuda:0", ndim=1, contiguous=True):
        self.shape = (n,) if ndim == 1 else (n, 3)
        self.ndim = ndim
        self.size = n * (3 if ndim == 2 else 1)
        self.dtype = dtype
        self.device = Device(device)
        self.is_contiguous = contiguous
        self.ptr = 0

    def __len__(self):
        return self.shape[0]

    def __bool__(self):
        return True


def test_bvh_constructor_enum():
    assert int(wp.BvhConstructor.SAH) == 0 and int(wp.BvhConstructor.MEDIAN) == 1
    assert int(wp.BvhConstructor.LBVH) == 2 and int(wp.BvhConstructor.CUBQL) == -1
    assert wp.BvhConstructor.from_str("lbvh") is wp.BvhConstructor.LBVH
    with pytest.raises(ValueError, match="Unknown BVH constructor"):
        wp.BvhConstructor.from_str("octree")


def test_mesh_validation_messages():
    pts, idx = FakeArray(8, wp.vec3), FakeArray(36, wp.int32)
    with pytest.raises(RuntimeError, match="must live on the same device"):
        wp.Mesh(pts, FakeArray(36, wp.int32, device="cuda:1"))
    with pytest.raises(RuntimeError, match="points should be a contiguous array of type wp.vec3"):
        wp.Mesh(FakeArray(8, wp.float32), idx)
    with pytest.raises(RuntimeError, match="points should be a contiguous array of type wp.vec3"):
        wp.Mesh(FakeArray(8, wp.vec3, contiguous=False), idx)
    with pytest.raises(RuntimeError, match="velocities should be a contiguous array of type wp.vec3"):
        wp.Mesh(pts, idx, velocities=FakeArray(8, wp.float32))
    with pytest.raises(RuntimeError, match="indices should be a contiguous array of type wp.int32"):
        wp.Mesh(pts, FakeArray(36, wp.float32))
    with pytest.raises(RuntimeError, match="flattened 1d array"):
        wp.Mesh(pts, FakeArray(12, wp.int32, ndim=2))
    with pytest.raises(RuntimeError, match="groups must have the same length as indices / 3"):
        wp.Mesh(pts, idx, groups=FakeArray(5, wp.int32))
    with pytest.raises(ValueError, match="bvh_leaf_size must be greater than or equal to 1"):
        wp.Mesh(pts, idx, bvh_leaf_size=0)
    with pytest.raises(ValueError, match="Unknown BVH constructor"):
        wp.Mesh(pts, idx, bvh_constructor="nope")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        wp.Mesh(FakeArray(8, wp.vec3, device="cpu"), FakeArray(36, wp.int32, device="cpu"))


def test_bvh_validation_messages():
    lo, hi = FakeArray(10, wp.vec3), FakeArray(10, wp.vec3)
    with pytest.raises(RuntimeError, match="same number of lower and upper bounds"):
        wp.Bvh(lo, FakeArray(9, wp.vec3))
    with pytest.raises(RuntimeError, match="must live on the same device"):
        wp.Bvh(lo, FakeArray(10, wp.vec3, device="cuda:1"))
    with pytest.raises(RuntimeError, match="lowers should be a contiguous array of type wp.vec3"):
        wp.Bvh(FakeArray(10, wp.float32), hi)
    with pytest.raises(RuntimeError, match="uppers should be a contiguous array of type wp.vec3"):
        wp.Bvh(lo, FakeArray(10, wp.float32))
    with pytest.raises(RuntimeError, match="groups must have the same length"):
        wp.Bvh(lo, hi, groups=FakeArray(3, wp.int32))
    with pytest.raises(ValueError, match="leaf_size must be greater than or equal to 1"):
        wp.Bvh(lo, hi, leaf_size=0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        wp.Bvh(FakeArray(10, wp.vec3, device="cpu"), FakeArray(10, wp.vec3, device="cpu"))


def test_meshgen_shapes():
    from warp_b200 import meshgen as mg

    p, i = mg.icosphere(3)
    assert i.size // 3 == 20 * 4**3 and p.shape[0] == 10 * 4**3 + 2
    assert np.allclose(np.linalg.norm(p, axis=1), 1.0, atol=1e-6)
    p, i = mg.heightfield(17)
    assert i.size // 3 == 2 * 16 * 16 and i.max() == p.shape[0] - 1
    s, d = mg.pinhole_rays(8, 4)
    assert s.shape == d.shape == (32, 3) and np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-6)


def test_query_surface_mirrors_the_reference_builtins():
    """Every batched query keeps the name of the Warp builtin it evaluates (warp/_src/builtins.py mesh_query_* /
    mesh_eval_* / bvh_query_* / bvh_get_group_root) and says in its docstring which reference lines it follows."""
    import re

    import warp_b200 as wp

    names = [
        "mesh_query_point", "mesh_query_point_no_sign", "mesh_query_point_sign_normal", "mesh_query_point_sign_parity",
        "mesh_query_furthest_point_no_sign", "mesh_query_ray", "mesh_query_ray_anyhit", "mesh_query_ray_count_intersections",
        "mesh_query_aabb", "mesh_query_sphere", "mesh_eval_position", "mesh_eval_velocity", "mesh_eval_face_normal",
        "bvh_query_aabb", "bvh_query_ray", "bvh_query_sphere", "bvh_query_capsule", "bvh_get_group_root",
    ]  # fmt: skip
    for name in names:
        fn = getattr(wp, name)
        assert callable(fn), name
        assert re.search(r"(mesh|bvh|intersect)\.h:\d+", fn.__doc__ or ""), f"{name}: docstring cites no reference line"
    # not built: asking for them fails loudly instead of answering something else
    assert not hasattr(wp, "mesh_query_point_sign_winding_number")
