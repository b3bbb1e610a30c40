"""warp_b200 -- B200-native (sm_100a) mesh / BVH spatial queries behind Warp's ``Mesh`` / ``Bvh`` API.

Drop-in for ONE path of NVIDIA/warp: ``wp.Mesh(points, indices, bvh_constructor="lbvh")`` /
``wp.Bvh(lowers, uppers)`` with ``.id`` / ``.refit()`` / ``.rebuild()``, plus batched equivalents of
``mesh_query_point_no_sign``, ``mesh_query_point`` and ``mesh_query_ray``.  Pure-Python host code
(numpy only) over a ctypes C ABI (``include/warp_b200.h``) into hand-written CUDA.  No Warp
codegen / NVRTC, no Triton, no torch, no CPU fallback.
"""

from .types import (  # noqa: F401
    Bvh,
    BvhConstructor,
    Device,
    Mesh,
    array,
    bool_,
    empty,
    float32,
    from_numpy,
    get_device,
    int32,
    synchronize,
    synchronize_device,
    uint8,
    uint32,
    uint64,
    vec3,
    zeros,
)
from .queries import (  # noqa: F401
    QUERY_ORDER_AUTO,
    QUERY_ORDER_INPUT,
    QUERY_ORDER_MORTON,
    REFIT_ATOMIC,
    REFIT_AUTO,
    REFIT_WAVEFRONT,
    get_refit_mode,
    set_refit_mode,
    MeshQueryPoint,
    get_query_order,
    get_ray_order,
    set_query_order,
    set_ray_order,
    MeshQueryRay,
    mesh_query_point,
    mesh_query_point_no_sign,
    mesh_query_point_sign_parity,
    mesh_query_point_sign_normal,
    mesh_average_edge_length,
    mesh_eval_position,
    mesh_eval_velocity,
    mesh_eval_face_normal,
    mesh_query_furthest_point_no_sign,
    mesh_query_ray,
    mesh_query_ray_anyhit,
    mesh_query_ray_count_intersections,
    query_stats,
)

from .bvh_queries import (  # noqa: F401,E402
    BvhQueryResult,
    bvh_get_group_root,
    bvh_query_aabb,
    bvh_query_ray,
    bvh_query_sphere,
    mesh_query_sphere,
    bvh_query_capsule,
    mesh_query_aabb,
)

from .capture import (  # noqa: F401,E402
    Graph,
    ScopedCapture,
    ScopedStream,
    Stream,
    capture_begin,
    capture_end,
    capture_launch,
)

__version__ = "0.1.0"


def is_cuda_available() -> bool:
    from . import _lib

    try:
        return _lib.core().wp_cuda_device_get_count() > 0
    except RuntimeError:
        return False


def get_error_string() -> str:
    from . import _lib

    return _lib.error_string()
