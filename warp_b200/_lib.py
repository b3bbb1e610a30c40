"""ctypes loader for ``libwarp_b200.so`` -- the only native dependency of the package.

Mirrors the binding block of the reference (``warp/_src/context.py:7000-7070``): same symbols,
same argtypes / restypes, plus the batched-query and runtime-slice entry points declared in
``include/warp_b200.h``.  There is no CPU fallback: if the library is missing or no CUDA device is
visible, every entry into the path raises.
"""

from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libwarp_b200.so")


class array_t(ctypes.Structure):
    """``wp::array_t`` passed by value (``warp/_src/types.py:2417-2425``)."""

    _fields_ = (
        ("data", ctypes.c_uint64),
        ("grad", ctypes.c_uint64),
        ("shape", ctypes.c_int32 * 4),
        ("strides", ctypes.c_int32 * 4),
        ("ndim", ctypes.c_uint16),
        ("flags", ctypes.c_uint16),
    )


class bvh_info_t(ctypes.Structure):
    _fields_ = (
        ("num_items", ctypes.c_int),
        ("leaf_size", ctypes.c_int),
        ("max_nodes", ctypes.c_int),
        ("root", ctypes.c_int),
        ("height", ctypes.c_int),
        ("deep", ctypes.c_int),
        ("key_bits", ctypes.c_int),
        ("total_lower", ctypes.c_float * 3),
        ("total_upper", ctypes.c_float * 3),
        ("inv_edges", ctypes.c_float * 3),
    )


# every symbol include/warp_b200.h declares: name -> (restype, argtypes)
_vp, _i, _u64, _i64, _f, _sz = (ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_int64, ctypes.c_float,
                                 ctypes.c_size_t)  # fmt: skip
SIGNATURES = {
    # part 1: drop-in
    "wp_bvh_create_device": (_u64, [_vp, _vp, _vp, _i, _i, _vp, _i]),
    "wp_bvh_destroy_device": (None, [_u64]),
    "wp_bvh_refit_device": (None, [_u64]),
    "wp_bvh_rebuild_device": (None, [_u64]),
    "wp_mesh_create_device": (_u64, [_vp, array_t, array_t, array_t, _i, _i, _i, _i, _vp, _i]),
    "wp_mesh_destroy_device": (None, [_u64]),
    "wp_mesh_refit_device": (_i, [_u64]),
    "wp_mesh_set_points_device": (_i, [_u64, array_t]),
    "wp_mesh_set_velocities_device": (None, [_u64, array_t]),
    "wp_get_error_string": (ctypes.c_char_p, []),
    # part 2: runtime slice
    "wp_init": (_i, [ctypes.c_char_p]),
    "wp_is_cuda_enabled": (_i, []),
    "wp_cuda_device_get_count": (_i, []),
    "wp_cuda_device_get_primary_context": (_vp, [_i]),
    "wp_cuda_context_get_current": (_vp, []),
    "wp_cuda_context_set_current": (None, [_vp]),
    "wp_cuda_context_synchronize": (None, [_vp]),
    "wp_cuda_context_get_stream": (_vp, [_vp]),
    "wp_cuda_context_set_stream": (None, [_vp, _vp, _i]),
    "wp_cuda_stream_create": (_vp, [_vp, _i]),
    "wp_cuda_stream_destroy": (None, [_vp, _vp]),
    "wp_cuda_stream_synchronize": (None, [_vp]),
    "wp_cuda_event_create": (_vp, [_vp, ctypes.c_uint]),
    "wp_cuda_event_destroy": (None, [_vp]),
    "wp_cuda_event_record": (None, [_vp, _vp, _i]),
    "wp_cuda_event_synchronize": (None, [_vp]),
    "wp_cuda_event_elapsed_time": (_f, [_vp, _vp]),
    "wp_cuda_graph_begin_capture": (_i, [_vp, _vp, _i, _i]),
    "wp_cuda_graph_end_capture": (_i, [_vp, _vp, ctypes.POINTER(ctypes.c_void_p)]),
    "wp_cuda_graph_create_exec": (_i, [_vp, _vp, _vp, ctypes.POINTER(ctypes.c_void_p)]),
    "wp_cuda_graph_launch": (_i, [_vp, _vp]),
    "wp_cuda_graph_destroy": (_i, [_vp, _vp]),
    "wp_cuda_graph_exec_destroy": (_i, [_vp, _vp]),
    "wp_alloc_device": (_vp, [_vp, _sz, ctypes.c_char_p]),
    "wp_free_device": (None, [_vp, _vp]),
    "wp_alloc_pinned": (_vp, [_sz, ctypes.c_char_p]),
    "wp_free_pinned": (None, [_vp]),
    "wp_memcpy_h2d": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "wp_memcpy_d2h": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "wp_memcpy_d2d": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "wp_memset_device": (_i, [_vp, _vp, _i, _sz, _vp]),
    "wp_b200_device_attr": (_i, [_i, ctypes.c_char_p, ctypes.POINTER(ctypes.c_longlong)]),
    "wp_b200_pointer_device": (_i, [_vp, ctypes.POINTER(ctypes.c_longlong)]),
    "wp_b200_device_name": (_i, [_i, ctypes.c_char_p, _i]),
    # part 3: batched queries
    "wp_b200_mesh_query_point_no_sign": (_i, [_u64, _vp, _i64, _f, _vp, _vp, _vp, _vp]),
    "wp_b200_mesh_query_point": (_i, [_u64, _vp, _i64, _f, _vp, _vp, _vp, _vp, _vp]),
    "wp_b200_mesh_query_point_sign_parity": (_i, [_u64, _vp, _i64, _f, _i, _f, _vp, _vp, _vp, _vp, _vp]),
    "wp_b200_mesh_query_point_sign_normal": (_i, [_u64, _vp, _i64, _f, _f, _vp, _vp, _vp, _vp, _vp]),
    "wp_b200_mesh_average_edge_length": (_i, [_u64, _vp]),
    "wp_b200_mesh_query_ray": (_i, [_u64, _vp, _vp, _i64, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "wp_b200_mesh_query_ray_anyhit": (_i, [_u64, _vp, _vp, _i64, _f, _vp, _vp]),
    "wp_b200_mesh_query_ray_count_intersections": (_i, [_u64, _vp, _vp, _i64, _vp, _vp]),
    "wp_b200_mesh_eval_position": (_i, [_u64, _vp, _vp, _vp, _i64, _vp]),
    "wp_b200_mesh_eval_velocity": (_i, [_u64, _vp, _vp, _vp, _i64, _vp]),
    "wp_b200_mesh_eval_face_normal": (_i, [_u64, _vp, _i64, _vp]),
    "wp_b200_mesh_eval_face_normal_masked": (_i, [_u64, _vp, _vp, _i64, _vp]),
    "wp_b200_mesh_query_furthest_point_no_sign": (_i, [_u64, _vp, _i64, _f, _vp, _vp, _vp, _vp]),
    "wp_b200_mesh_query_point_no_sign_host":(_i, [_u64, _vp, _i64, _f, _vp, _vp, _vp, _vp]),
    "wp_b200_mesh_query_point_host": (_i, [_u64, _vp, _i64, _f, _vp, _vp, _vp, _vp, _vp]),
    "wp_b200_mesh_query_ray_host": (_i, [_u64, _vp, _vp, _i64, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "wp_b200_set_morton_bits": (None, [_i]),
    "wp_b200_get_morton_bits": (_i, []),
    "wp_b200_set_query_order": (None, [_i]),
    "wp_b200_get_query_order": (_i, []),
    "wp_b200_set_ray_order": (None, [_i]),
    "wp_b200_get_ray_order": (_i, []),
    "wp_b200_kernel_timing_enable": (None, [_i]),
    "wp_b200_kernel_timing_read": (None, [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)]),
    "wp_b200_query_stats_enable": (None, [_i]),
    "wp_b200_query_stats_read": (None, [ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_ulonglong)]),
    "wp_b200_bvh_query_aabb_count": (_i, [_u64, _vp, _vp, _vp, _i64, _vp]),
    "wp_b200_bvh_query_aabb_fill": (_i, [_u64, _vp, _vp, _vp, _i64, _vp, _vp]),
    "wp_b200_bvh_query_ray_count": (_i, [_u64, _vp, _vp, _vp, _i64, _f, _vp]),
    "wp_b200_bvh_query_ray_fill": (_i, [_u64, _vp, _vp, _vp, _i64, _f, _vp, _vp]),
    "wp_b200_mesh_query_sphere_count": (_i, [_u64, _vp, _vp, _i64, _vp]),
    "wp_b200_mesh_query_sphere_fill": (_i, [_u64, _vp, _vp, _i64, _vp, _vp]),
    "wp_b200_bvh_query_sphere_count": (_i, [_u64, _vp, _vp, _vp, _i64, _vp]),
    "wp_b200_bvh_query_sphere_fill": (_i, [_u64, _vp, _vp, _vp, _i64, _vp, _vp]),
    "wp_b200_bvh_query_capsule_count": (_i, [_u64, _vp, _vp, _vp, _vp, _i64, _f, _vp]),
    "wp_b200_bvh_query_capsule_fill": (_i, [_u64, _vp, _vp, _vp, _vp, _i64, _f, _vp, _vp]),
    "wp_b200_bvh_get_group_root": (_i, [_u64, _vp, _i64, _vp]),
    "wp_b200_mesh_query_aabb_count": (_i, [_u64, _vp, _vp, _i64, _vp]),
    "wp_b200_mesh_query_aabb_fill": (_i, [_u64, _vp, _vp, _i64, _vp, _vp]),
    "wp_b200_exclusive_scan_i32": (_i, [_vp, _vp, _i64]),
    "wp_b200_set_refit_mode": (None, [_i]),
    "wp_b200_get_refit_mode": (_i, []),
    "wp_b200_mesh_rebuild_device": (_i, [_u64]),
    "wp_b200_bvh_create_device_ex": (_u64, [_vp, _vp, _vp, _i, _i, _vp, _i, _i]),
    "wp_b200_mesh_create_device_ex": (_u64, [_vp, array_t, array_t, array_t, _i, _i, _i, _i, _vp, _i, _i]),
    "wp_b200_bvh_set_option": (_i, [_u64, ctypes.c_char_p, _i]),
    "wp_b200_bvh_get_option": (_i, [_u64, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int)]),
    "wp_b200_host_build_order": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "wp_b200_set_experiment": (_i, [ctypes.c_char_p, _i]),
    "wp_b200_set_auto_reference_layout": (None, [_i]),
    "wp_b200_get_auto_reference_layout": (_i, []),
    "wp_b200_bvh_info": (_i, [_u64, ctypes.POINTER(bvh_info_t)]),
    "wp_b200_bvh_sync_reference_layout": (_i, [_u64]),
    "wp_b200_bvh_download": (_i, [_u64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "wp_b200_experiment_parallel_topology": (ctypes.c_float, [_u64, _vp, _i]),
    "wp_b200_gen_box_queries": (_i, [_vp, _i64, _i64, _u64, _vp, _vp]),
    "wp_b200_gen_cloth_points": (_i, [_vp, _i, _vp, _i]),
    "wp_b200_gen_cloth_queries": (_i, [_vp, _i64, _i, _vp, _f]),
    "wp_b200_counter_add": (_i, [_vp, _i]),
    # multi-GPU
    "wp_b200_nccl_load": (_i, [ctypes.c_char_p]),
    "wp_b200_nccl_unique_id": (_i, [_vp]),
    "wp_b200_nccl_init": (_i, [_vp, _i, _i]),
    "wp_b200_nccl_allgather": (_i, [_vp, _vp, _sz]),
    "wp_b200_nccl_allgather_part": (_i, [_vp, _vp, _sz, _sz, _sz]),
    "wp_b200_nccl_allgather_multi": (_i, [_vp, _vp, _vp, _i, _i]),
    "wp_b200_ipc_get_handle": (_i, [_vp, _vp]),
    "wp_b200_ipc_open_handle": (_vp, [_vp]),
    "wp_b200_ipc_close_handle": (None, [_vp, _vp]),
    "wp_b200_p2p_allgather_multi": (_i, [_vp, _vp, _vp, _vp, _i, _i]),
    "wp_b200_nccl_comm_stream": (_vp, []),
    "wp_b200_nccl_mark": (_i, [_i]),
    "wp_b200_nccl_wait_mark": (_i, [_i]),
    "wp_b200_nccl_fork": (_i, []),
    "wp_b200_nccl_join": (_i, []),
    "wp_b200_nccl_allreduce_max_f32": (_i, [_vp, _sz]),
    "wp_b200_nccl_barrier": (_i, []),
    "wp_b200_nccl_destroy": (None, []),
}

_core = None


def load(path: str | None = None) -> ctypes.CDLL:
    """dlopen the native library and bind every declared symbol; raises if it is not built."""
    global _core
    if _core is not None and path is None:
        return _core
    p = path or os.environ.get("WARP_B200_LIB") or LIB_PATH  # WARP_B200_LIB: A/B builds (scripts/build_variants.sh)
    if not os.path.exists(p):
        raise RuntimeError(
            f"warp_b200: native library {p} is not built; run `python -m warp_b200.build` "
            "(needs nvcc).  There is no CPU fallback for this path."
        )
    core = ctypes.CDLL(p)  # ctypes.CDLL releases the GIL around calls, like the reference (context.py:8452-8454)
    variant = p != LIB_PATH  # an A/B build of an older revision may lack newer entry points; the product library may not
    for name, (restype, argtypes) in SIGNATURES.items():
        try:
            fn = getattr(core, name)  # AttributeError here == the .so does not export a declared symbol
        except AttributeError:
            if variant:
                continue
            raise
        fn.restype = restype
        fn.argtypes = argtypes
    if path is None:
        _core = core
    return core


def core() -> ctypes.CDLL:
    return load()


def error_string() -> str:
    return core().wp_get_error_string().decode("utf-8", "replace")


def require_cuda() -> int:
    """Number of CUDA devices; raises (never falls back) when there is none."""
    n = core().wp_cuda_device_get_count()
    if n <= 0:
        raise RuntimeError("warp_b200: no CUDA device is visible and this path has no CPU fallback")
    return n
