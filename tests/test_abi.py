"""CPU tests of the drop-in boundary: the shared library loads without a GPU, exports every symbol
include/warp_b200.h declares, and the by-value structs have the reference's sizes."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "warp_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"WP_B200_API\s+[\w\s\*]+?\b(wp_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from warp_b200 import build

    build.build()  # no-op when up to date; nvcc cross-compiles without a GPU
    return ctypes.CDLL(build.LIB)


def test_header_declares_the_reference_entry_points():
    syms = declared_symbols()
    for name in (
        "wp_bvh_create_device", "wp_bvh_destroy_device", "wp_bvh_refit_device", "wp_bvh_rebuild_device",
        "wp_mesh_create_device", "wp_mesh_destroy_device", "wp_mesh_refit_device", "wp_mesh_set_points_device",
        "wp_mesh_set_velocities_device", "wp_get_error_string",
    ):  # warp/native/warp.h:93-140
        assert name in syms
    assert len(syms) > 40


def test_library_exports_every_declared_symbol(lib):
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_bindings_cover_the_header(lib):
    from warp_b200 import _lib

    assert sorted(_lib.SIGNATURES) == declared_symbols()
    _lib.load()  # binds argtypes / restypes for all of them


def test_struct_sizes_match_reference_layout():
    from warp_b200 import _lib

    assert ctypes.sizeof(_lib.array_t) == 56  # wp::array_t, warp/native/array.h:173-277


def test_no_gpu_means_loud_failure_not_fallback(lib):
    import warp_b200 as wp

    if wp.is_cuda_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        wp.get_device("cuda:0")
    with pytest.raises(RuntimeError):
        wp.array([[0, 0, 0]], dtype=wp.vec3, device="cpu")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "warp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "lbvh_oracle" not in src, f
