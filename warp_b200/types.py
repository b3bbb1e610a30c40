"""Host-side mirror of the reference's Python surface for the mesh/BVH path.

``Mesh`` / ``Bvh`` / ``BvhConstructor`` follow ``warp/_src/types.py:5774-6310`` -- same constructor
arguments and defaults, same validation order and error messages, same ``id`` / ``device`` /
``refit()`` / ``rebuild()`` / ``points`` setter semantics -- but there is no Warp type system
underneath: ``array`` is a minimal device-buffer holder (pointer, shape, dtype) with numpy
round-trips, enough to carry points / indices / bounds / results.
"""

from __future__ import annotations

import builtins
import ctypes
import enum
import warnings

import numpy as np

from . import _lib

# ------------------------------------------------------------------------------------------------
# scalar / vector type markers (wp.vec3, wp.int32, ...)
# ------------------------------------------------------------------------------------------------


class _DType:
    def __init__(self, name, np_dtype, length=1):
        self.name, self.np_dtype, self.length = name, np.dtype(np_dtype), length
        self.itemsize = self.np_dtype.itemsize * length

    def __repr__(self):
        return f"wp.{self.name}"


vec3 = _DType("vec3", np.float32, 3)
float32 = _DType("float32", np.float32)
int32 = _DType("int32", np.int32)
uint8 = _DType("uint8", np.uint8)
uint32 = _DType("uint32", np.uint32)
uint64 = _DType("uint64", np.uint64)
bool_ = uint8  # query `result` arrays are one byte per query

_BY_NP = {np.dtype(np.float32): float32, np.dtype(np.int32): int32, np.dtype(np.uint8): uint8,
          np.dtype(np.uint32): uint32, np.dtype(np.uint64): uint64, np.dtype(np.bool_): uint8}  # fmt: skip


def _as_dtype(dtype):
    if isinstance(dtype, _DType):
        return dtype
    if dtype is int:
        return int32
    if dtype is float:
        return float32
    try:
        return _BY_NP[np.dtype(dtype)]
    except (KeyError, TypeError):
        raise RuntimeError(f"unsupported array dtype {dtype!r}") from None


# ------------------------------------------------------------------------------------------------
# devices
# ------------------------------------------------------------------------------------------------


class Device:
    """``cuda:N`` only.  ``cpu`` exists as a value so the reference's device checks can be mirrored,
    but nothing can be allocated or built on it (no CPU fallback)."""

    def __init__(self, alias: str):
        self.alias = alias
        self.is_cpu = alias == "cpu"
        self.is_cuda = not self.is_cpu
        self.ordinal = -1 if self.is_cpu else int(alias.split(":")[1])
        self._context = None

    @property
    def context(self):
        """The device's primary ``CUcontext`` (what the reference's ``Device.context`` holds and passes to the native
        creators, ``warp/_src/types.py:5947-5950``), resolved on first use."""
        if self.is_cpu:
            return None
        if self._context is None:
            self._context = ctypes.c_void_p(_lib.core().wp_cuda_device_get_primary_context(self.ordinal))
        return self._context

    def __eq__(self, other):
        return isinstance(other, Device) and other.alias == self.alias

    def __hash__(self):
        return hash(self.alias)

    def __repr__(self):
        return f"'{self.alias}'"

    __str__ = __repr__


_devices: dict[str, Device] = {}


def get_device(ident=None) -> Device:
    if isinstance(ident, Device):
        return ident
    if ident is None or ident == "cuda":
        ident = "cuda:0"
    if ident not in _devices:
        if ident != "cpu":
            if not (isinstance(ident, str) and ident.startswith("cuda:") and ident[5:].isdigit()):
                raise RuntimeError(f"Invalid device identifier: {ident}")
            n = _lib.require_cuda()
            if int(ident[5:]) >= n:
                raise RuntimeError(f"Invalid device ordinal {ident}; {n} CUDA device(s) visible")
        _devices[ident] = Device(ident)
    return _devices[ident]


def synchronize_device(device=None):
    _lib.core().wp_cuda_context_synchronize(get_device(device).context)


synchronize = synchronize_device

# ------------------------------------------------------------------------------------------------
# device array
# ------------------------------------------------------------------------------------------------


class array:
    """Contiguous device buffer: ``array(data, dtype=wp.vec3, device="cuda:0")``.

    ``data`` may be a numpy array / nested sequence (copied to the device), ``None`` together with
    ``shape`` (uninitialised), or an object exposing ``__cuda_array_interface__`` (wrapped, zero
    copy; the owner must outlive the array).  ``ptr=`` wraps a raw device pointer.
    """

    def __init__(self, data=None, dtype=None, shape=None, device=None, ptr=None, owner=None):
        self.owner = owner
        self._owns = False
        self.ptr = None
        if data is not None and hasattr(data, "__cuda_array_interface__") and not isinstance(data, np.ndarray):
            cai = data.__cuda_array_interface__
            if cai.get("strides") is not None:
                raise RuntimeError("only contiguous __cuda_array_interface__ objects can be wrapped")
            npdt = np.dtype(cai["typestr"])
            shp = tuple(cai["shape"])
            dt = _as_dtype(dtype) if dtype is not None else _as_dtype(npdt)
            if dt.length > 1:
                if shp[-1] != dt.length or npdt != dt.np_dtype:
                    raise RuntimeError(f"cannot view shape {shp} {npdt} as {dt}")
                shp = shp[:-1]
            self.dtype, self.shape = dt, shp
            self.ptr = int(cai["data"][0])
            if device is None and self.ptr:
                # the buffer says where it lives; labelling a cuda:1 pointer cuda:0 would send kernels to the wrong device
                ordinal = ctypes.c_longlong(-1)
                if _lib.core().wp_b200_pointer_device(ctypes.c_void_p(self.ptr), ctypes.byref(ordinal)) and ordinal.value >= 0:
                    device = f"cuda:{ordinal.value}"
            self.device = get_device(device)
            self.owner = data
            return
        self.device = get_device(device)
        if self.device.is_cpu:
            raise RuntimeError("warp_b200 arrays live on CUDA devices only (no CPU fallback for this path)")
        if ptr is not None:
            self.dtype = _as_dtype(dtype)
            self.shape = (shape,) if isinstance(shape, int) else tuple(shape)
            self.ptr = int(ptr)
            return
        if data is not None:
            dt = _as_dtype(dtype) if dtype is not None else None
            host = np.asarray(data, dtype=dt.np_dtype if dt else None)
            if dt is None:
                if host.dtype == np.float64:
                    host = host.astype(np.float32)
                elif host.dtype == np.int64:
                    host = host.astype(np.int32)
                dt = _as_dtype(host.dtype)
            if dt.length > 1:
                if host.ndim < 1 or host.shape[-1] != dt.length:
                    if host.size % dt.length:
                        raise RuntimeError(f"cannot interpret data of shape {host.shape} as {dt}")
                    host = host.reshape(-1, dt.length)
                shp = host.shape[:-1]
            else:
                shp = host.shape
            self.dtype, self.shape = dt, tuple(shp)
            self._alloc()
            self.assign(host)
            return
        self.dtype = _as_dtype(dtype if dtype is not None else float32)
        if shape is None:
            shape = 0
        self.shape = (shape,) if isinstance(shape, (int, np.integer)) else tuple(int(s) for s in shape)
        self._alloc()

    # -- storage
    def _alloc(self):
        nbytes = self.nbytes
        self._owns = True
        if nbytes == 0:
            self.ptr = None
            return
        p = _lib.core().wp_alloc_device(self.device.context, nbytes, b"(warp_b200:array)")
        if not p:
            raise RuntimeError(f"Failed to allocate {nbytes} bytes on {self.device}: {_lib.error_string()}")
        self.ptr = int(p)

    def __del__(self):
        try:
            if self._owns and self.ptr:
                _lib.core().wp_free_device(self.device.context, ctypes.c_void_p(self.ptr))
                self.ptr = None
        except (TypeError, AttributeError):
            pass

    # -- metadata
    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64)) if len(self.shape) else 1

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def nbytes(self):
        return self.size * self.dtype.itemsize

    @property
    def is_contiguous(self):
        return True

    def __len__(self):
        return self.shape[0] if self.shape else 0

    def __bool__(self):
        return True

    def __ctype__(self) -> _lib.array_t:
        a = _lib.array_t()
        a.data = self.ptr or 0
        a.grad = 0
        stride = self.dtype.itemsize
        for k in range(self.ndim - 1, -1, -1):
            a.shape[k] = self.shape[k]
            a.strides[k] = stride
            stride *= self.shape[k]
        a.ndim = self.ndim
        a.flags = 0
        return a

    @property
    def __cuda_array_interface__(self):
        shp = self.shape + ((self.dtype.length,) if self.dtype.length > 1 else ())
        return {"shape": shp, "typestr": self.dtype.np_dtype.str, "data": (self.ptr or 0, False), "version": 2}

    # -- transfers (synchronous with respect to the device's current stream)
    def _stream(self):
        return _lib.core().wp_cuda_context_get_stream(self.device.context)

    def assign(self, src):
        """Overwrite the contents from a numpy array / sequence (H2D) or another ``array`` (D2D)."""
        c = _lib.core()
        if isinstance(src, array):
            if src.nbytes != self.nbytes:
                raise RuntimeError("array.assign: size mismatch")
            if self.nbytes:
                c.wp_memcpy_d2d(self.device.context, self.ptr, src.ptr, self.nbytes, self._stream())
            return self
        host = np.ascontiguousarray(src, dtype=self.dtype.np_dtype)
        if host.nbytes != self.nbytes:
            raise RuntimeError(f"array.assign: source has {host.nbytes} bytes, array has {self.nbytes}")
        if self.nbytes:
            if not c.wp_memcpy_h2d(self.device.context, self.ptr, host.ctypes.data, self.nbytes, self._stream()):
                raise RuntimeError(_lib.error_string())
            c.wp_cuda_stream_synchronize(self._stream())  # `host` may be a temporary
        return self

    def numpy(self) -> np.ndarray:
        shp = self.shape + ((self.dtype.length,) if self.dtype.length > 1 else ())
        out = np.empty(shp, dtype=self.dtype.np_dtype)
        if self.nbytes:
            c = _lib.core()
            if not c.wp_memcpy_d2h(self.device.context, out.ctypes.data, self.ptr, self.nbytes, self._stream()):
                raise RuntimeError(_lib.error_string())
            c.wp_cuda_stream_synchronize(self._stream())
        return out

    def zero_(self):
        if self.nbytes:
            _lib.core().wp_memset_device(self.device.context, self.ptr, 0, self.nbytes, self._stream())
        return self

    def __repr__(self):
        return f"array(shape={self.shape}, dtype={self.dtype}, device={self.device})"


def empty(shape, dtype=float32, device=None) -> array:
    return array(None, dtype=dtype, shape=shape, device=device)


def zeros(shape, dtype=float32, device=None) -> array:
    return empty(shape, dtype, device).zero_()


def from_numpy(arr, dtype=None, device=None) -> array:
    return array(arr, dtype=dtype, device=device)


# ------------------------------------------------------------------------------------------------
# Bvh / Mesh
# ------------------------------------------------------------------------------------------------


class BvhConstructor(enum.IntEnum):
    """BVH construction algorithm selection (``warp/_src/types.py:5774-5793``)."""

    SAH = 0
    MEDIAN = 1
    LBVH = 2
    CUBQL = -1

    @classmethod
    def from_str(cls, value: str) -> "BvhConstructor":
        try:
            return cls[value.upper()]
        except KeyError:
            raise ValueError(
                f"Unknown BVH constructor '{value}', expected one of: {', '.join(m.name.lower() for m in cls)}"
            ) from None


def _check_morton_bits(bits, groups):
    """Extension (not in the reference API): ``morton_bits=63`` builds the tree on a 2 097 152^3 Morton grid
    instead of the reference's 1024^3 -- a quality option for meshes with far more than 2^20 occupied cells.
    Only ``morton_bits=30`` (default) reproduces the reference tree bit for bit.  Passed to the native creator as an
    argument (``wp_b200_*_create_device_ex``), so it is per object and thread-safe."""
    if bits not in (30, 63):
        raise ValueError(f"morton_bits must be 30 or 63, current value: {bits}")
    if bits == 63 and groups is not None:
        raise RuntimeError("morton_bits=63 cannot be combined with groups (the key holds group << 32 | 30-bit code)")
    return int(bits)


class _Options:
    """Per-object native options (``wp_b200_bvh_set_option``): ``refit_mode``, ``query_order``, ``ray_order``,
    ``auto_reference_layout``; -1 = follow the process-wide default."""

    def set_option(self, name: str, value: int) -> None:
        if not _lib.core().wp_b200_bvh_set_option(self.id, name.encode(), int(value)):
            raise RuntimeError(_lib.error_string())

    def get_option(self, name: str) -> int:
        v = ctypes.c_int(0)
        if not _lib.core().wp_b200_bvh_get_option(self.id, name.encode(), ctypes.byref(v)):
            raise RuntimeError(_lib.error_string())
        return v.value


def _void_p(arr):
    return ctypes.c_void_p(arr.ptr) if arr is not None and arr.ptr else ctypes.c_void_p(0)


class Bvh(_Options):
    """Bounding volume hierarchy over AABBs -- ``warp/_src/types.py:5796-6078``.

    Only GPU trees exist here; ``constructor=None`` or ``"lbvh"`` builds with the B200 LBVH builder,
    the host constructors (``"sah"``, ``"median"``) and ``"cubql"`` are rejected by the native library.
    """

    def __new__(cls, *args, **kwargs):
        instance = super().__new__(cls)
        instance.id = None
        return instance

    def __init__(self, lowers, uppers, constructor=None, groups=None, leaf_size: int = 1, morton_bits: int = 30):
        if len(lowers) != len(uppers):
            raise RuntimeError("The same number of lower and upper bounds must be provided")
        if lowers.device != uppers.device:
            raise RuntimeError("Lower and upper bounds must live on the same device")
        if lowers.dtype != vec3 or not lowers.is_contiguous:
            raise RuntimeError("lowers should be a contiguous array of type wp.vec3")
        if uppers.dtype != vec3 or not uppers.is_contiguous:
            raise RuntimeError("uppers should be a contiguous array of type wp.vec3")
        if groups is not None:
            if groups.dtype != int32 or not groups.is_contiguous:
                raise RuntimeError("groups should be a contiguous array of type wp.int32")
            if groups.device != lowers.device:
                raise RuntimeError("groups must live on the same device as lowers/uppers")
            if len(groups) != len(lowers):
                raise RuntimeError("groups must have the same length as lowers/uppers")

        self.device = lowers.device
        self.lowers = lowers
        self.uppers = uppers
        self.groups = groups

        if constructor is None:
            constructor = BvhConstructor.LBVH  # GPU tree default (types.py:5924-5928)
        if not isinstance(constructor, BvhConstructor):
            constructor = BvhConstructor.from_str(constructor)
        if constructor == BvhConstructor.CUBQL and groups is not None:
            raise RuntimeError("Grouped BVHs are not supported with constructor='cubql'")
        if leaf_size < 1:
            raise ValueError(f"leaf_size must be greater than or equal to 1, current value: {leaf_size}")
        if self.device.is_cpu:
            raise RuntimeError("warp_b200.Bvh: CPU trees are not available (no CPU fallback for this path)")

        self.id = _lib.core().wp_b200_bvh_create_device_ex(
            self.device.context, _void_p(lowers), _void_p(uppers), len(lowers), int(constructor), _void_p(groups),
            leaf_size, _check_morton_bits(morton_bits, groups),
        )  # fmt: skip
        self._constructor = constructor
        self.leaf_size = leaf_size
        if not self.id:
            raise RuntimeError(f"Failed to create BVH: {_lib.error_string()}")

    def __del__(self):
        if not self.id:
            return
        try:
            _lib.core().wp_bvh_destroy_device(self.id)
        except (TypeError, AttributeError):
            pass

    def refit(self):
        """Refit the BVH after ``lowers`` / ``uppers`` were modified in place."""
        _lib.core().wp_bvh_refit_device(self.id)

    def rebuild(self, constructor=None):
        """Rebuild the hierarchy in place from the current bounds (no allocation)."""
        if constructor is None:
            constructor = BvhConstructor.LBVH
        if not isinstance(constructor, BvhConstructor):
            constructor = BvhConstructor.from_str(constructor)
        if constructor == BvhConstructor.CUBQL:
            raise ValueError("Cannot rebuild a non-cuBQL BVH with constructor='cubql'; create a new BVH instead")
        if constructor != BvhConstructor.LBVH:
            warnings.warn(
                "In-place rebuild method on the CUDA device only supports LBVH constructor. "
                "Falling back to LBVH constructor.",
                stacklevel=2,
            )
        _lib.core().wp_bvh_rebuild_device(self.id)
        self._constructor = BvhConstructor.LBVH

    # parity / debugging helper (not in the reference API)
    def download_tree(self):
        return _download_tree(self.id, len(self.lowers))


class Mesh(_Options):
    """Triangle mesh with an LBVH for closest-point and ray queries -- ``warp/_src/types.py:6081-6310``."""

    def __new__(cls, *args, **kwargs):
        instance = super().__new__(cls)
        instance.id = None
        return instance

    def __init__(
        self,
        points,
        indices,
        velocities=None,
        support_winding_number: builtins.bool = False,
        bvh_constructor=None,
        bvh_leaf_size: int | None = None,
        groups=None,
        morton_bits: int = 30,
    ):
        if points.device != indices.device:
            raise RuntimeError("Mesh points and indices must live on the same device")
        if points.dtype != vec3 or not points.is_contiguous:
            raise RuntimeError("Mesh points should be a contiguous array of type wp.vec3")
        if velocities and (velocities.dtype != vec3 or not velocities.is_contiguous):
            raise RuntimeError("Mesh velocities should be a contiguous array of type wp.vec3")
        if indices.dtype != int32 or not indices.is_contiguous:
            raise RuntimeError("Mesh indices should be a contiguous array of type wp.int32")
        if indices.ndim > 1:
            raise RuntimeError("Mesh indices should be a flattened 1d array of indices")
        if groups is not None:
            if groups.dtype != int32 or not groups.is_contiguous:
                raise RuntimeError("groups should be a contiguous array of type wp.int32")
            if groups.device != points.device:
                raise RuntimeError("groups must live on the same device as points")
            if len(groups) != len(indices) // 3:
                raise RuntimeError("groups must have the same length as indices / 3")

        self.device = points.device
        self._points = points
        self._velocities = velocities
        self.indices = indices
        self.groups = groups

        if bvh_constructor is None:
            bvh_constructor = BvhConstructor.LBVH
        if not isinstance(bvh_constructor, BvhConstructor):
            bvh_constructor = BvhConstructor.from_str(bvh_constructor)
        if bvh_constructor == BvhConstructor.CUBQL:
            if groups is not None:
                raise RuntimeError("Grouped mesh queries are not supported with bvh_constructor='cubql'")
            if support_winding_number:
                raise RuntimeError("support_winding_number=True is not supported with bvh_constructor='cubql'")
            if bvh_leaf_size is None:
                bvh_leaf_size = 0
            elif bvh_leaf_size < 0:
                raise ValueError(f"bvh_leaf_size must be greater than or equal to 0, current value: {bvh_leaf_size}")
        else:
            if bvh_leaf_size is None:
                bvh_leaf_size = 4
            elif bvh_leaf_size < 1:
                raise ValueError(f"bvh_leaf_size must be greater than or equal to 1, current value: {bvh_leaf_size}")
        if self.device.is_cpu:
            raise RuntimeError("warp_b200.Mesh: CPU meshes are not available (no CPU fallback for this path)")

        self.bvh_leaf_size = bvh_leaf_size
        self.id = _lib.core().wp_b200_mesh_create_device_ex(
            self.device.context, points.__ctype__(),
            velocities.__ctype__() if velocities else _lib.array_t(), indices.__ctype__(),
            len(points), int(indices.size // 3), int(support_winding_number), int(bvh_constructor),
            _void_p(groups), bvh_leaf_size, _check_morton_bits(morton_bits, groups),
        )  # fmt: skip
        if not self.id:
            raise RuntimeError(f"Failed to create mesh: {_lib.error_string()}")

    def __del__(self):
        if not self.id:
            return
        try:
            _lib.core().wp_mesh_destroy_device(self.id)
        except (TypeError, AttributeError):
            pass

    def refit(self):
        """Refit the BVH to ``points`` after they were modified in place."""
        if not _lib.core().wp_mesh_refit_device(self.id):
            raise RuntimeError(f"Failed to refit mesh: {_lib.error_string()}")

    def rebuild(self):
        """Rebuild the LBVH in place from the current ``points`` (extension: the reference offers
        ``rebuild`` on ``Bvh`` only).  No allocation; buffers and ``id`` are unchanged."""
        if not _lib.core().wp_b200_mesh_rebuild_device(self.id):
            raise RuntimeError(f"Failed to rebuild mesh: {_lib.error_string()}")

    @property
    def points(self):
        return self._points

    @points.setter
    def points(self, points_new):
        if points_new.device != self._points.device:
            raise RuntimeError(
                "The new points and the original points must live on the same device, the "
                f"new points are on {points_new.device} while the old points are on {self._points.device}."
            )
        if points_new.ndim != 1 or points_new.shape[0] != self._points.shape[0]:
            raise RuntimeError(
                "The new points and the original points must have the same shape, the "
                f"new points' shape is {points_new.shape}, while the old points' shape is {self._points.shape}."
            )
        self._points = points_new
        if not _lib.core().wp_mesh_set_points_device(self.id, points_new.__ctype__()):
            raise RuntimeError(f"Failed to set mesh points: {_lib.error_string()}")

    @property
    def velocities(self):
        return self._velocities

    @velocities.setter
    def velocities(self, velocities_new):
        if velocities_new.device != self._velocities.device:
            raise RuntimeError(
                "The new points and the original points must live on the same device, the "
                f"new points are on {velocities_new.device} while the old points are on {self._velocities.device}."
            )
        if velocities_new.ndim != 1 or velocities_new.shape[0] != self._velocities.shape[0]:
            raise RuntimeError(
                "The new points and the original points must have the same shape, the "
                f"new points' shape is {velocities_new.shape}, while the old points' shape is {self._velocities.shape}."
            )
        self._velocities = velocities_new
        _lib.core().wp_mesh_set_velocities_device(self.id, velocities_new.__ctype__())

    # parity / debugging helper (not in the reference API)
    def download_tree(self):
        return _download_tree(self.id, int(self.indices.size // 3))

    def download_primitive_indices(self):
        """``bvh.primitive_indices`` (sorted position -> face); also available for sah / median trees."""
        n = int(self.indices.size // 3)
        out = np.zeros(n, np.int32)
        if n and not _lib.core().wp_b200_bvh_download(self.id, None, out.ctypes.data, None, None, None, None):
            raise RuntimeError(_lib.error_string())
        return out


HALF_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("ib", "<u4")])


def _download_tree(id_, n):
    """Host copies of the LBVH products in the reference's own layout (bvh.h:161-207)."""
    c = _lib.core()
    info = _lib.bvh_info_t()
    if not c.wp_b200_bvh_info(id_, ctypes.byref(info)):
        raise RuntimeError(_lib.error_string())
    m = max(2 * n - 1, 0)
    out = {
        "n": n,
        "leaf_size": info.leaf_size,
        "keys": np.zeros(n, np.uint64 if info.key_bits == 64 else np.uint32),
        "key_bits": info.key_bits,
        "primitive_indices": np.zeros(n, np.int32),
        "node_lowers": np.zeros(m, HALF_DTYPE),
        "node_uppers": np.zeros(m, HALF_DTYPE),
        "parents": np.zeros(m, np.int32),
        "height": info.height,
        "deep": info.deep,
        "total_lower": np.array(info.total_lower[:], np.float32),
        "total_upper": np.array(info.total_upper[:], np.float32),
        "inv_edges": np.array(info.inv_edges[:], np.float32),
    }
    root = np.full(1, -1, np.int32)
    if n > 0:
        ok = c.wp_b200_bvh_download(
            id_, out["keys"].ctypes.data, out["primitive_indices"].ctypes.data, out["node_lowers"].ctypes.data,
            out["node_uppers"].ctypes.data, out["parents"].ctypes.data, root.ctypes.data,
        )  # fmt: skip
        if not ok:
            raise RuntimeError(_lib.error_string())
    out["root"] = int(root[0])
    return out
