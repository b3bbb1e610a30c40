"""Query sharding across GPUs: one process per GPU, mesh + BVH replicated, query batch partitioned
contiguously, SoA results gathered with NCCL over NVLink.  Build and refit are
"replicas only": every rank builds its own (bit-identical, deterministic) tree.

The reference has no multi-GPU driver for this path (its FAQ points at nccl4py / mpi4py,
docs/user_guide/faq.rst:405-452); this module is the new call site.

Host logic (``ShardPlan``, ``gather_fields``) is backend agnostic so it can be exercised on CPU with
a gloo communicator (tests/test_distributed_cpu.py); the product backend is ``NcclCommunicator``.
"""

from __future__ import annotations

import ctypes
import os
import socket
import struct
import time

import numpy as np

from . import _lib


class ShardPlan:
    """Contiguous, equal-size (padded) partition of ``n`` queries over ``world`` ranks.

    Equal sizes make the gather a plain all-gather; the last ranks may own fewer (or zero) real
    queries, their tail is padding that is computed on nothing and trimmed after the gather.
    """

    def __init__(self, n: int, world: int):
        if world < 1 or n < 0:
            raise ValueError("ShardPlan: need world >= 1 and n >= 0")
        self.n, self.world = int(n), int(world)
        self.shard = (self.n + self.world - 1) // self.world if self.n else 0
        self.padded = self.shard * self.world

    def range(self, rank: int):
        """[start, end) of the real queries owned by ``rank``."""
        start = min(rank * self.shard, self.n)
        return start, min(start + self.shard, self.n)

    def count(self, rank: int) -> int:
        s, e = self.range(rank)
        return e - s


class NcclCommunicator:
    """Thin wrapper over the native NCCL entry points (device pointers, library's current stream)."""

    def __init__(self, rank: int, world: int, unique_id: bytes):
        self.rank, self.world = rank, world
        buf = ctypes.create_string_buffer(unique_id, 128)
        if not _lib.core().wp_b200_nccl_init(buf, world, rank):
            raise RuntimeError("NCCL communicator initialisation failed")

    @staticmethod
    def load(path: str | None = None):
        cand = [path] if path else [None]
        try:  # the wheel-bundled NCCL (newer than the system one) if present
            import importlib.util

            spec = importlib.util.find_spec("nvidia.nccl")
            if spec and spec.submodule_search_locations:
                cand.insert(0, os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2"))
        except Exception:
            pass
        for c in cand:
            if c is None or os.path.exists(c):
                if _lib.core().wp_b200_nccl_load(c.encode() if c else None):
                    return
        raise RuntimeError("cannot load libnccl.so.2")

    @staticmethod
    def new_unique_id() -> bytes:
        NcclCommunicator.load()
        buf = ctypes.create_string_buffer(128)
        if not _lib.core().wp_b200_nccl_unique_id(buf):
            raise RuntimeError("ncclGetUniqueId failed")
        return buf.raw

    def allgather(self, send, recv, nbytes_per_rank: int):
        """``send`` / ``recv``: objects with ``.ptr`` (device arrays)."""
        if not _lib.core().wp_b200_nccl_allgather(ctypes.c_void_p(send.ptr), ctypes.c_void_p(recv.ptr), nbytes_per_rank):
            raise RuntimeError("ncclAllGather failed")

    def allgather_part(self, send_ptr: int, recv, part_bytes: int, shard_stride_bytes: int, offset_bytes: int):
        """One part of every rank's shard into the rank-major ``recv``, on the communication stream."""
        if not _lib.core().wp_b200_nccl_allgather_part(ctypes.c_void_p(send_ptr), ctypes.c_void_p(recv.ptr), part_bytes,
                                                       shard_stride_bytes, offset_bytes):
            raise RuntimeError("pipelined all-gather failed")

    def allgather_multi(self, fields, comm_stream: bool = False):
        """``fields``: list of (send, recv, nbytes_per_rank) device arrays -- gathered in ONE NCCL launch (a group of
        all-gathers), on the current stream or, with ``comm_stream=True``, on the library's communication stream."""
        k = len(fields)
        send = (ctypes.c_void_p * k)(*[f[0].ptr for f in fields])
        recv = (ctypes.c_void_p * k)(*[f[1].ptr for f in fields])
        nbytes = (ctypes.c_size_t * k)(*[int(f[2]) for f in fields])
        if not _lib.core().wp_b200_nccl_allgather_multi(send, recv, nbytes, k, 1 if comm_stream else 0):
            raise RuntimeError("grouped ncclAllGather failed")

    def map_peers(self, buffers):
        """CUDA IPC mapping of every rank's ``buffers`` (a list of device arrays, same order on every rank) into this
        process: returns a :class:`PeerMap` for :meth:`p2p_allgather_multi`.  Collective (one NCCL all-gather of the
        72-byte handles)."""
        from .types import array, uint8

        c = _lib.core()
        k = len(buffers)
        mine = np.zeros((k, 72), np.uint8)
        for i, b in enumerate(buffers):
            if not c.wp_b200_ipc_get_handle(ctypes.c_void_p(b.ptr), mine[i].ctypes.data):
                raise RuntimeError("cudaIpcGetMemHandle failed")
        send = array(mine.reshape(-1), dtype=uint8, device=buffers[0].device)
        recv = array(np.zeros(self.world * k * 72, np.uint8), dtype=uint8, device=buffers[0].device)
        self.allgather(send, recv, k * 72)
        c.wp_cuda_stream_synchronize(c.wp_cuda_context_get_stream(None))
        handles = recv.numpy().reshape(self.world, k, 72).copy()
        table = (ctypes.c_void_p * (self.world * k))()
        for r in range(self.world):
            for i in range(k):
                if r == self.rank:
                    continue
                p = c.wp_b200_ipc_open_handle(handles[r, i].ctypes.data)
                if not p:
                    raise RuntimeError(f"cudaIpcOpenMemHandle failed for rank {r}'s buffer {i}")
                table[r * k + i] = p
        return PeerMap(self, buffers, handles, table)

    def p2p_allgather_multi(self, sends, peer_map, nbytes):
        """Every rank pushes its shard of each field into every rank's buffer (copy engines over NVLink), on the
        communication stream; ``sends[i]`` / ``nbytes[i]`` per field, ``peer_map`` from :meth:`map_peers`."""
        k = len(sends)
        send = (ctypes.c_void_p * k)(*[a.ptr for a in sends])
        own = (ctypes.c_void_p * k)(*[b.ptr for b in peer_map.buffers])
        nb = (ctypes.c_size_t * k)(*[int(x) for x in nbytes])
        if not _lib.core().wp_b200_p2p_allgather_multi(send, own, peer_map.table, nb, k, self.rank):
            raise RuntimeError("peer-memory all-gather failed")

    def on_comm_stream(self):
        """Context manager: the library's communication stream is the device's current stream inside the block."""
        return _CommStreamScope()

    def mark(self, k: int):
        if not _lib.core().wp_b200_nccl_mark(int(k)):
            raise RuntimeError("NCCL mark failed")

    def wait_mark(self, k: int):
        if not _lib.core().wp_b200_nccl_wait_mark(int(k)):
            raise RuntimeError("NCCL wait_mark failed")

    def fork(self):
        if not _lib.core().wp_b200_nccl_fork():
            raise RuntimeError("NCCL fork failed")

    def join(self):
        if not _lib.core().wp_b200_nccl_join():
            raise RuntimeError("NCCL join failed")

    def allreduce_max(self, arr):
        if not _lib.core().wp_b200_nccl_allreduce_max_f32(ctypes.c_void_p(arr.ptr), arr.size):
            raise RuntimeError("ncclAllReduce failed")

    def barrier(self):
        if not _lib.core().wp_b200_nccl_barrier():
            raise RuntimeError("NCCL barrier failed")

    def close(self):
        _lib.core().wp_b200_nccl_destroy()


class _CommStreamScope:
    def __enter__(self):
        c = _lib.core()
        self.saved = c.wp_cuda_context_get_stream(None)
        c.wp_cuda_context_set_stream(None, c.wp_b200_nccl_comm_stream(), 0)
        return self

    def __exit__(self, *exc):
        _lib.core().wp_cuda_context_set_stream(None, self.saved, 0)
        return False


class PeerMap:
    """Peers' buffers mapped into this process (CUDA IPC); unmapped when the object goes away."""

    def __init__(self, comm, buffers, handles, table):
        self.comm, self.buffers, self.handles, self.table = comm, buffers, handles, table

    def close(self):
        c = _lib.core()
        k = len(self.buffers)
        for r in range(self.comm.world):
            for i in range(k):
                p = self.table[r * k + i]
                if p:
                    c.wp_b200_ipc_close_handle(ctypes.c_void_p(p), self.handles[r, i].ctypes.data)
                    self.table[r * k + i] = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def exchange_unique_id(rank: int, world: int, make_id, addr: str | None = None, port: int | None = None,
                       timeout: float = 300.0) -> bytes:
    """Rank 0 creates an id with ``make_id()`` and serves it over TCP; the others fetch it.

    Defaults come from the torchrun environment (MASTER_ADDR, MASTER_PORT + 1 -- torchrun's own
    store owns MASTER_PORT); override the port with WARP_B200_BOOTSTRAP_PORT.
    """
    addr = addr or os.environ.get("MASTER_ADDR", "127.0.0.1")
    if port is None:
        port = int(os.environ.get("WARP_B200_BOOTSTRAP_PORT", int(os.environ.get("MASTER_PORT", "29500")) + 1))
    if world == 1:
        return make_id()
    if rank == 0:
        uid = make_id()
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind((addr, port))
        srv.listen(world)
        srv.settimeout(timeout)
        served = 0
        while served < world - 1:
            conn, _ = srv.accept()
            conn.sendall(struct.pack("<I", len(uid)) + uid)
            conn.close()
            served += 1
        srv.close()
        return uid
    deadline = time.time() + timeout
    while True:
        try:
            s = socket.create_connection((addr, port), timeout=5.0)
            break
        except OSError:
            if time.time() > deadline:
                raise RuntimeError(f"rank {rank}: cannot reach the bootstrap server at {addr}:{port}") from None
            time.sleep(0.1)
    data = b""
    while len(data) < 4:
        data += s.recv(4 - len(data))
    (ln,) = struct.unpack("<I", data)
    uid = b""
    while len(uid) < ln:
        chunk = s.recv(ln - len(uid))
        if not chunk:
            raise RuntimeError("bootstrap connection closed early")
        uid += chunk
    s.close()
    return uid


def init_from_env():
    """(rank, world, local_rank, communicator-or-None) from RANK / WORLD_SIZE / LOCAL_RANK (torchrun)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    _lib.core().wp_cuda_context_set_current(ctypes.c_void_p(local_rank + 1))
    if world == 1:
        return rank, world, local_rank, None
    NcclCommunicator.load()
    uid = exchange_unique_id(rank, world, NcclCommunicator.new_unique_id)
    return rank, world, local_rank, NcclCommunicator(rank, world, uid)


FIELD_BYTES = {"result": 1, "sign": 4, "face": 4, "t": 4, "u": 4, "v": 4, "normal": 12}


def gather_fields(local: dict, plan: ShardPlan, comm, alloc):
    """All-gather every field of a per-rank SoA result (arrays of ``plan.shard`` entries) into arrays
    of ``plan.padded`` entries laid out rank-major, i.e. in global query order.

    ``alloc(field, count)`` returns a destination buffer; ``comm.allgather(send, recv, nbytes)``
    moves the bytes.  Returns {field: gathered buffer}; callers trim to ``plan.n``.
    """
    out = {name: alloc(name, plan.padded) for name in local}
    if hasattr(comm, "allgather_multi"):  # one launch for the whole result set
        comm.allgather_multi([(arr, out[name], plan.shard * FIELD_BYTES[name]) for name, arr in local.items()])
    else:
        for name, arr in local.items():
            comm.allgather(arr, out[name], plan.shard * FIELD_BYTES[name])
    return out


def _sharded(local: dict, dtypes: dict, plan: ShardPlan, comm, dev, global_out):
    from .types import empty

    if comm is None:
        return local, plan.n
    if global_out is None:
        global_out = {k: empty(plan.padded, dtypes[k], dev) for k in local}
    return gather_fields(local, plan, comm, lambda name, count: global_out[name]), plan.n


def part_ranges(shard: int, parts: int):
    """Contiguous sub-ranges of a shard for the pipelined gather: ``parts`` nearly equal pieces (fewer for tiny shards)."""
    parts = max(1, min(int(parts), shard)) if shard > 0 else 1
    step = -(-shard // parts)
    return [(a, min(a + step, shard)) for a in range(0, max(shard, 1), max(step, 1)) if a < shard or shard == 0]


def sharded_query_point_no_sign(mesh, local_points, plan: ShardPlan, max_dist: float, comm, rank: int,
                                local_out=None, global_out=None, parts: int = 1):
    """Each rank answers its shard of a global batch; every rank ends up with all ``plan.n`` answers.

    ``local_points``: device array with ``plan.shard`` vec3 entries (entries past ``plan.count(rank)``
    are padding).  Returns ``(global_fields, n)``; fields are device arrays of ``plan.padded`` entries.
    ``parts`` > 1 pipelines the step: the shard is answered in that many pieces and the gather of
    piece k runs on the communication stream under the traversal of piece k + 1.
    """
    from . import _lib
    from .queries import MeshQueryPoint, mesh_query_point_no_sign
    from .types import empty, float32, int32, uint8

    if comm is not None and parts > 1 and hasattr(comm, "allgather_part") and plan.shard > 0:
        import ctypes

        dev = mesh.device
        n = plan.shard
        if local_out is None:
            local_out = MeshQueryPoint(empty(n, uint8, dev), empty(n, float32, dev).zero_(), empty(n, int32, dev),
                                       empty(n, float32, dev), empty(n, float32, dev))  # fmt: skip
        dtypes = {"result": uint8, "face": int32, "u": float32, "v": float32}
        if global_out is None:
            global_out = {k: empty(plan.padded, dt, dev) for k, dt in dtypes.items()}
        c = _lib.core()
        vp = ctypes.c_void_p
        for a, b in part_ranges(n, parts):
            ok = c.wp_b200_mesh_query_point_no_sign(mesh.id, vp(local_points.ptr + 12 * a), b - a, float(max_dist),
                                                    vp(local_out.result.ptr + a), vp(local_out.face.ptr + 4 * a),
                                                    vp(local_out.u.ptr + 4 * a), vp(local_out.v.ptr + 4 * a))  # fmt: skip
            if not ok:
                raise RuntimeError(f"mesh_query_point_no_sign failed: {_lib.error_string()}")
            comm.fork()  # the communication stream picks up after this piece's traversal
            for name in dtypes:
                w = FIELD_BYTES[name]
                comm.allgather_part(getattr(local_out, name).ptr + w * a, global_out[name], w * (b - a), w * n, w * a)
        comm.join()  # the step ends when the last piece has been gathered
        return global_out, plan.n

    res = mesh_query_point_no_sign(mesh, local_points, max_dist, out=local_out)
    local = {"result": res.result, "face": res.face, "u": res.u, "v": res.v}
    dtypes = {"result": uint8, "face": int32, "u": float32, "v": float32}
    return _sharded(local, dtypes, plan, comm, mesh.device, global_out)


def sharded_query_point(mesh, local_points, plan: ShardPlan, max_dist: float, comm, rank: int, local_out=None,
                        global_out=None):
    """Signed closest point (``mesh_query_point``), sharded like :func:`sharded_query_point_no_sign`; adds ``sign``."""
    from .queries import mesh_query_point
    from .types import float32, int32, uint8

    res = mesh_query_point(mesh, local_points, max_dist, out=local_out)
    local = {"result": res.result, "sign": res.sign, "face": res.face, "u": res.u, "v": res.v}
    dtypes = {"result": uint8, "sign": float32, "face": int32, "u": float32, "v": float32}
    return _sharded(local, dtypes, plan, comm, mesh.device, global_out)


RAY_WIRE_FIELDS = ("result", "sign", "face", "t", "u", "v")  # 21 bytes per ray cross NVLink; `normal` does not


def ray_normals_from_faces(mesh, result, face, out):
    """``normal`` of ``mesh_query_ray`` recomputed from (``result``, ``face``) on this rank's replica of the mesh:
    ``normalize(cross(b - a, c - a))`` of the hit face (``mesh.h:1880-1886``), the zero vector for a miss -- the same
    arithmetic on the same vertices, hence the same bits as the normal the tracing rank produced."""
    n = len(face)
    ok = _lib.core().wp_b200_mesh_eval_face_normal_masked(mesh.id, ctypes.c_void_p(face.ptr), ctypes.c_void_p(result.ptr), n,
                                                           ctypes.c_void_p(out.ptr))
    if not ok:
        raise RuntimeError(f"ray normal evaluation failed: {_lib.error_string()}")
    return out


def sharded_query_ray(mesh, local_starts, local_dirs, plan: ShardPlan, max_t: float, comm, rank: int, local_out=None,
                      global_out=None):
    """Closest ray hit (``mesh_query_ray``) for this rank's shard of a global ray batch.  ``result``, ``sign``,
    ``face``, ``t``, ``u``, ``v`` (21 bytes per ray) are all-gathered into global ray order in one NCCL launch; the
    12-byte ``normal`` is NOT sent -- every rank holds a replica of the mesh and recomputes it from the gathered
    (``result``, ``face``), bit for bit (:func:`ray_normals_from_faces`)."""
    from .queries import mesh_query_ray
    from .types import empty, float32, int32, uint8, vec3

    res = mesh_query_ray(mesh, local_starts, local_dirs, max_t, out=local_out)
    dtypes = {"result": uint8, "sign": float32, "face": int32, "t": float32, "u": float32, "v": float32, "normal": vec3}
    if comm is None:
        return {k: getattr(res, k) for k in dtypes}, plan.n
    if global_out is None:
        global_out = {k: empty(plan.padded, dt, mesh.device) for k, dt in dtypes.items()}
    local = {k: getattr(res, k) for k in RAY_WIRE_FIELDS}
    gather_fields(local, plan, comm, lambda name, count: global_out[name])
    ray_normals_from_faces(mesh, global_out["result"], global_out["face"], global_out["normal"])
    return global_out, plan.n


class QueryPipeline:
    """Software pipeline over a stream of equally sized query batches ("pipeline in chunks of 8-16 M
    queries so the gather of chunk k overlaps the traversal of chunk k+1"): ``submit()`` answers this rank's shard of
    batch k on the compute stream into one of two local result sets and hands the gather of that set -- one grouped
    NCCL launch -- to the communication stream, where it runs under the traversal of batch k + 1.  ``result(k)`` /
    ``finish()`` order the compute stream after the gathers.  ``kind``: "point_no_sign" | "point" | "ray"."""

    POINT_FIELDS = ("result", "face", "u", "v")

    def __init__(self, mesh, plan: ShardPlan, comm, kind: str = "point_no_sign", max_dist: float = 1.0e6, depth: int = 2):
        from .queries import MeshQueryPoint, MeshQueryRay
        from .types import empty, float32, int32, uint8, vec3

        self.mesh, self.plan, self.comm, self.kind, self.max_dist, self.depth = mesh, plan, comm, kind, float(max_dist), depth
        dev, n = mesh.device, plan.shard
        transport = os.environ.get("WARP_B200_GATHER", "p2p") if (comm is not None and hasattr(comm, "map_peers")) else "nccl"
        if kind == "ray":
            # the 12-byte normal: sent with the other fields when the gather runs on copy engines (peer-memory pushes cost
            # no SM time; measured at 8 GPUs: recomputing 134 M normals per step takes ~5 ms of SM / HBM time on every rank,
            # sending them 2 ms more of NVLink time that hides under the traversal), recomputed from the gathered
            # (result, face) when the gather is an SM-driven NCCL launch.  WARP_B200_RAY_WIRE=all / faces overrides.
            wire = os.environ.get("WARP_B200_RAY_WIRE", "all" if transport == "p2p" else "faces")
            self.wire = RAY_WIRE_FIELDS + (("normal",) if wire == "all" else ())
            mk = lambda cnt: MeshQueryRay(empty(cnt, uint8, dev), empty(cnt, float32, dev), empty(cnt, int32, dev),  # noqa: E731
                                          empty(cnt, float32, dev), empty(cnt, float32, dev), empty(cnt, float32, dev),
                                          empty(cnt, vec3, dev))
        else:
            self.wire = self.POINT_FIELDS + (("sign",) if kind == "point" else ())
            mk = lambda cnt: MeshQueryPoint(empty(cnt, uint8, dev), empty(cnt, float32, dev), empty(cnt, int32, dev),  # noqa: E731
                                            empty(cnt, float32, dev), empty(cnt, float32, dev))
        self.local = [mk(n) for _ in range(depth)]
        self.gathered = [mk(plan.padded) for _ in range(depth)] if comm is not None else self.local
        self.submitted = 0
        # transport of the gather: "p2p" = every rank pushes its shard into the peers' buffers with copy-engine memcpys
        # over NVLink (no SM taken from the traversal it runs under), "nccl" = one grouped ncclAllGather launch
        self.transport = transport
        self.peers = None
        if comm is not None and self.transport == "p2p":
            try:
                self.peers = [comm.map_peers([getattr(g, f) for f in self.wire]) for g in self.gathered]
            except RuntimeError:
                self.transport, self.peers = "nccl", None

    def submit(self, *inputs):
        from .queries import mesh_query_point, mesh_query_point_no_sign, mesh_query_ray

        k = self.submitted
        slot = k % self.depth
        if self.comm is not None and k >= self.depth:
            self.comm.wait_mark(slot)  # the gather that last read this local set (and wrote this gathered set) is done
        if self.kind == "ray":
            mesh_query_ray(self.mesh, inputs[0], inputs[1], self.max_dist, out=self.local[slot])
        elif self.kind == "point":
            mesh_query_point(self.mesh, inputs[0], self.max_dist, out=self.local[slot])
        else:
            mesh_query_point_no_sign(self.mesh, inputs[0], self.max_dist, out=self.local[slot])
        if self.comm is not None:
            self.comm.fork()
            if self.peers is not None:
                self.comm.p2p_allgather_multi([getattr(self.local[slot], f) for f in self.wire], self.peers[slot],
                                              [self.plan.shard * FIELD_BYTES[f] for f in self.wire])
            else:
                self.comm.allgather_multi([(getattr(self.local[slot], f), getattr(self.gathered[slot], f),
                                            self.plan.shard * FIELD_BYTES[f]) for f in self.wire], comm_stream=True)
            if self.kind == "ray" and "normal" not in self.wire and hasattr(self.comm, "on_comm_stream"):
                # the normals of ALL ranks' rays, recomputed from the gathered (result, face) on this rank's replica --
                # in communication order, so that they too run under the next batch's traversal
                g = self.gathered[slot]
                with self.comm.on_comm_stream():
                    ray_normals_from_faces(self.mesh, g.result, g.face, g.normal)
            self.comm.mark(slot)
        self.submitted += 1
        return k

    def result(self, k: int):
        """Result set of batch ``k`` (valid for the last ``depth`` batches), ordered after its gather (and, for rays, the
        recomputation of the normals) on the compute stream."""
        slot = k % self.depth
        if self.comm is not None:
            self.comm.wait_mark(slot)
            if self.kind == "ray" and "normal" not in self.wire and not hasattr(self.comm, "on_comm_stream"):
                g = self.gathered[slot]
                ray_normals_from_faces(self.mesh, g.result, g.face, g.normal)
        return self.gathered[slot]

    def finish(self):
        if self.comm is not None:
            self.comm.join()


class GlooCommunicator:
    """CPU stand-in used by the world_size-2 tests: same ``allgather`` contract over numpy buffers."""

    def __init__(self):
        import torch.distributed as dist

        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def allgather(self, send: np.ndarray, recv: np.ndarray, nbytes_per_rank: int):
        import torch

        s = torch.from_numpy(send.view(np.uint8).reshape(-1)[:nbytes_per_rank].copy())
        parts = [torch.empty(nbytes_per_rank, dtype=torch.uint8) for _ in range(self.world)]
        self.dist.all_gather(parts, s)
        flat = recv.view(np.uint8).reshape(-1)
        for r, p in enumerate(parts):
            flat[r * nbytes_per_rank : (r + 1) * nbytes_per_rank] = p.numpy()

    def allgather_part(self, send: np.ndarray, recv: np.ndarray, part_bytes: int, shard_stride_bytes: int, offset_bytes: int):
        """CPU stand-in for NcclCommunicator.allgather_part (``send`` = the part itself, as a numpy view)."""
        import torch

        world = self.dist.get_world_size()
        mine = torch.from_numpy(np.ascontiguousarray(send).view(np.uint8).reshape(-1)[:part_bytes].copy())
        bufs = [torch.empty(part_bytes, dtype=torch.uint8) for _ in range(world)]
        self.dist.all_gather(bufs, mine)
        flat = recv.view(np.uint8).reshape(-1)
        for r, b in enumerate(bufs):
            flat[r * shard_stride_bytes + offset_bytes : r * shard_stride_bytes + offset_bytes + part_bytes] = b.numpy()

    def allgather_multi(self, fields, comm_stream: bool = False):
        for send, recv, nbytes in fields:
            self.allgather(send, recv, nbytes)

    def mark(self, k: int):
        pass

    def wait_mark(self, k: int):
        pass

    def fork(self):
        pass

    def join(self):
        pass

    def barrier(self):
        self.dist.barrier()
