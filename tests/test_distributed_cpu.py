"""CPU tests (gloo, world_size 2) of the host-side sharding logic that the NCCL path uses on GPUs:
ShardPlan partitioning, padded all-gather reassembly in global query order, and the TCP bootstrap
that distributes the NCCL unique id."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from warp_b200.distributed import ShardPlan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_plan_partitions_exactly():
    for n in (0, 1, 7, 8, 9, 1000, 1001, 16777216):
        for world in (1, 2, 3, 4, 8):
            plan = ShardPlan(n, world)
            ranges = [plan.range(r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            assert all(0 <= e - s <= plan.shard for s, e in ranges)
            assert sum(plan.count(r) for r in range(world)) == n
            assert plan.padded == plan.shard * world >= n


WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, {root!r})
import torch.distributed as dist
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
import oracle
from warp_b200 import meshgen as mg
from warp_b200.distributed import ShardPlan, GlooCommunicator, gather_fields, exchange_unique_id, part_ranges, FIELD_BYTES
rank = dist.get_rank()
comm = GlooCommunicator()
P, I = mg.noisy_sphere(2, 0.05, 11)
tree = oracle.mesh_lbvh_build(P, I, 4)          # "replicated" mesh + tree: every rank builds its own
Q = mg.box_queries(P, 1001, seed=5)              # odd count: last shard is padded
plan = ShardPlan(Q.shape[0], 2)
s, e = plan.range(rank)
local_q = np.zeros((plan.shard, 3), np.float32); local_q[: e - s] = Q[s:e]
res = oracle.query_point_no_sign(P, I, tree, local_q, 1e6)   # the oracle stands in for the GPU kernel here
local = {{k: res[k] for k in ("result", "face", "u", "v")}}
dt = {{"result": np.uint8, "face": np.int32, "u": np.float32, "v": np.float32}}
glob = gather_fields(local, plan, comm, lambda name, count: np.zeros(count, dt[name]))
want = oracle.query_point_no_sign(P, I, tree, Q, 1e6)
for k in local:
    assert np.array_equal(glob[k][: plan.n], want[k]), (rank, k)
# pipelined gather: the shard goes out in pieces, each placed at its offset inside every rank's slot
for parts in (2, 3, 7):
    pg = {{k: np.zeros(plan.padded, dt[k]) for k in local}}
    pieces = part_ranges(plan.shard, parts)
    assert pieces[0][0] == 0 and pieces[-1][1] == plan.shard and all(x[1] == y[0] for x, y in zip(pieces, pieces[1:]))
    for a, b in pieces:
        for k in local:
            w = FIELD_BYTES[k]
            comm.allgather_part(local[k][a:b], pg[k], w * (b - a), w * plan.shard, w * a)
    for k in local:
        assert np.array_equal(pg[k][: plan.n], want[k]), (rank, parts, k)
# rays: seven fields incl. the 12-byte normals (what sharded_query_ray gathers on GPUs)
S, D = mg.random_rays(P, 777, seed=6)
rplan = ShardPlan(S.shape[0], 2)
s, e = rplan.range(rank)
ls = np.zeros((rplan.shard, 3), np.float32); ld = np.ones((rplan.shard, 3), np.float32)
ls[: e - s] = S[s:e]; ld[: e - s] = D[s:e]
rres = oracle.query_ray(P, I, tree, ls, ld, 1e6)
rdt = {{"result": np.uint8, "sign": np.float32, "face": np.int32, "t": np.float32, "u": np.float32, "v": np.float32}}
rglob = gather_fields({{k: rres[k] for k in ("result", "sign", "face", "t", "u", "v", "normal")}}, rplan, comm,
                      lambda name, count: np.zeros((count, 3), np.float32) if name == "normal" else np.zeros(count, rdt[name]))
rwant = oracle.query_ray(P, I, tree, S, D, 1e6)
for k in rglob:
    assert np.array_equal(rglob[k][: rplan.n], rwant[k]), (rank, k)
# bootstrap: rank 0 invents 128 bytes, rank 1 must receive exactly those
uid = exchange_unique_id(rank, 2, lambda: bytes(range(128)), addr="127.0.0.1", port={port2})
assert uid == bytes(range(128))
comm.barrier()
dist.destroy_process_group()
print("OK", rank)
"""


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_gather_matches_unsharded(oracle_mod):
    pytest.importorskip("torch")
    code = WORKER.format(root=ROOT, port=_free_port(), port2=_free_port())
    procs = [subprocess.Popen([sys.executable, "-c", code, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"OK {r}" in o, o[-2000:]
