"""torchrun -n 2..8: the pipelined gather (parts = 4, 3) gives bit for bit what the plain all-gather gives, and rank 0
checks the gathered answers of ALL ranks against single-GPU answers on the concatenated batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import warp_b200 as wp
from warp_b200 import distributed, meshgen as mg

rank, world, local_rank, comm = distributed.init_from_env()
dev = wp.get_device(f"cuda:{local_rank}")
P, I = mg.noisy_sphere(6, 0.02, 1)
mesh = wp.Mesh(wp.array(P, dtype=wp.vec3, device=dev), wp.array(I, dtype=wp.int32, device=dev))
n_total = 1_000_003 * world
Q = mg.box_queries(P, n_total, seed=9)
plan = distributed.ShardPlan(n_total, world)
s, e = plan.range(rank)
local = np.zeros((plan.shard, 3), np.float32)
local[: e - s] = Q[s:e]
q_dev = wp.array(local, dtype=wp.vec3, device=dev)
ref, _ = distributed.sharded_query_point_no_sign(mesh, q_dev, plan, 1e6, comm, rank)
ref = {k: v.numpy().copy() for k, v in ref.items()}
for parts in (4, 3, 2):
    got, _ = distributed.sharded_query_point_no_sign(mesh, q_dev, plan, 1e6, comm, rank, parts=parts)
    wp.synchronize()
    for k in ref:
        assert np.array_equal(got[k].numpy(), ref[k]), (rank, parts, k)
if rank == 0:
    # rank-major layout: rank r's answers sit at [r * shard, r * shard + count(r))
    whole = wp.mesh_query_point_no_sign(mesh, wp.array(Q, dtype=wp.vec3, device=dev), 1e6).numpy()
    for r in range(world):
        a, b = plan.range(r)
        for k in ("result", "face", "u", "v"):
            assert np.array_equal(ref[k][r * plan.shard : r * plan.shard + (b - a)], whole[k][a:b]), (r, k)
# QueryPipeline (cross-batch pipelining of the gather) over both transports: peer-memory pushes and grouped NCCL
S, D = mg.random_rays(P, plan.shard, seed=20 + rank)
s_d, d_d = wp.array(S, dtype=wp.vec3, device=dev), wp.array(D, dtype=wp.vec3, device=dev)
ray_ref, _ = distributed.sharded_query_ray(mesh, s_d, d_d, plan, 1e6, comm, rank)
ray_ref = {k: v.numpy().copy() for k, v in ray_ref.items()}
for transport in ("p2p", "nccl"):
    os.environ["WARP_B200_GATHER"] = transport
    pipe = distributed.QueryPipeline(mesh, plan, comm, "point_no_sign", 1e6)
    assert pipe.transport == transport, (pipe.transport, transport)
    for step in range(5):
        k = pipe.submit(q_dev)
    g = pipe.result(k)
    pipe.finish()
    wp.synchronize()
    for f in ("result", "face", "u", "v"):
        assert np.array_equal(getattr(g, f).numpy(), ref[f]), (rank, transport, f)
    rp = distributed.QueryPipeline(mesh, plan, comm, "ray", 1e6)
    for step in range(4):
        k = rp.submit(s_d, d_d)
    g = rp.result(k)
    rp.finish()
    wp.synchronize()
    for f in ("result", "sign", "face", "t", "u", "v", "normal"):
        assert np.array_equal(getattr(g, f).numpy(), ray_ref[f]), (rank, transport, f)
    del pipe, rp, g
    comm.barrier()
comm.barrier()
print("sharded OK", rank, flush=True)
comm.close()
