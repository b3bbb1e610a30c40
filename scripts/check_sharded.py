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
comm.barrier()
print("sharded OK", rank, flush=True)
comm.close()
