"""torchrun -n N: what limits the sharded ray step -- the gather alone (per transport), the traversal alone, both pipelined."""
import os, sys, json, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import warp_b200 as wp
from warp_b200 import _lib, distributed, meshgen as mg
from bench import event_ms, max_over_ranks

core = _lib.core()
rank, world, local_rank, comm = distributed.init_from_env()
dev = f"cuda:{local_rank}"
stream = core.wp_cuda_context_get_stream(None)
Ph, Ih = mg.heightfield(2237, 4)
hm = wp.Mesh(wp.array(Ph, dtype=wp.vec3, device=dev), wp.array(Ih, dtype=wp.int32, device=dev))
S, D = mg.pinhole_rays(4096, 4096, eye=(0.5 + 0.02 * rank, -0.6, 0.9))
s_d, d_d = wp.array(S, dtype=wp.vec3, device=dev), wp.array(D, dtype=wp.vec3, device=dev)
n = len(S)
plan = distributed.ShardPlan(n * world, world)
res = {"n_gpus": world, "rays_per_gpu": n}
for transport in ("p2p", "nccl"):
    os.environ["WARP_B200_GATHER"] = transport
    rp = distributed.QueryPipeline(hm, plan, comm, "ray", 1.0e6)
    for _ in range(2):
        rp.submit(s_d, d_d)
    rp.finish(); comm.barrier(); core.wp_cuda_context_synchronize(None)
    K = 12
    def piped():
        for _ in range(K):
            rp.submit(s_d, d_d)
        rp.finish()
    ms_p = max_over_ranks(wp, comm, event_ms(core, piped, stream) / K, dev)
    # gather alone: same communication calls, no traversal in between
    fields = [getattr(rp.local[0], f) for f in rp.wire]
    nbytes = [plan.shard * distributed.FIELD_BYTES[f] for f in rp.wire]
    def gathers():
        for _ in range(K):
            comm.fork()
            if rp.peers is not None:
                comm.p2p_allgather_multi(fields, rp.peers[0], nbytes)
            else:
                comm.allgather_multi([(a, getattr(rp.gathered[0], f), b) for a, f, b in zip(fields, rp.wire, nbytes)], comm_stream=True)
        comm.join()
    gathers(); comm.barrier(); core.wp_cuda_context_synchronize(None)
    ms_g = max_over_ranks(wp, comm, event_ms(core, gathers, stream) / K, dev)
    res[transport] = {"pipelined_ms_per_step": ms_p, "gather_alone_ms": ms_g, "wire_bytes_in_per_rank": sum(nbytes) * (world - 1),
                      "gather_alone_GBps_in_per_rank": sum(nbytes) * (world - 1) / (ms_g * 1e-3) / 1e9}
    del rp
    comm.barrier()
out = wp.mesh_query_ray(hm, s_d, d_d, 1.0e6)
core.wp_cuda_context_synchronize(None)
res["traversal_alone_ms"] = statistics.median([event_ms(core, lambda: wp.mesh_query_ray(hm, s_d, d_d, 1.0e6, out=out), stream) for _ in range(5)])
if rank == 0:
    print("GATHER_PROBE " + json.dumps(res), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/gather_probe_n{world}.json", "w"), indent=1)
comm.barrier()
comm.close()
