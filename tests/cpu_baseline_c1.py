"""Config C1 on the host CPU through the UNMODIFIED reference C++ (oracle/_ref): SAH build + 1 M
mesh_query_point_no_sign queries on an 81 920-triangle icosphere, 1 thread (the reference's CPU launch is a
serial loop) and all threads (courtesy)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from warp_b200 import meshgen as mg
P, I = mg.icosphere(6)
t0 = time.perf_counter(); m = oracle.RefMesh(P, I, oracle.SAH, 4); build = time.perf_counter() - t0
Q = mg.cube_queries(1_000_000, 1.5, 42)
res = {"triangles": len(I) // 3, "sah_build_ms": 1e3 * build, "host_threads": oracle.ref_max_threads()}
for th in (1, 0):
    t0 = time.perf_counter(); r = m.query_point_no_sign(Q, 1e6, nthreads=th); dt = time.perf_counter() - t0
    res[f"qps_{'1_thread' if th == 1 else 'all_threads'}"] = len(Q) / dt
res["found"] = int(r["result"].sum())
print(json.dumps(res))
