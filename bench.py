#!/usr/bin/env python
"""Benchmark of the mesh/BVH spatial-query hot path (BASELINE.json metric, config C2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one batch of 16 777 216 closest-point queries (mesh_query_point_no_sign, max_dist 1e6)
against the LBVH of a 1 310 720-triangle noisy sphere (icosphere, 8 subdivisions) -- BASELINE.json
configs[1].  `value` is whole-job queries/s with inputs and outputs resident in HBM; `e2e` is the same
batch through the public API with pinned HOST buffers (H2D + D2H inside the timed region).  With
N > 1 (launched by torchrun, one rank per GPU) the mesh and tree are replicated, every rank answers
its own 16 M-query shard (weak scaling) and the SoA results are all-gathered with NCCL inside the
timed region.  LBVH build / refit times, ray throughput (config C3) and their rooflines ride along
in `extra`.  `--impl reference` times the reference's own CPU implementation (oracle/_ref, the
unmodified reference C++; falls back to the C port) on a bounded sample of the same workload.

No torch: device memory, streams, events and NCCL all go through libwarp_b200.so.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "closest_point_queries_per_s"
UNIT = "queries/s"
SUBDIV = 8                 # 20 * 4**8 = 1 310 720 triangles
NQ = 1 << 24               # 16 777 216 queries per GPU
MAX_DIST = 1.0e6
WORKLOAD = "C2: LBVH of 1.31M-triangle noisy icosphere; 16.8M mesh_query_point_no_sign queries in 1.2x AABB per GPU"


# pieces of the pipelined gather at N > 1 (distributed.sharded_query_point_no_sign(parts=)).  Measured at 2 GPUs:
# 1 piece 24.9 ms/step, 4 pieces 32.6 ms -- a quarter-size batch is Morton-sorted on its own and is sparser than the
# whole batch, which costs more than the 0.7 ms (2 GPUs) / 2.7 ms (8 GPUs) of gather it hides.  So: one piece.
GATHER_PARTS = int(os.environ.get("BENCH_GATHER_PARTS", "1"))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)  # fmt: skip
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass
        return False

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, val in zip(names, r[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ------------------------------------------------------------------------------------------------

def cpu_reference_setup():
    """Reference CPU objects for the C2 mesh: what `wp.Mesh(..., device='cpu')` gives a user = SAH tree,
    leaf 4 (warp/_src/types.py:6186-6191), queried by the reference's own mesh.h code."""
    import oracle
    from warp_b200 import meshgen as mg

    P, I = mg.noisy_sphere(SUBDIV, 0.02, 1)
    if oracle.ref_available():
        t0 = time.perf_counter()
        mesh = oracle.RefMesh(P, I, oracle.SAH, 4)
        build_s = time.perf_counter() - t0
        cores = oracle.ref_max_threads()

        def run(q, threads):
            return mesh.query_point_no_sign(q, MAX_DIST, nthreads=threads)

        return P, I, run, "reference", cores, build_s
    t0 = time.perf_counter()
    tree = oracle.mesh_lbvh_build(P, I, 4)
    build_s = time.perf_counter() - t0

    def run(q, threads):
        return oracle.query_point_no_sign(P, I, tree, q, MAX_DIST)

    return P, I, run, "port", 1, build_s


def cpu_baseline(sample_queries: int):
    from warp_b200 import meshgen as mg

    P, I, run, kind, cores, build_s = cpu_reference_setup()
    q = mg.box_queries(P, sample_queries, seed=2)
    run(q[:2048], cores)
    t0 = time.perf_counter()
    run(q, cores)
    dt = time.perf_counter() - t0
    return {
        "value": sample_queries / dt, "unit": UNIT, "cores": cores, "kind": kind,
        "sample": f"{sample_queries} of the {NQ} queries (seed 2) on the reference's CPU tree "
                  f"({'SAH, leaf 4, built in %.2f s' % build_s if kind == 'reference' else 'oracle LBVH, leaf 4'}), "
                  f"{cores} host thread(s), {dt:.2f} s",
    }  # fmt: skip


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from warp_b200 import meshgen as mg

    P, I, run, kind, cores, build_s = cpu_reference_setup()
    sample = 1 << 18
    q = mg.box_queries(P, sample * (args.steps + args.warmup), seed=2)
    for w in range(args.warmup):
        run(q[w * sample : (w + 1) * sample], cores)
    t0 = time.perf_counter()
    for k in range(args.steps):
        s = (args.warmup + k) * sample
        run(q[s : s + sample], cores)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "step": f"bounded sample: {sample} queries per step on the host CPU"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{sample} queries/step x {args.steps} steps, reference CPU tree built in {build_s:.2f} s"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

class Pinned:
    """numpy view over page-locked host memory from wp_alloc_pinned."""

    def __init__(self, core, shape, dtype):
        self.core = core
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = core.wp_alloc_pinned(max(n, 1), b"bench")
        if not self.ptr:
            raise RuntimeError("pinned allocation failed")
        buf = (ctypes.c_char * max(n, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        self.core.wp_free_pinned(ctypes.c_void_p(self.ptr))


def event_ms(core, fn, stream):
    a, b = core.wp_cuda_event_create(None, 0), core.wp_cuda_event_create(None, 0)
    core.wp_cuda_event_record(a, stream, 0)
    fn()
    core.wp_cuda_event_record(b, stream, 0)
    core.wp_cuda_event_synchronize(b)
    ms = core.wp_cuda_event_elapsed_time(a, b)
    core.wp_cuda_event_destroy(a)
    core.wp_cuda_event_destroy(b)
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--queries", type=int, default=NQ, help="queries per GPU per step (default = config C2)")
    ap.add_argument("--no-extra", action="store_true", help="skip the build/refit/ray side measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference_arm(args)

    import warp_b200 as wp
    from warp_b200 import _lib, distributed, meshgen as mg

    core = _lib.core()
    rank, world, local_rank, comm = distributed.init_from_env()
    if world != args.gpus and rank == 0:
        print(f"[bench] note: WORLD_SIZE={world} but --gpus {args.gpus}; using WORLD_SIZE", file=sys.stderr)
    dev = f"cuda:{local_rank}"
    stream = core.wp_cuda_context_get_stream(None)
    peak_gbs, peak_src = measured_peaks()
    nq = args.queries

    # ---- workload (synthetic, generated on the host once) -----------------------------------------
    P, I = mg.noisy_sphere(SUBDIV, 0.02, 1)
    T = len(I) // 3
    pts = wp.array(P, dtype=wp.vec3, device=dev)
    idx = wp.array(I, dtype=wp.int32, device=dev)
    core.wp_cuda_context_synchronize(None)
    t0 = time.perf_counter()
    mesh = wp.Mesh(pts, idx, bvh_constructor="lbvh")  # leaf 4 (reference default)
    core.wp_cuda_context_synchronize(None)
    first_build_ms = 1e3 * (time.perf_counter() - t0)
    Qh = mg.box_queries(P, nq, seed=2 + rank)
    q_dev = wp.array(Qh, dtype=wp.vec3, device=dev)
    out = wp.MeshQueryPoint(*(wp.empty(nq, dt, dev) for dt in (wp.uint8, wp.float32, wp.int32, wp.float32, wp.float32)))
    flush = wp.empty(256 << 20, wp.uint8, dev)  # L2 flush buffer (B200 L2 = 126 MB)
    plan = distributed.ShardPlan(nq * world, world)
    gathered = None
    if comm is not None:
        gathered = {"result": wp.empty(plan.padded, wp.uint8, dev), "face": wp.empty(plan.padded, wp.int32, dev),
                    "u": wp.empty(plan.padded, wp.float32, dev), "v": wp.empty(plan.padded, wp.float32, dev)}  # fmt: skip

    def step():
        # one launch of k_query_point (+ 4 NCCL all-gathers when sharded)
        # sharded: the shard is answered in GATHER_PARTS pieces, the gather of a piece runs under the next piece's traversal
        distributed.sharded_query_point_no_sign(mesh, q_dev, plan, MAX_DIST, comm, rank, local_out=out, global_out=gathered,
                                                parts=GATHER_PARTS if comm is not None else 1)

    def l2_flush():
        core.wp_memset_device(None, ctypes.c_void_p(flush.ptr), 0, flush.nbytes, stream)

    def barrier():
        core.wp_cuda_context_synchronize(None)
        if comm is not None:
            comm.barrier()
        core.wp_cuda_context_synchronize(None)

    # ---- timed region: W warm-up steps, then exactly K steps, device-timed, L2 flushed in between -----
    step_ms = []
    with ClockSampler(local_rank) as clocks:  # started before the warm-up so nvidia-smi is already sampling when timing starts
        for _ in range(args.warmup):
            step()
        barrier()
        wall0 = time.perf_counter()
        for _ in range(args.steps):
            l2_flush()
            step_ms.append(event_ms(core, step, stream))
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - wall0)
    total_ms = float(sum(step_ms))
    if comm is not None:  # max over ranks, on the device
        tmax = wp.array(np.array([total_ms], np.float32), dtype=wp.float32, device=dev)
        comm.allreduce_max(tmax)
        total_ms = float(tmax.numpy()[0])
    value = nq * world * args.steps / (total_ms * 1e-3)

    # ---- traversal byte accounting (separate counted launch, not timed) -----------------------------
    with wp.query_stats() as st:
        wp.mesh_query_point_no_sign(mesh, q_dev, MAX_DIST, out=out)
        core.wp_cuda_context_synchronize(None)
    algo_bytes = nq * (12 + 13) + 64 * st.pair_fetches + 48 * st.tri_fetches
    kernel_ms = statistics.median(step_ms) if comm is None else None
    if comm is not None:  # kernel alone, without the gather
        kernel_ms = statistics.median([event_ms(core, lambda: wp.mesh_query_point_no_sign(mesh, q_dev, MAX_DIST, out=out), stream)
                                       for _ in range(3)])  # fmt: skip
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = None
    try:  # DRAM bytes of this kernel from the committed ncu --set full capture (same workload), scaled to this batch
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["k_query_point<no_sign>"]
        traffic = tj["dram_bytes_per_launch"] * nq / tj["queries"]
    except Exception:
        pass
    roofline = {
        "kernel": "k_query_point<no_sign>", "bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
        "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": algo_bytes, "launch_ms": kernel_ms,
        "pair_fetches_per_query": st.pair_fetches / nq, "tri_fetches_per_query": st.tri_fetches / nq,
        "nodes_per_s": 2 * st.pair_fetches / (kernel_ms * 1e-3),
        "note": "bytes = 25 B/query I/O + 64 B per sibling-pair fetch + 48 B per packed-triangle fetch (counted); "
                "SURVEY.md 8(d)'s traversal formula.  The tree (84 MB) and the packed triangles (63 MB) are L1 / L2 resident, "
                "so these fetches are served on chip: `traffic` (ncu, DRAM read + write of the same launch) is ~15x smaller, and "
                "`frac` -- fetched bytes against the HBM copy peak -- can exceed 1.  The kernel is issue bound "
                "(profiles/r01_summary.md: issue slots 70 % busy at 12 of 32 lanes), not bandwidth bound",
        "dram_frac": (traffic / (kernel_ms * 1e-3) / 1e9 / peak_gbs) if traffic else None,
    }  # fmt: skip

    # ---- end to end through the public API with pinned host buffers -------------------------------
    hp = Pinned(core, (nq, 3), np.float32)
    hp.array[:] = Qh
    h_out = wp.MeshQueryPoint(*(Pinned(core, (nq,), dt).array for dt in (np.uint8, np.float32, np.int32, np.float32, np.float32)))
    e2e_steps = max(3, min(args.steps, 5))
    wp.mesh_query_point_no_sign(mesh, hp.array, MAX_DIST, out=h_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        wp.mesh_query_point_no_sign(mesh, hp.array, MAX_DIST, out=h_out)
    core.wp_cuda_context_synchronize(None)
    e2e_s = time.perf_counter() - t0
    if comm is not None:
        tmax = wp.array(np.array([e2e_s], np.float32), dtype=wp.float32, device=dev)
        comm.allreduce_max(tmax)
        e2e_s = float(tmax.numpy()[0])
    e2e = {"value": nq * world * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 12 * nq, "d2h_bytes_per_step": 13 * nq,
           "steps": e2e_steps, "api": "warp_b200.mesh_query_point_no_sign(mesh, pinned numpy, ...) -> wp_b200_mesh_query_point_no_sign_host"}  # fmt: skip
    found = int(h_out.result.sum())

    # ---- side measurements: build / refit (C2) and rays (C3) ---------------------------------------
    extra = {"triangles": T, "first_build_ms_incl_alloc": first_build_ms, "queries_found": found}
    if not args.no_extra:
        extra.update(side_measurements(wp, core, mg, dev, stream, mesh, pts, P, I, peak_gbs, rank, comm, world))

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "triangles": T, "queries_per_gpu": nq, "leaf_size": 4, "max_dist": MAX_DIST,
                   "l2": "256 MB memset between timed steps; query inputs (201 MB) + outputs (218 MB) exceed the 126 MB L2; "
                         "the tree is meant to stay L2-resident",
                   "gather": ("ncclAllGather of result/face/u/v inside the step" if GATHER_PARTS == 1 else
                              f"result/face/u/v gathered inside the step, pipelined in {GATHER_PARTS} pieces") if comm else "none (1 GPU)"},
        "clocks": clocks.summary(), "e2e": e2e,
        # per step: Morton ordering of the batch (k_scene_bounds, k_morton_hist, 4 x k_onesweep_pass) + k_query_point
        "gpu_launches": args.steps * 7 * (GATHER_PARTS if comm is not None else 1), "roofline": roofline,
        "wall_ms_timed_region": wall_ms, "extra": extra,
    }  # fmt: skip
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(1 << 19)
    if rank == 0:
        print(json.dumps(line), flush=True)
    barrier()
    if comm is not None:
        comm.close()
    return 0


def side_measurements(wp, core, mg, dev, stream, mesh, pts, P, I, peak_gbs, rank, comm=None, world=1):
    """LBVH build + refit ms on C2 (with their HBM rooflines) and ray throughput on C3."""
    T = len(I) // 3
    out = {}
    # build: full constructor (allocations + descriptor upload + 8 launches), like timing wp.Mesh(...) on the reference
    idx_d = mesh.indices
    builds = []
    for _ in range(5):
        core.wp_cuda_context_synchronize(None)
        t0 = time.perf_counter()
        m2 = wp.Mesh(pts, idx_d, bvh_constructor="lbvh")
        core.wp_cuda_context_synchronize(None)
        builds.append(1e3 * (time.perf_counter() - t0))
        del m2
    out["build_ms_constructor"] = statistics.median(builds)
    # build kernels only: in-place rebuild of a Bvh over the same triangle boxes shares every kernel but the gather
    # refit: vertices re-noised in place, then mesh.refit()
    P2 = mg.renoise_sphere(P, 0.02, 3)
    refits = []
    for k in range(7):
        pts.assign(P2 if k % 2 == 0 else P)
        core.wp_cuda_context_synchronize(None)
        refits.append(event_ms(core, mesh.refit, stream))
    pts.assign(P)
    mesh.refit()
    out["refit_ms"] = statistics.median(refits[2:])
    out["refit_roofline"] = {"bound": "hbm", "algorithmic_bytes": 189 * T, "achieved": 189 * T / (out["refit_ms"] * 1e-3) / 1e9,
                             "peak": peak_gbs, "unit": "GB/s", "frac": 189 * T / (out["refit_ms"] * 1e-3) / 1e9 / peak_gbs}  # fmt: skip
    lib_build = build_kernel_ms(wp, core, stream, mesh)
    if lib_build is not None:
        out["build_ms_kernels"] = lib_build
        out["build_roofline"] = {"bound": "hbm", "algorithmic_bytes": 396 * T, "achieved": 396 * T / (lib_build * 1e-3) / 1e9,
                                 "peak": peak_gbs, "unit": "GB/s", "frac": 396 * T / (lib_build * 1e-3) / 1e9 / peak_gbs}  # fmt: skip
    # signed closest point (mesh_query_point: + three axis-probe traversals per query) on 2 M of the C2 queries
    try:
        ns = 1 << 21
        qs = wp.array(mg.box_queries(P, ns, seed=2), dtype=wp.vec3, device=dev)
        s_out = wp.MeshQueryPoint(*(wp.empty(ns, dt, dev) for dt in (wp.uint8, wp.float32, wp.int32, wp.float32, wp.float32)))
        run_s = lambda: wp.mesh_query_point(mesh, qs, MAX_DIST, out=s_out)  # noqa: E731
        run_s()
        ms = statistics.median([event_ms(core, run_s, stream) for _ in range(3)])
        out["signed"] = {"workload": "C2 mesh, 2 097 152 mesh_query_point (signed) queries", "queries_per_s": ns / (ms * 1e-3),
                         "ms": ms, "inside_fraction": float((s_out.sign.numpy() < 0).mean())}  # fmt: skip
    except Exception as e:
        out["signed"] = {"error": repr(e)}
    # rays: config C3 (10M-triangle heightfield, 4096 x 4096 primary rays)
    try:
        Ph, Ih = mg.heightfield(2237, 4)
        hm = wp.Mesh(wp.array(Ph, dtype=wp.vec3, device=dev), wp.array(Ih, dtype=wp.int32, device=dev), bvh_constructor="lbvh")
        S, D = mg.pinhole_rays(4096, 4096)
        s_d, d_d = wp.array(S, dtype=wp.vec3, device=dev), wp.array(D, dtype=wp.vec3, device=dev)
        n = S.shape[0]
        r_out = wp.MeshQueryRay(*(wp.empty(n, dt, dev) for dt in (wp.uint8, wp.float32, wp.int32, wp.float32, wp.float32, wp.float32, wp.vec3)))
        run = lambda: wp.mesh_query_ray(hm, s_d, d_d, 1.0e6, out=r_out)  # noqa: E731
        run()
        ms = statistics.median([event_ms(core, run, stream) for _ in range(3)])
        with wp.query_stats() as st:
            run()
            core.wp_cuda_context_synchronize(None)
        out["rays"] = {"workload": "C3: 9 999 392-triangle heightfield, 4096x4096 pinhole rays", "rays_per_s": n / (ms * 1e-3),
                       "ms": ms, "hit_fraction": float(r_out.result.numpy().mean()),
                       "pair_fetches_per_ray": st.pair_fetches / n, "tri_fetches_per_ray": st.tri_fetches / n,
                       "bytes_fetched_GBps": (n * (24 + 37) + 64 * st.pair_fetches + 48 * st.tri_fetches) / (ms * 1e-3) / 1e9}  # fmt: skip
        if comm is not None:
            # rays at N GPUs (weak scaling): every rank traces its own 4096 x 4096 image of the replicated terrain
            # (eye shifted per rank) and the seven result fields are all-gathered into global ray order
            from warp_b200.distributed import ShardPlan, sharded_query_ray

            S2, D2 = mg.pinhole_rays(4096, 4096, eye=(0.5 + 0.02 * rank, -0.6, 0.9))
            s_d.assign(S2), d_d.assign(D2)
            plan = ShardPlan(n * world, world)
            dts = {"result": wp.uint8, "sign": wp.float32, "face": wp.int32, "t": wp.float32, "u": wp.float32,
                   "v": wp.float32, "normal": wp.vec3}
            g_out = {k: wp.empty(plan.padded, dt, dev) for k, dt in dts.items()}
            run_n = lambda: sharded_query_ray(hm, s_d, d_d, plan, 1.0e6, comm, rank, local_out=r_out, global_out=g_out)  # noqa: E731
            run_n()
            comm.barrier()
            core.wp_cuda_context_synchronize(None)
            ms_n = statistics.median([event_ms(core, run_n, stream) for _ in range(3)])
            tmax = wp.array(np.array([ms_n], np.float32), dtype=wp.float32, device=dev)
            comm.allreduce_max(tmax)
            ms_n = float(tmax.numpy()[0])
            out["rays_sharded"] = {"workload": "C3 terrain replicated, 4096x4096 rays per GPU, 7 fields all-gathered (NCCL)",
                                   "n_gpus": world, "rays_per_s": n * world / (ms_n * 1e-3), "ms": ms_n,
                                   "hits_all_ranks": int(g_out["result"].numpy().sum())}  # fmt: skip
    except Exception as e:  # the headline must not die on a side measurement
        out["rays"] = {"error": repr(e)}
    # collision loop: config C4 (4 M-triangle deforming cloth; per frame = refit() + 8.4 M closest-point queries within 0.05)
    try:
        try:
            del hm, s_d, d_d, r_out
        except NameError:
            pass
        nc, nqc = 1415, 1 << 23
        Pc, Ic = mg.cloth(nc, 0)
        cpts = wp.array(Pc, dtype=wp.vec3, device=dev)
        cm = wp.Mesh(cpts, wp.array(Ic, dtype=wp.int32, device=dev), bvh_constructor="lbvh")
        c_out = wp.MeshQueryPoint(*(wp.empty(nqc, dt, dev) for dt in (wp.uint8, wp.float32, wp.int32, wp.float32, wp.float32)))
        frames, refits, prev = [], [], Pc
        for f in range(1, 7):
            Pf, _ = mg.cloth(nc, f)
            rng = np.random.default_rng(5 + f)
            Qc = (prev[rng.integers(0, prev.shape[0], nqc)] + rng.normal(0, 0.01, (nqc, 3))).astype(np.float32)
            qc = wp.array(Qc, dtype=wp.vec3, device=dev)
            cpts.assign(Pf)  # vertex update (not timed: a simulation writes the positions on the device)
            core.wp_cuda_context_synchronize(None)

            def frame():
                cm.refit()
                wp.mesh_query_point_no_sign(cm, qc, 0.05, out=c_out)

            frames.append(event_ms(core, frame, stream))
            refits.append(event_ms(core, cm.refit, stream))
            prev = Pf
            del qc
        out["cloth"] = {"workload": "C4: 3 998 792-triangle cloth, per frame refit() + 8 388 608 mesh_query_point_no_sign within 0.05",
                        "frame_ms": statistics.median(frames[1:]), "refit_ms": statistics.median(refits[1:]),
                        "queries_per_s": nqc / (statistics.median(frames[1:]) * 1e-3), "frames": len(frames),
                        "found_fraction": float(c_out.result.numpy().mean())}  # fmt: skip
    except Exception as e:
        out["cloth"] = {"error": repr(e)}
    return out


def build_kernel_ms(wp, core, stream, mesh):
    """Device time of the build kernels alone (no allocation): wp_b200_mesh_rebuild_device if exported."""
    fn = getattr(core, "wp_b200_mesh_rebuild_device", None)
    if fn is None:
        return None
    fn.restype, fn.argtypes = ctypes.c_int, [ctypes.c_uint64]
    fn(mesh.id)
    core.wp_cuda_context_synchronize(None)
    return statistics.median([event_ms(core, lambda: fn(mesh.id), stream) for _ in range(7)])


if __name__ == "__main__":
    sys.exit(main())
