#!/usr/bin/env python
"""Benchmark of the cluster-LOD DAG build (BASELINE.json metric: Mtris/s, full DAG build).

  python bench.py --gpus N --steps K --warmup W            this framework (CUDA, one process per GPU)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU implementation on the host cores

A step is one full DAG build of the workload mesh. N = 1: config C3, the 100 M-triangle noisy displaced icosphere with UV
seams, the mesh the >= 50x target of BASELINE.json is quoted on and the largest single-GPU configuration
(CLODB200_BENCH_WORKLOAD=C2 selects the 10 M-triangle heightfield, =C4 the scene batch). Under torchrun (N > 1) every rank
builds its own mesh of that class (independent meshes sharded by mesh: weak scaling); there is no data-path collective, only
the gather of the per-mesh cache metadata blobs (the library's own NCCL all-gather on a side stream, csrc/comm.cu), the
barrier and the max-over-ranks of the timing. The line also carries, at N = 1, the C2 measurement ("also"), a parity block
(our DAG next to the reference's on the CPU arm's sample mesh) and the CPU baseline.

A build is the reference's outer builder call (BuildClusterLODArtifactsFromGeometry, ClusterLODUtilities.cpp:5325, mesh
mode): DAG to a single root cluster + group tables + traversal hierarchy + mesh-wide page packing, finished
ClusterLODPrebuiltData and page blobs in host memory. `value` is timed with the geometry resident in HBM
(clodb200_geometryBuildArtifacts); `e2e` goes through the host-pointer C ABI (clodb200_buildArtifacts: upload + build +
page read-back) from pinned host buffers.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from basicrenderer_b200 import meshgen  # noqa: E402

WORKLOAD_N = 2236  # (n+1)^2 grid => 9 999 392 triangles
ATTR_WEIGHTS = np.ones(3, np.float32)  # normals x3, weight 1.0 (ClusterLODUtilities.cpp:5368-5385)
PROTECT_MASK = 7

# Algorithmic bytes per thread of the kernels that can dominate a step (DESIGN.md "Kernels" table derives each figure):
# what the kernel must read and write once if every operand moved exactly once.
KERNEL_BYTES_PER_THREAD = {
    # segmented SAH area scan, 8 elements per thread (SA_ITEMS): order u32 + gathered box 32 B + node id u32 + area f32 out
    "k_sa_chained": 8 * (4 + 32 + 4 + 4),
    # radix sort scatter, 8 keys per thread: key in + key out + value in + value out (u32 keys)
    "(k_rs_scatter<K>)": 8 * 16,
    "k_rs_scatter<K>": 8 * 16,
    "(k_rs_hist<K>)": 8 * 4,
    "k_rs_hist<K>": 8 * 4,
    # pivot search, 4 positions per thread: node id + prefix area + suffix area (node tables stay in cache)
    "k_pivot_large": 4 * (4 + 4 + 4),
    "k_partition": 4 + 4 + 4 + 1 + 4 + 4,
    # fused side flags + segmented count + stable partition, 8 positions per thread: order + node id + side byte in, order out
    "k_partition_chained": 8 * (4 + 4 + 1 + 4),
    "k_side_flags": 4 + 1 + 4,
    "k_mark_sides": 4 + 4 + 1 + 8,
    "k_update_node_of_pos": 4 + 4 + 4,
    # persistent wavefront kernel: the profile reports (candidate x round) work items instead of threads; per item one publish
    # and one decide step: sorted id 4 + (v0, v1) 8 + remap 8 + three 8-byte vertex minima
    "k_wave_rounds": 4 + 8 + 8 + 3 * 8,
    "k_wave_decide": 1 + 4 + 8 + 8 + 16,
    "k_wave_publish": 1 + 4 + 8 + 3 * 8,
    "k_rank": 9 + 2 * (44 + 12) + 4 + 8,
    "k_build_clusters": 128 * 3 * 4 * 2 + 4 * 4,
    "k_cluster_bounds": 128 * 3 * 16 + 16,
    # chained single-pass scan, 8 u32 elements per thread, read once + written once
    "(k_scan_chained<T, Op>)": 8 * 8,
    "(k_scan_chained<T, Op, SCAN_THREADS, SCAN_ITEMS>)": 8 * 8,
    "(k_scan_chained<T, Op, SCAN_LARGE_THREADS, SCAN_LARGE_ITEMS>)": 16 * 8,
    # per candidate: sorted id 4 + status 1 + group 4 + error 4 + the three prefix arrays 4 + 4 + 8
    "k_cut_find": 29,
    "k_cut_inputs": 4 + 1 + 4 + 4 + 4 + 1 + 4 + 4 + 8,
    "k_cut_apply": 4 + 1 + 4 + 4,
    # listing pass, one thread per examined position: status 1 + window offsets (cached) + for listed candidates sorted id 4, endpoints 8,
    # position classes 8, record 16, three publishes 24 (about 40 % of the positions are listed)
    "k_wave_list": 1 + 24,
    "k_pick_flags": 4 + 4 + 4 + 4 + 2 + 4,
    "k_pick_emit": 4 + 4 + 4,
    "k_adj_count": 4 + 4,
    "k_adj_fill": 4 + 4 + 4 + 4,
}

# Kernels without a byte model are not bandwidth kernels; what bounds them, from the ncu captures in profiles/r02_ncu_kernels.md
KERNEL_LIMITER = {
    "k_update_quadrics": "latency: one thread per accepted collapse (compacted list), each a chain of dependent scattered read-modify-writes of 44..156-byte quadrics",
    "k_fill_attribute_quadrics": "gather + compute: per vertex, the attribute quadrics and gradients of its adjacent triangles in corner order (the reference's summation order)",
    "k_merge_rounds": "latency + grid barriers: persistent cooperative kernel, 6 barriers per merge round over a shrinking edge list (barrier 55 warps per issue)",
    "k_build_clusters_warp": "compute: meshlet assembly + optimizeMeshlet per warp, SM throughput 88 %",
    "k_cluster_bounds_warp": "compute: 7-axis extremal search + sequential sphere growth per cluster",
    "k_leaf_test_warp": "latency: one warp per tree node, small launches",
    "k_resolve_nodes": "launch latency: a few thousand nodes per launch",
}


# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (profiles/), for the
# largest launch of the kernel on this workload; compare with roofline.algorithmic_bytes_per_launch of the same launch size
KERNEL_DRAM_TRAFFIC = {}
KERNEL_DRAM_TRAFFIC_SOURCE = None
try:
    _t = json.load(open(os.path.join(ROOT, "profiles", "kernel_dram_traffic.json")))
    KERNEL_DRAM_TRAFFIC = _t.get("kernels", {})
    KERNEL_DRAM_TRAFFIC_SOURCE = _t.get("source")
except Exception:
    pass


# dram__bytes_read.sum + dram__bytes_write.sum summed over every kernel of ONE build step (one `ncu --metrics` pass, committed under
# profiles/): {"C2": {"bytes": ..., "source": ...}, ...}
WHOLE_STEP_DRAM = {}
try:
    WHOLE_STEP_DRAM = json.load(open(os.path.join(ROOT, "profiles", "whole_step_dram.json")))
except Exception:
    pass


def _kernel_table(rows, peak_gbs, limit=10):
    """Top kernels of one profiled step: time, launches and, where DESIGN.md has a byte model for the kernel, the algorithmic
    GB/s and its fraction of the measured HBM peak. rows: (kernel, launches, total_ms, total_threads_or_items)."""
    out = []
    for name, launches, ms, threads in rows[:limit]:
        entry = {"kernel": name, "launches": int(launches), "ms": round(float(ms), 3)}
        bpt = KERNEL_BYTES_PER_THREAD.get(name)
        if bpt and ms > 0:
            gbs = bpt * threads / (ms * 1e-3) / 1e9
            entry["algorithmic_gbs"] = round(gbs, 1)
            entry["frac"] = round(gbs / peak_gbs, 4)
        elif name in KERNEL_LIMITER:
            entry["limiter"] = KERNEL_LIMITER[name]
        out.append(entry)
    return out


def _sample_clocks(stop_event, out):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    idx = os.environ.get("LOCAL_RANK", "0")
    while not stop_event.is_set():
        try:
            r = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", idx], capture_output=True, text=True, timeout=5)
            parts = [p.strip() for p in r.stdout.strip().split(",")]
            if len(parts) >= 6:
                out.append(parts)
        except Exception:
            pass
        stop_event.wait(0.5)


def _clock_summary(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
    reasons = set()
    for s in samples:
        for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
            if v.lower().startswith("active"):
                reasons.add(name)
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(samples[0][1]) if samples[0][1].isdigit() else None, "reasons": sorted(reasons)}


def _dist():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


WORKLOAD = os.environ.get("CLODB200_BENCH_WORKLOAD", "C3").upper()  # C3 (default) | C2 | C4 (scene batch, see run_scene_batch)
C3_F = 2236  # icosphere frequency => 99 993 920 triangles
METRIC = "Mtris/s full cluster-LOD DAG build"


def _workload_text(workload: str, world: int) -> str:
    if workload == "C3":
        f = int(os.environ.get("CLODB200_BENCH_ICO_F", C3_F))
        t = f"C3: {20 * f * f}-triangle noisy displaced icosphere (f={f}) with normals + per-face UV atlas seams, 7 simplification attributes (normal + tangent xyz + sign), full DAG to a single root cluster; MikkTSpace tangents generated inside the timed call, as in the reference"
    else:
        n = int(os.environ.get("CLODB200_BENCH_GRID", WORKLOAD_N))
        t = f"C2: {2 * n * n}-triangle displaced heightfield grid (n={n}), pos+normal, full DAG to a single root cluster"
    return t + ("" if world == 1 else f"; one such mesh per GPU ({world} independent meshes with distinct noise seeds, sharded by mesh)")


def _config(workload: str, world: int) -> dict:
    """Names the workload; identical in both arms (`--impl reference` times the reference's CPU implementation on this config)."""
    return {
        "workload": _workload_text(workload, world),
        "builder": "clodDefaultConfig(128) + BasicRenderer overrides (128/128/64, partition 384, 8 refined ids, spatial clusters)",
        "scope": "BuildClusterLODArtifactsFromGeometry, mesh mode: remap, clusterize, partition, lock, simplify, bounds, error rule, group tables, traversal hierarchy, mesh page packing, page blobs in host memory",
        "l2": "inputs and per-level working sets (>= 240 MB) exceed the 126 MB L2; no explicit flush",
    }


def _make_mesh(workload: str, rank: int, world: int):
    """The rank's mesh: (vertices [V, stride/4] f32, indices u32, flags)."""
    from basicrenderer_b200 import artifacts as art

    if workload == "C3":
        f = int(os.environ.get("CLODB200_BENCH_ICO_F", C3_F))
        # the generator's analytic tangents are not used: the MikkTSpace tangent stream is generated inside every build call
        # (csrc/mikk.cu), as the reference generates it inside BuildClusterLODArtifactsFromGeometry
        mesh, _analytic_tangents = meshgen.icosphere_seams_torch(f, seed=42 + rank)
        import torch

        torch.cuda.empty_cache()
        return mesh.vertices, mesh.indices, art.VERTEX_NORMALS | art.VERTEX_TEXCOORDS, mesh.triangle_count
    seed = 1234 if world == 1 else 1234 + rank
    n = int(os.environ.get("CLODB200_BENCH_GRID", WORKLOAD_N))
    m = meshgen.grid(n, seed=seed)
    return art.interleave(m.positions, m.normals), m.indices, art.VERTEX_NORMALS, m.triangle_count


def _pin(a: np.ndarray) -> np.ndarray:
    import torch

    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    _PINNED.append(t)  # keep the pinned allocation alive for the numpy view
    return t.numpy()


_PINNED = []


def _share_host_cores(world: int, builds_in_flight: int = 1):
    """Several ranks (and several builds in flight per rank) share the box's host cores: cap the host threads each build's replay
    stages may use (csrc/rt.cuh host_parallel_for) so that the processes do not oversubscribe them. Must run before the library loads."""
    cores = os.cpu_count() or 1
    os.environ.setdefault("CLODB200_HOST_THREADS", str(max(1, min(16, cores // max(1, world * builds_in_flight)))))


def _comm_setup(lib, rank, world):
    """The library's own NCCL communicator for the metadata gather: rank 0 makes the id, torch.distributed only ships it."""
    if world == 1:
        return
    import torch.distributed as dist

    box = [lib.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    lib.comm_init(box[0], world, rank)
    # warm-up exchanges: NCCL sets its channels up lazily inside the first collective of a communicator (hundreds of ms), and again
    # when a message is large enough to change protocol; the batch gather carries a few MB per rank
    for size in (16, 1 << 20, 24 << 20):
        lib.gather_end(lib.gather_begin(bytes(size)))


def _timed_builds(lib, steps, build_once, rank, world):
    """K builds. With N > 1 the K steps are one scene batch of K meshes per rank: every build's cache metadata blob is kept and
    the batch ends with ONE gather of all blobs to every rank (the library's NCCL all-gather, csrc/comm.cu), inside the timed
    region. (A gather per step would make the ranks run in lock step; a gather left pending across builds parks an NCCL kernel
    on the GPU that waits for the slowest peer and keeps this rank's cooperative kernels from becoming co-resident.)
    Returns (device ms, last record)."""
    from basicrenderer_b200 import sharding

    blobs = []
    lib.timer_start()
    rec = None
    for s in range(steps):
        rec = build_once(world > 1)
        if world > 1:
            blobs.append(lib.serialize_metadata(rec, f"clod_mesh{rank}_{s}.clodbin", "bench", f"/mesh{rank}_{s}"))
            lib.free_artifacts(rec)
    if world > 1:
        gathered = sharding.gather_metadata_end(lib, sharding.gather_metadata_begin(lib, [rank * steps + s for s in range(steps)], blobs))
        assert len(gathered) == world * steps
    return lib.timer_stop_ms(), rec


def _measure(lib, workload, rank, world, steps, warmup, barrier, with_clocks=True):
    """Resident-input loop (`value`) and host-buffer loop (`e2e`) on this rank's mesh of `workload`."""
    vertices, indices, flags, T = _make_mesh(workload, rank, world)
    vertices = _pin(vertices)
    indices = _pin(indices)
    handle = lib.upload_geometry(vertices, indices, flags)
    for _ in range(warmup):
        rec = lib.build_artifacts_resident(handle, views=True)
    stop = threading.Event()
    clock_samples = []
    sampler = threading.Thread(target=_sample_clocks, args=(stop, clock_samples), daemon=True)
    with_clocks = with_clocks and rank == 0  # one nvidia-smi poller per box: NVML queries take driver locks the builds also need
    if with_clocks:
        sampler.start()
    barrier()
    launches0 = lib.launch_count
    ms, rec_last = _timed_builds(lib, steps, lambda keep: lib.build_artifacts_resident(handle, views=True, keep_handle=keep), rank, world)
    launches = lib.launch_count - launches0
    barrier()
    if world > 1:  # the gathered runs freed their records; one more (untimed) build for the output statistics
        rec_last = lib.build_artifacts_resident(handle, views=True)
    stat = dict(rec_last.stat)
    # ---- end to end through the host-pointer C ABI (upload + build + page read-back every step)
    for _ in range(min(warmup, 2)):
        lib.build_artifacts(vertices, indices, flags, views=True)
    barrier()
    ms_e2e, rec_e2e = _timed_builds(lib, steps, lambda keep: lib.build_artifacts(vertices, indices, flags, views=True, keep_handle=keep), rank, world)
    barrier()
    if with_clocks:
        stop.set()
        sampler.join()
    d2h = int(stat["d2h_bytes"])
    return {"T": T, "ms": ms, "ms_e2e": ms_e2e, "launches": launches, "stat": stat, "h2d": vertices.nbytes + indices.nbytes, "d2h": d2h, "clocks": clock_samples,
            "handle": handle}


def _batch_shape():
    count = int(os.environ.get("CLODB200_BENCH_MESHES", 4096 if WORKLOAD == "C5" else 512))
    total = float(os.environ.get("CLODB200_BENCH_BATCH_TRIS", 8.0e9 if WORKLOAD == "C5" else 1.0e9))
    return count, total


def _batch_config(world: int) -> dict:
    count, total = _batch_shape()
    c = _config("C2", 1)
    c["workload"] = (f"{WORKLOAD if WORKLOAD in ('C4', 'C5') else 'C4'}: scene batch of {count} independent synthetic meshes (sphere / heightfield / torus by index, pos+normal, log-uniform triangle "
                     f"budgets rescaled to {total:.3g} triangles in total), every mesh built to a single root cluster, sharded by mesh across the GPUs (LPT), one metadata gather per batch")
    c["l2"] = "meshes are built back to back, several in flight; the shard (>= 3 GB per GPU) exceeds the 126 MB L2"
    return c


def _reference_batch_step(meshes, cores):
    """The CPU arm's scene-batch step: the meshes are the primitives of one file - outer parallel-for over meshes
    (GlTFGeometryExtractor.cpp:1349), each mesh built by the reference's builder with its group-level parallel-for."""
    from concurrent.futures import ThreadPoolExecutor

    inner = max(1, cores // max(1, min(len(meshes), cores)))
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=min(len(meshes), cores)) as ex:
        list(ex.map(lambda m: _reference_build(m, inner, "C2"), meshes))
    return time.perf_counter() - t0


def run_reference_batch(args):
    """Bounded sample of the scene batch for the CPU arm: 2 x cores meshes of the batch's own generator and budget distribution
    (every k-th mesh of the batch, budgets scaled so that the sample holds CLODB200_REF_BATCH_TRIS triangles, default 6 M)."""
    count, total = _batch_shape()
    cores = os.cpu_count() or 1
    budgets = meshgen.scene_batch_sizes(count, total)
    take = min(count, 2 * cores)
    ids = [int(round(k * (count - 1) / max(1, take - 1))) for k in range(take)] if take > 1 else [0]
    scale = float(os.environ.get("CLODB200_REF_BATCH_TRIS", 6.0e6)) / float(sum(budgets[i] for i in ids))
    meshes = [meshgen.scene_mesh(i, max(2000.0, budgets[i] * scale)) for i in ids]
    for _ in range(args.warmup):
        _reference_batch_step(meshes, cores)
    dt = sum(_reference_batch_step(meshes, cores) for _ in range(args.steps)) / max(1, args.steps)
    tris = sum(m.triangle_count for m in meshes)
    value = tris / dt / 1e6
    sample = (f"{len(meshes)} meshes of the batch generator (every {max(1, count // take)}-th mesh, budgets scaled to {tris} triangles in total), built concurrently by the unmodified "
              f"BuildClusterLODArtifactsFromGeometry (mesh mode): outer parallel-for over meshes on {min(len(meshes), cores)} threads, {dt:.2f} s per step")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mtris/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+u32", "data": "synthetic", "config": _batch_config(max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": "Mtris/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "Mtris/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_scene_batch(args):
    """CLODB200_BENCH_WORKLOAD=C4 | C5 (opt-in): the scene batch configurations of BASELINE.json. C4: 512 independent meshes of the
    scene generator (type = i mod 3 in {sphere, heightfield, torus}, log-uniform triangle budgets in [1e4, 2e6] rescaled to
    1.0 B triangles in total); C5: 4 096 meshes, 8.0 B triangles. The meshes are sharded by mesh across the ranks (LPT,
    basicrenderer_b200/sharding.py), every rank generates its shard on its GPU, keeps several builds in flight (one host thread
    and build context each) and the batch ends with one NCCL gather of the cache metadata blobs. Total work is fixed as N grows
    (strong scaling). CLODB200_BENCH_MESHES / CLODB200_BENCH_BATCH_TRIS override the batch size."""
    import torch
    import torch.distributed as dist

    rank, world, local = _dist()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    from basicrenderer_b200 import artifacts as art
    from basicrenderer_b200 import load, sharding

    workers = max(1, int(os.environ.get("CLODB200_BENCH_THREADS", 4)))
    _share_host_cores(world, workers)
    lib = load(local)
    _comm_setup(lib, rank, world)
    count, total = _batch_shape()
    budgets = meshgen.scene_batch_sizes(count, total)
    mine = sharding.assign_meshes([int(b) for b in budgets], world)[rank]
    meshes = [meshgen.scene_mesh_torch(i, budgets[i]) for i in mine]
    torch.cuda.empty_cache()
    pin = total <= 2.0e9  # the 8 B-triangle aggregate keeps its host copies pageable (24 GB per rank at N = 8)
    host = [((_pin(m.vertices), _pin(m.indices)) if pin else (m.vertices, m.indices)) for m in meshes]
    handles = [lib.upload_geometry(v, i, art.VERTEX_NORMALS) for v, i in host]
    my_tris = sum(m.triangle_count for m in meshes)
    del meshes

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # several builds in flight per GPU: every worker thread owns a build context (stream, arenas) inside the library
    # largest meshes first, workers take the next mesh when they finish one (the way the reference's parallel-for hands out primitives)
    order = sorted(range(len(mine)), key=lambda k: -int(host[k][1].size))
    import queue
    from concurrent.futures import ThreadPoolExecutor

    pool = ThreadPoolExecutor(max_workers=workers)

    lane_launches = []

    def build_lane(todo, resident):
        out = []
        l0 = lib.launch_count  # per calling thread
        while True:
            try:
                k = todo.get_nowait()
            except queue.Empty:
                break
            rec = lib.build_artifacts_resident(handles[k], views=True, keep_handle=True) if resident else lib.build_artifacts(host[k][0], host[k][1], art.VERTEX_NORMALS, views=True, keep_handle=True)
            out.append((mine[k], lib.serialize_metadata(rec, f"clod_mesh{mine[k]}.clodbin", "bench", f"/mesh{mine[k]}")))
            lib.free_artifacts(rec)
        lane_launches.append(lib.launch_count - l0)
        return out

    def step(resident):
        todo = queue.SimpleQueue()
        for k in order:
            todo.put(k)
        done = [r for lane_out in pool.map(lambda _w: build_lane(todo, resident), range(workers)) for r in lane_out]
        done.sort()
        # one gather per batch, through the library's own NCCL all-gather (csrc/comm.cu)
        gathered = sharding.gather_metadata_end(lib, sharding.gather_metadata_begin(lib, [i for i, _ in done], [b for _, b in done]))
        assert len(gathered) == count
        return sum(len(b) for b in gathered.values())

    for _ in range(args.warmup):
        step(True)
    stop = threading.Event()
    clock_samples = []
    sampler = threading.Thread(target=_sample_clocks, args=(stop, clock_samples), daemon=True)
    if rank == 0:
        sampler.start()
    barrier()
    lane_launches.clear()
    # the builds run on several streams: the timed region is bracketed by device-wide synchronisations and read on the host
    t0 = time.perf_counter()
    for _ in range(args.steps):
        blob_bytes = step(True)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    launches = sum(lane_launches)
    barrier()
    stop.set()
    if rank == 0:
        sampler.join()
    step(False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(False)
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    barrier()
    tris = float(my_tris)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        tt = torch.tensor([tris], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        tris = float(tt[0])
    if rank == 0:
        ms_per_step = ms / args.steps
        print(json.dumps({
            "metric": "Mtris/s full cluster-LOD DAG build", "value": tris / (ms_per_step * 1e-3) / 1e6, "unit": "Mtris/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+u32", "data": "synthetic",
            "config": _batch_config(world),
            "output": {"meshes": count, "triangles": int(tris), "mesh_triangles_min": int(budgets.min()), "mesh_triangles_max": int(budgets.max()), "builds_in_flight_per_gpu": workers, "metadata_bytes_gathered": blob_bytes,
                       "timing": "host clock between device-wide synchronisations (several build streams per GPU), max over ranks"},
            "clocks": _clock_summary(clock_samples),
            "e2e": {"value": tris / (ms_e2e / args.steps * 1e-3) / 1e6, "unit": "Mtris/s", "h2d_bytes_per_step": int(sum(v.nbytes + i.nbytes for v, i in host)), "d2h_bytes_per_step": None},
            "gpu_launches": int(launches),
        }))
    for h in handles:
        lib.free_geometry(h)
    if world > 1:
        dist.barrier()
        lib.comm_destroy()
        dist.destroy_process_group()


def run_ours(args):
    import torch
    import torch.distributed as dist

    if WORKLOAD in ("C4", "C5"):
        return run_scene_batch(args)
    rank, world, local = _dist()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)

    from basicrenderer_b200 import load

    _share_host_cores(world)
    lib = load(local)
    _comm_setup(lib, rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    m = _measure(lib, WORKLOAD, rank, world, args.steps, args.warmup, barrier)
    ms, ms_e2e, T = m["ms"], m["ms_e2e"], m["T"]

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        tt = torch.tensor([T], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        total_tris = float(tt[0])
    else:
        total_tris = float(T)

    result = None
    if rank == 0:
        # ---- per-kernel breakdown of one more step, CUDA events on the build stream (not part of the timed region)
        lib.profile_enable(True)
        lib.build_artifacts_resident(m["handle"], views=True)
        rows = lib.profile_report()
        lib.profile_enable(False)
        total_kernel_ms = sum(r[2] for r in rows)
        if os.environ.get("CLODB200_PROFILE_OUT"):
            with open(os.environ["CLODB200_PROFILE_OUT"], "w") as f:
                f.write("kernel,launches,total_ms,total_threads\n")
                for r in rows:
                    f.write(f"{r[0]},{r[1]},{r[2]:.4f},{r[3]}\n")
        top = rows[0]
        peaks = {}
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peaks = json.load(open(peaks_path))
        peak = float(peaks.get("hbm_gbs", 6650.0))
        bpt = KERNEL_BYTES_PER_THREAD.get(top[0])
        achieved = (bpt * top[3] / (top[2] * 1e-3) / 1e9) if bpt else None
        traffic = KERNEL_DRAM_TRAFFIC.get(top[0])
        roofline = {
            "bound": "hbm",
            "kernel": top[0],
            "achieved": achieved,
            "peak": peak,
            "peak_source": "MEASURED_PEAKS.json (measured copy bandwidth)" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)",
            "unit": "GB/s",
            "frac": (achieved / peak) if achieved else None,
            "traffic": traffic.get(WORKLOAD) if isinstance(traffic, dict) else traffic,
            "traffic_source": KERNEL_DRAM_TRAFFIC_SOURCE if top[0] in KERNEL_DRAM_TRAFFIC else None,
            "algorithmic_bytes_per_launch": (bpt * top[3] / max(1, top[1])) if bpt else None,
            "kernel_share_of_step": top[2] / total_kernel_ms if total_kernel_ms else None,
            "launches": top[1],
            "avg_launch_us": top[2] * 1e3 / max(1, top[1]),
            "whole_step_dram": WHOLE_STEP_DRAM.get(WORKLOAD),
            "top_kernels": _kernel_table(rows, peak, limit=15),
        }
        lib.free_geometry(m["handle"])
        m["handle"] = None

        ms_per_step = ms / args.steps
        stat = m["stat"]
        result = {
            "metric": METRIC,
            "value": total_tris / (ms_per_step * 1e-3) / 1e6,
            "unit": "Mtris/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms_per_step,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f32+u32",
            "data": "synthetic",
            "config": _config(WORKLOAD, world),
            "output": {k: stat[k] for k in ("levels", "groups", "meshlets", "segments", "pages", "page_bytes")},
            "whole_job_roofline": _whole_job_roofline(stat, T, ms / args.steps, peak, 24 if WORKLOAD != "C3" else 48),
            "clocks": _clock_summary(m["clocks"]),
            "e2e": {"value": total_tris / (ms_e2e / args.steps * 1e-3) / 1e6, "unit": "Mtris/s", "h2d_bytes_per_step": int(m["h2d"]), "d2h_bytes_per_step": m["d2h"]},
            "gpu_launches": int(m["launches"]),
            "roofline": roofline,
        }
        if world == 1:
            if WORKLOAD == "C3" and not os.environ.get("CLODB200_BENCH_SKIP_ALSO"):
                # the 10 M-triangle heightfield (BASELINE.json configs[1]) in the same run, short loop
                c2 = _measure(lib, "C2", 0, 1, 5, 3, barrier, with_clocks=False)
                lib.free_geometry(c2["handle"])
                result["also"] = {"config": _config("C2", 1)["workload"], "steps": 5, "warmup": 3, "ms_per_step": c2["ms"] / 5, "value": c2["T"] / (c2["ms"] / 5 * 1e-3) / 1e6,
                                  "e2e": c2["T"] / (c2["ms_e2e"] / 5 * 1e-3) / 1e6, "unit": "Mtris/s", "groups": c2["stat"]["groups"], "meshlets": c2["stat"]["meshlets"], "gpu_launches_per_step": c2["launches"] // 5}
            result["parity"] = parity_block(lib)
            result["cpu_baseline"] = cpu_baseline_sample()
    if m.get("handle"):
        lib.free_geometry(m["handle"])
    if world > 1:
        dist.barrier()
        lib.comm_destroy()
        dist.destroy_process_group()
    if result is not None:
        print(json.dumps(result))


def _whole_job_roofline(stat, T0, ms_per_step, peak_gbs, s_v=24):
    """compulsory bytes of the whole build from the builder's own output sums,
    B = 24*sum T_l + S_v*sum V_l + 12*sum T_{l+1} + page bytes (S_v = 24 B for pos+normal, 48 B with uv + tangent4)."""
    b = 24 * stat["level_triangles"] + s_v * stat["group_vertices"] + 12 * stat["simplified_triangles"] + stat["page_bytes"]
    gbs = b / (ms_per_step * 1e-3) / 1e9
    return {"algorithmic_bytes": int(b), "bytes_per_input_triangle": b / T0, "achieved_gbs": gbs, "frac": gbs / peak_gbs}


# ---- CPU arm: the compiled, unmodified reference on the host cores ------------------------------------------------------------
def _reference_sample_mesh(workload: str = None, index: int = 0):
    """Bounded sample of the workload: a 1 M-triangle mesh of the same generator (C3: f = 224 icosphere with the same displacement,
    normals and UV atlas seams; C2: a tile of the same heightfield at the same density). The reference is ~80 % serial per mesh and
    gets SLOWER per triangle on larger meshes (10 M-triangle C2: 0.226 vs 0.241 Mtris/s on 8 cores), so the sample flatters it."""
    workload = workload or WORKLOAD
    if workload == "C3":
        return meshgen.icosphere(int(os.environ.get("CLODB200_REF_ICO_F", 224)), True, True, seed=42 + index)
    n = int(os.environ.get("CLODB200_REF_GRID", 707))
    return meshgen.grid(n, seed=1234 + index)


def _reference_build(m, threads, workload: str = None):
    """One call of the reference's own builder (oracle/_ref/libclodref_full.so: the unmodified
    BuildClusterLODArtifactsFromGeometry with its own clodBuildEx, mesh mode) on host arrays; returns seconds inside the call."""
    from oracle import clodfull

    if (workload or WORKLOAD) == "C3":
        return clodfull.build(m.vertices, m.indices, flags=m.flags, threads=threads).seconds
    v = clodfull.interleave(m.positions, m.normals)
    return clodfull.build(v, m.indices, flags=clodfull.VERTEX_NORMALS, threads=threads).seconds


def _reference_sample_text(m, threads, dt, meshes=1):
    what = (f"{m.triangle_count}-triangle icosphere of the C3 generator (f=224, same displacement, normals and UV atlas seams), incl. the reference's MikkTSpace tangent generation"
            if WORKLOAD == "C3" else f"{m.triangle_count}-triangle tile of the C2 heightfield (same generator and density)")
    fan = "" if meshes == 1 else f"; {meshes} such meshes built concurrently (outer parallel-for over meshes, GlTFGeometryExtractor.cpp:1349; inner parallel-for over groups), "
    return (f"{what}{fan}; unmodified BuildClusterLODArtifactsFromGeometry in mesh mode (meshoptimizer v1.0 + clusterlod.h + L3 builder, std::thread pool of {threads} per mesh "
            f"in place of oneTBB), {dt:.2f} s per step; includes the reference's unused VoxelSourceTriangleBVH::Build (SURVEY.md §8a a20)")


def _reference_step(meshes, cores):
    """One step of the CPU arm: every mesh of the step built by the reference, all of them concurrently when there are several
    (the way an importer fans the primitives of one file out over the worker threads). Returns wall seconds."""
    if len(meshes) == 1:
        return _reference_build(meshes[0], cores)
    inner = max(1, cores // len(meshes))
    t0 = time.perf_counter()
    threads = [threading.Thread(target=_reference_build, args=(m, inner)) for m in meshes]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    return time.perf_counter() - t0


def cpu_baseline_sample():
    """The compiled reference on a bounded sample of the workload; reported, not a target."""
    try:
        from oracle import clodfull, clodref

        if not clodfull.available():
            return {"value": None, "unit": "Mtris/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/libclodref_full.so missing"}
        m = _reference_sample_mesh()
        threads = os.cpu_count() or 1
        dt = _reference_build(m, threads)
        out = {"value": m.triangle_count / dt / 1e6, "unit": "Mtris/s", "cores": threads, "kind": "reference", "sample": _reference_sample_text(m, threads, dt)}
        if clodref.available():
            t0 = time.perf_counter()
            clodref.dag_build_timed(np.ascontiguousarray(m.positions), m.indices, np.ascontiguousarray(m.normals), ATTR_WEIGHTS, PROTECT_MASK, threads=threads)
            out["clodBuildEx_only_mtris_s"] = m.triangle_count / (time.perf_counter() - t0) / 1e6
        return out
    except Exception as e:  # pragma: no cover
        return {"value": None, "unit": "Mtris/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}


def parity_block(lib):
    """Our DAG next to the reference's on the CPU arm's sample mesh (inner boundary, normals as simplification attributes):
    per-level triangle delta, fallback-group counts and max-error ratio. tests/test_scale_parity.py asserts the same quantities
    on C1 / C2-sized meshes; tools/noise_floor.py measures how far the reference's own per-level max moves with its input order."""
    try:
        import importlib.util

        spec = importlib.util.spec_from_file_location("scale_parity", os.path.join(ROOT, "tools", "scale_parity.py"))
        sp = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(sp)
        m = _reference_sample_mesh()
        ours = sp.ours_stats(lib, m)
        ref = sp.ref_stats(m)
        rows, summary = sp.compare(ours, ref)
        summary["sample"] = f"{m.triangle_count}-triangle sample mesh of the workload generator (the CPU arm's sample), clodBuildEx boundary, normals x3 weight 1"
        summary["bars"] = {"triangles_per_level": "+-2 %", "max_error": "reference's own input-order spread is 1.2x-1.9x per level (profiles/r02_reference_noise_floor.txt)"}
        summary["levels_detail"] = [{k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k in ("depth", "triangles", "ref_triangles", "delta_pct", "sloppy", "ref_sloppy", "error_ratio")} for r in rows]
        return summary
    except Exception as e:  # pragma: no cover
        return {"error": str(e)}


def run_reference(args):
    rank, world, _ = _dist()
    if rank != 0:
        return
    if WORKLOAD in ("C4", "C5"):
        return run_reference_batch(args)
    n_meshes = max(1, args.gpus)  # the workload of the N-GPU arm is N independent meshes
    meshes = [_reference_sample_mesh(index=i) for i in range(n_meshes)]
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        _reference_step(meshes, cores)
    total = 0.0
    for _ in range(args.steps):
        total += _reference_step(meshes, cores)
    dt = total / args.steps
    tris = sum(m.triangle_count for m in meshes)
    value = tris / dt / 1e6
    sample = _reference_sample_text(meshes[0], cores if n_meshes == 1 else max(1, cores // n_meshes), dt, n_meshes)
    print(json.dumps({
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": "Mtris/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": dt * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32+u32",
        "data": "synthetic",
        "config": _config(WORKLOAD if WORKLOAD != "C4" else "C2", max(1, args.gpus)),
        "cpu_baseline": {"value": value, "unit": "Mtris/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "Mtris/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # exactly one line on stdout: libraries that write to fd 1 (NCCL's version banner) are sent to stderr for the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
