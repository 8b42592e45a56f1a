#!/usr/bin/env python
"""Benchmark of the cluster-LOD DAG build (BASELINE.json metric: Mtris/s, full DAG build).

  python bench.py --gpus N --steps K --warmup W            this framework (CUDA, one process per GPU)
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU implementation on the host cores

A step is one full DAG build of the workload mesh (N = 1: config C2, the 10 M-triangle displaced heightfield). Under
torchrun (N > 1) every rank builds its own shard of a scene batch of independent meshes (config C4 shape, weak scaling);
there is no data-path collective, only the barrier and the max-over-ranks of the timing.

`value` is timed with inputs resident in HBM; `e2e` goes through the host-pointer C ABI (clodb200_buildEx-shaped call:
upload + build + read-back of the whole callback stream) from pinned host buffers.
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
    # segmented SAH area scan, 4 elements per thread: order u32 + gathered box 32 B + node id u32 (+ area f32 out)
    "k_sa_chained": 4 * (4 + 32 + 4 + 4),
    # radix sort scatter, 8 keys per thread: key in + key out + value in + value out (u32 keys)
    "(k_rs_scatter<K>)": 8 * 16,
    "k_rs_scatter<K>": 8 * 16,
    "(k_rs_hist<K>)": 8 * 4,
    "k_rs_hist<K>": 8 * 4,
    "k_pivot_large": 4 + 4 + 4 + 8,
    "k_partition": 4 + 4 + 4 + 1 + 4 + 4,
    "k_side_flags": 4 + 1 + 4,
    "k_mark_sides": 4 + 4 + 1 + 8,
    "k_update_node_of_pos": 4 + 4 + 4,
    "k_wave_decide": 1 + 4 + 8 + 8 + 16,
    "k_wave_publish": 1 + 4 + 8 + 3 * 8,
    "k_rank": 9 + 2 * (44 + 12) + 4 + 8,
    "k_build_clusters": 128 * 3 * 4 * 2 + 4 * 4,
    "k_cluster_bounds": 128 * 3 * 16 + 16,
    # chained single-pass scan, 8 u32 elements per thread, read once + written once
    "(k_scan_chained<T, Op>)": 8 * 8,
}


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
        stop_event.wait(0.2)


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


def _workload(rank: int, world: int):
    """N = 1: the C2 heightfield. N > 1: mesh `rank` of a batch of equally sized heightfields with distinct noise seeds
    (weak scaling: per-GPU work is fixed; meshes are independent, the C4 sharding rule)."""
    seed = 1234 if world == 1 else 1234 + rank
    n = int(os.environ.get("CLODB200_BENCH_GRID", WORKLOAD_N))
    return meshgen.grid(n, seed=seed)


def _pin(a: np.ndarray) -> np.ndarray:
    import torch

    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    _PINNED.append(t)  # keep the pinned allocation alive for the numpy view
    return t.numpy()


_PINNED = []


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank, world, local = _dist()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)

    from basicrenderer_b200 import load, sharding

    lib = load(local)
    mesh = _workload(rank, world)
    T = mesh.triangle_count
    positions = _pin(mesh.positions.copy())
    normals = _pin(mesh.normals.copy())
    indices = _pin(mesh.indices)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-input timing (value)
    handle = lib.upload_mesh(positions, indices, attributes=normals, attribute_weights=ATTR_WEIGHTS, protect_mask=PROTECT_MASK)
    for _ in range(args.warmup):
        rec = lib.build_dag_resident(handle, keep_indices=False)
    stop = threading.Event()
    clock_samples = []
    sampler = threading.Thread(target=_sample_clocks, args=(stop, clock_samples), daemon=True)
    sampler.start()
    barrier()
    launches0 = lib.launch_count
    lib.timer_start()
    for _ in range(args.steps):
        rec = lib.build_dag_resident(handle, keep_indices=False)
        if world > 1:
            # the path's only exchange: per-mesh metadata to every rank (NCCL all-gather of sizes + padded blobs)
            gathered = sharding.gather_metadata([rank], [sharding.dag_summary_blob(rec)])
            assert len(gathered) == world
    ms = lib.timer_stop_ms()
    launches = lib.launch_count - launches0
    barrier()
    stop.set()
    sampler.join()

    # ---- end to end through the host-pointer C ABI (upload + build + read-back every step)
    for _ in range(min(args.warmup, 2)):
        lib.build_dag(positions, indices, attributes=normals, attribute_weights=ATTR_WEIGHTS, protect_mask=PROTECT_MASK, views=True)
    barrier()
    lib.timer_start()
    for _ in range(args.steps):
        rec_e2e = lib.build_dag(positions, indices, attributes=normals, attribute_weights=ATTR_WEIGHTS, protect_mask=PROTECT_MASK, views=True)
    ms_e2e = lib.timer_stop_ms()
    barrier()
    h2d = positions.nbytes + normals.nbytes + indices.nbytes
    d2h = int(rec_e2e.stats[4])

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
        lib.build_dag_resident(handle, keep_indices=False)
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
        roofline = {
            "bound": "hbm",
            "kernel": top[0],
            "achieved": achieved,
            "peak": peak,
            "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)",
            "unit": "GB/s",
            "frac": (achieved / peak) if achieved else None,
            "traffic": None,
            "kernel_share_of_step": top[2] / total_kernel_ms if total_kernel_ms else None,
            "launches": top[1],
            "avg_launch_us": top[2] * 1e3 / max(1, top[1]),
            "top_kernels": [{"kernel": r[0], "launches": r[1], "ms": round(r[2], 3)} for r in rows[:8]],
        }

        cpu = cpu_baseline_sample()
        ms_per_step = ms / args.steps
        value = total_tris / (ms_per_step * 1e-3) / 1e6
        e2e_value = total_tris / (ms_e2e / args.steps * 1e-3) / 1e6
        result = {
            "metric": "Mtris/s full cluster-LOD DAG build",
            "value": value,
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
            "config": {
                "workload": f"C2: {T}-triangle displaced heightfield grid (n={int(round((T / 2) ** 0.5))}), pos+normal, full DAG to a single root cluster" + ("" if world == 1 else f"; one such mesh per GPU ({world} independent meshes, sharded by mesh)"),
                "builder": "clodDefaultConfig(128) + BasicRenderer overrides (128/128/64, partition 384, 8 refined ids, spatial clusters)",
                "l2": "inputs and per-level working sets (>= 240 MB) exceed the 126 MB L2; no explicit flush",
                "levels": int(rec.levels),
                "groups": int(rec.groups),
                "clusters": int(rec.total_clusters),
                "scope": "clodBuildEx path (remap, clusterize, partition, lock, simplify, bounds, error rule, callback stream to host)",
            },
            "clocks": _clock_summary(clock_samples),
            "e2e": {"value": e2e_value, "unit": "Mtris/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": cpu,
        }
    lib.free_mesh(handle)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if result is not None:
        print(json.dumps(result))


def _reference_sample_mesh():
    # bounded sample of the workload: a 1 M-triangle tile of the same heightfield generator and density
    n = int(os.environ.get("CLODB200_REF_GRID", 707))
    return meshgen.grid(n, seed=1234)


def cpu_baseline_sample():
    """The compiled reference (oracle/_ref, clodBuildEx with its per-group tasks on all host threads) on a bounded
    sample of the workload; reported, not a target."""
    try:
        from oracle import clodref

        if not clodref.available():
            return {"value": None, "unit": "Mtris/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/libclodref.so missing"}
        m = _reference_sample_mesh()
        threads = os.cpu_count() or 1
        t0 = time.perf_counter()
        clodref.dag_build_timed(m.positions, m.indices, m.normals, ATTR_WEIGHTS, PROTECT_MASK, threads=threads)
        dt = time.perf_counter() - t0
        return {"value": m.triangle_count / dt / 1e6, "unit": "Mtris/s", "cores": threads, "kind": "reference",
                "sample": f"{m.triangle_count}-triangle tile of the C2 heightfield (same generator/density), reference clodBuildEx, {threads} threads, {dt:.2f} s"}
    except Exception as e:  # pragma: no cover
        return {"value": None, "unit": "Mtris/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}


def run_reference(args):
    rank, world, _ = _dist()
    if rank != 0:
        return
    from oracle import clodref

    m = _reference_sample_mesh()
    threads = os.cpu_count() or 1
    for _ in range(min(args.warmup, 1)):
        clodref.dag_build_timed(m.positions, m.indices, m.normals, ATTR_WEIGHTS, PROTECT_MASK, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        clodref.dag_build_timed(m.positions, m.indices, m.normals, ATTR_WEIGHTS, PROTECT_MASK, threads=threads)
    dt = (time.perf_counter() - t0) / args.steps
    value = m.triangle_count / dt / 1e6
    sample = f"{m.triangle_count}-triangle tile of the C2 heightfield per step (bounded sample of the 10 M workload; the reference is ~80 % serial so Mtris/s is size independent to first order)"
    print(json.dumps({
        "impl": "reference",
        "metric": "Mtris/s full cluster-LOD DAG build",
        "value": value,
        "unit": "Mtris/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": min(args.warmup, 1),
        "ms_per_step": dt * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32+u32",
        "data": "synthetic",
        "config": {"workload": "C2 heightfield, reference clodBuildEx (meshoptimizer v1.0 + clusterlod.h) on host cores", "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Mtris/s", "cores": threads, "kind": "reference", "sample": sample},
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
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
