"""Scene batches across GPUs: meshes are independent units (reference: one task per primitive,
BasicRenderer/src/Import/GlTFGeometryExtractor.cpp:1349), so a batch is sharded by mesh with no data-path collective.
The only exchange is the gather of the per-mesh metadata blobs; page/cluster payloads stay on the rank that built
them. On GPUs the gather is the library's own NCCL all-gather (csrc/comm.cu: gather_metadata_begin / gather_metadata_end below);
gather_metadata is the torch.distributed version kept for the world-size-2 gloo tests on CPU."""
from __future__ import annotations

import math
import struct
from typing import List, Sequence

import numpy as np


def mesh_cost(triangles: int) -> float:
    """Build cost model: the DAG build is O(T log T) (radix sorts + ~log2(T/128) tree and DAG levels)."""
    t = max(int(triangles), 1)
    return t * max(1.0, math.log2(t))


def assign_meshes(triangle_counts: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of mesh ids to ranks. Deterministic (ties by mesh id), every mesh appears
    exactly once; a single mesh never spans ranks."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(triangle_counts)), key=lambda i: (-mesh_cost(triangle_counts[i]), i))
    loads = [0.0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += mesh_cost(triangle_counts[i])
    for s in shards:
        s.sort()
    return shards


def pack_blobs(mesh_ids: Sequence[int], blobs: Sequence[bytes]) -> bytes:
    """[u32 count] then per mesh [u32 mesh id][u64 size][bytes]."""
    out = [struct.pack("<I", len(mesh_ids))]
    for i, b in zip(mesh_ids, blobs):
        out.append(struct.pack("<IQ", int(i), len(b)))
        out.append(bytes(b))
    return b"".join(out)


def unpack_blobs(buf: bytes):
    (n,) = struct.unpack_from("<I", buf, 0)
    off = 4
    out = {}
    for _ in range(n):
        i, size = struct.unpack_from("<IQ", buf, off)
        off += 12
        out[i] = bytes(buf[off : off + size])
        off += size
    return out


def gather_metadata(mesh_ids: Sequence[int], blobs: Sequence[bytes], device=None):
    """Gathers every rank's (mesh id -> metadata blob) to all ranks and returns the merged dict ordered by mesh id.
    Variable sizes: one all_gather of byte counts, then one all_gather of buffers padded to the largest."""
    import torch
    import torch.distributed as dist

    payload = pack_blobs(mesh_ids, blobs)
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(sorted(unpack_blobs(payload).items()))
    world = dist.get_world_size()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    size = torch.tensor([len(payload)], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(size) for _ in range(world)]
    dist.all_gather(sizes, size)
    max_size = max(int(s.item()) for s in sizes)
    buf = torch.zeros(max_size, dtype=torch.uint8, device=device)
    buf[: len(payload)] = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(device)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    merged = {}
    for s, b in zip(sizes, bufs):
        merged.update(unpack_blobs(b[: int(s.item())].cpu().numpy().tobytes()))
    return dict(sorted(merged.items()))


def gather_metadata_begin(lib, mesh_ids: Sequence[int], blobs: Sequence[bytes]):
    """The product path on GPUs: the library's own NCCL all-gather (csrc/comm.cu, clodb200_commGatherBegin) on its communication
    stream. Returns a handle; the caller keeps building and collects the merged dict later with gather_metadata_end."""
    return lib.gather_begin(pack_blobs(mesh_ids, blobs))


def gather_metadata_end(lib, handle):
    merged = {}
    for payload in lib.gather_end(handle):
        if payload:
            merged.update(unpack_blobs(payload))
    return dict(sorted(merged.items()))


def dag_summary_blob(rec) -> bytes:
    """Compact per-mesh summary of a recorded DAG build (clodBuildEx level): per group {depth, simplified bounds[5], cluster
    count}. The artifacts path gathers the reference's full metadata blob instead (ClodLib.serialize_metadata,
    CLodCache.cpp:169-207)."""
    depth = np.asarray(rec.group_depth, np.int32)
    simp = np.asarray(rec.group_simplified, np.float32).reshape(-1, 5)
    counts = np.diff(np.asarray(rec.group_cluster_offsets, np.uint32)).astype(np.uint32)
    head = struct.pack("<III", len(depth), int(rec.total_clusters), int(rec.levels))
    return head + depth.tobytes() + simp.tobytes() + counts.tobytes()
