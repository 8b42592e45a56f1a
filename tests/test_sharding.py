"""Multi-GPU host logic on CPU: mesh -> rank assignment and the metadata gather over a world_size-2 gloo group."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from basicrenderer_b200 import sharding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_assignment_covers_every_mesh_once_and_balances():
    rng = np.random.default_rng(7)
    counts = np.exp(rng.uniform(np.log(1e4), np.log(2e6), size=512)).astype(np.int64)
    for world in (1, 2, 4, 8):
        shards = sharding.assign_meshes(counts, world)
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(512))
        loads = [sum(sharding.mesh_cost(counts[i]) for i in s) for s in shards]
        assert max(loads) <= 1.02 * (sum(loads) / world) + sharding.mesh_cost(counts.max())
    assert sharding.assign_meshes([5, 5, 5], 2) == sharding.assign_meshes([5, 5, 5], 2)  # deterministic
    assert sharding.assign_meshes([], 4) == [[], [], [], []]


def test_pack_roundtrip():
    blobs = [b"", b"abc", bytes(range(256)) * 3]
    got = sharding.unpack_blobs(sharding.pack_blobs([4, 9, 2], blobs))
    assert got == {4: b"", 9: b"abc", 2: bytes(range(256)) * 3}


def test_gather_metadata_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent("""
        import os, sys, pickle
        sys.path.insert(0, "@ROOT@")
        import torch.distributed as dist
        from basicrenderer_b200 import sharding
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        counts = [100, 2000, 30, 400, 5000, 60, 7]
        mine = sharding.assign_meshes(counts, world)[rank]
        blobs = [bytes([i]) * (counts[i] % 97 + rank) for i in mine]
        merged = sharding.gather_metadata(mine, blobs)
        with open(os.path.join("@OUT@", f"out{rank}.pkl"), "wb") as f:
            pickle.dump((mine, merged), f)
        dist.barrier()
        dist.destroy_process_group()
    """).replace("@ROOT@", ROOT).replace("@OUT@", str(tmp_path)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r))) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=120) == 0
    import pickle

    results = [pickle.load(open(tmp_path / f"out{r}.pkl", "rb")) for r in range(2)]
    assert results[0][1] == results[1][1]  # every rank holds the merged table
    merged = results[0][1]
    assert sorted(merged) == list(range(7))
    counts = [100, 2000, 30, 400, 5000, 60, 7]
    for r in range(2):
        for i in results[r][0]:
            assert merged[i] == bytes([i]) * (counts[i] % 97 + r)
    assert sorted(results[0][0] + results[1][0]) == list(range(7))


def test_scene_batch_world_size_2_gloo_real_builds(tmp_path):
    """The N > 1 path of bench.py's scene-batch mode on CPU: meshes of the C4 generator sharded over two gloo ranks, every
    rank builds its meshes (kernel sources in host emulation) on two threads, serialises the cache metadata and the blobs
    are gathered; every rank must end up with one loadable blob per mesh, identical to a single-process build."""
    from basicrenderer_b200 import build

    emu = build.build_emu()
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent("""
        import os, sys, pickle
        from concurrent.futures import ThreadPoolExecutor
        sys.path.insert(0, "@ROOT@")
        import torch.distributed as dist
        from basicrenderer_b200 import artifacts as art, meshgen, sharding
        from basicrenderer_b200.api import ClodLib
        dist.init_process_group("gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        lib = ClodLib("@EMU@")
        budgets = meshgen.scene_batch_sizes(5, 30000, lo=2000, hi=12000)
        mine = sharding.assign_meshes([int(b) for b in budgets], world)[rank]
        def build_one(i):
            m = meshgen.scene_mesh(i, budgets[i])
            a = lib.build_artifacts(art.interleave(m.positions, m.normals), m.indices, art.VERTEX_NORMALS, keep_handle=True)
            blob = lib.serialize_metadata(a, f"clod_mesh{i}.clodbin", "scene", f"/mesh{i}")
            lib.free_artifacts(a)
            return blob
        with ThreadPoolExecutor(max_workers=2) as pool:
            blobs = list(pool.map(build_one, mine))
        merged = sharding.gather_metadata(mine, blobs)
        with open(os.path.join("@OUT@", f"out{rank}.pkl"), "wb") as f:
            pickle.dump((mine, merged), f)
        dist.barrier()
        dist.destroy_process_group()
    """).replace("@ROOT@", ROOT).replace("@OUT@", str(tmp_path)).replace("@EMU@", emu))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r))) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    import pickle

    from basicrenderer_b200 import artifacts as art
    from basicrenderer_b200 import cache, meshgen
    from basicrenderer_b200.api import ClodLib

    results = [pickle.load(open(tmp_path / f"out{r}.pkl", "rb")) for r in range(2)]
    assert results[0][1] == results[1][1] and sorted(results[0][0] + results[1][0]) == list(range(5))
    lib = ClodLib(emu)
    budgets = meshgen.scene_batch_sizes(5, 30000, lo=2000, hi=12000)
    for i, blob in results[0][1].items():
        meta = cache.read_metadata(blob)  # DeserializeMetadata acceptance rules
        assert meta["containerFileName"] == f"clod_mesh{i}.clodbin" and meta["primPath"] == f"/mesh{i}"
        m = meshgen.scene_mesh(i, budgets[i])
        a = lib.build_artifacts(art.interleave(m.positions, m.normals), m.indices, art.VERTEX_NORMALS, keep_handle=True)
        assert lib.serialize_metadata(a, f"clod_mesh{i}.clodbin", "scene", f"/mesh{i}") == blob
        lib.free_artifacts(a)
