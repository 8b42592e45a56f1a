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


def test_library_gather_loopback_and_errors(lib):
    """csrc/comm.cu with world size 1 (no NCCL involved): the gather returns this rank's payload; a second init is refused; the
    emulation refuses world sizes above 1 (it has no NCCL), the CUDA library needs a 128-byte id for them."""
    from basicrenderer_b200 import ClodbError

    lib.comm_init(None, 1, 0)
    try:
        h = sharding.gather_metadata_begin(lib, [3, 5], [b"abc", b"defgh"])
        assert sharding.gather_metadata_end(lib, h) == {3: b"abc", 5: b"defgh"}
        assert lib.gather_end(lib.gather_begin(b"")) == [b""]
    finally:
        lib.comm_destroy()
    import pytest

    with pytest.raises(ClodbError):
        lib.comm_init(None, 2, 0)  # more than one rank needs the unique id
    with pytest.raises(ClodbError):
        lib.comm_init(b"\0" * 128, 2, 5)  # rank out of range
    lib.comm_destroy()


@__import__("pytest").mark.gpu
def test_library_gather_two_ranks_nccl(tmp_path):
    """Two processes on one or two GPUs exchange payloads of different sizes through clodb200_comm* (NCCL all-gather over the
    library's own communicator); skipped when fewer than two GPUs are visible."""
    import pytest
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(textwrap.dedent("""
        import os, sys, pickle, time
        sys.path.insert(0, "@ROOT@")
        from basicrenderer_b200 import load, sharding
        rank = int(sys.argv[1])
        lib = load(rank)
        idf = os.path.join("@OUT@", "id.bin")
        if rank == 0:
            open(idf + ".tmp", "wb").write(lib.comm_unique_id()); os.replace(idf + ".tmp", idf)
        while not os.path.exists(idf):
            time.sleep(0.05)
        lib.comm_init(open(idf, "rb").read(), 2, rank)
        out = []
        for rnd in range(3):
            ids = [10 * rank + i for i in range(2 + rank)]
            blobs = [bytes([rank + 1]) * (1000 * (i + 1) + 7 * rnd + rank) for i in range(2 + rank)]
            out.append(sharding.gather_metadata_end(lib, sharding.gather_metadata_begin(lib, ids, blobs)))
        lib.comm_destroy()
        pickle.dump(out, open(os.path.join("@OUT@", f"out{rank}.pkl"), "wb"))
    """).replace("@ROOT@", ROOT).replace("@OUT@", str(tmp_path)))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    import pickle

    a, b = (pickle.load(open(tmp_path / f"out{r}.pkl", "rb")) for r in range(2))
    assert a == b and len(a) == 3
    for rnd, merged in enumerate(a):
        assert sorted(merged) == [0, 1, 10, 11, 12]
        assert merged[11] == bytes([2]) * (2000 + 7 * rnd + 1)
