"""Independent meshes can be built concurrently from several host threads of one process: every calling thread has its own
build context (stream, workspace arenas, scan-chain descriptors, pinned staging; csrc/rt.cuh). The reference builds one
primitive per worker thread in the same way (BasicRenderer/src/Import/GlTFGeometryExtractor.cpp:1349). Results must be
byte-identical to building the same meshes one after the other."""
import threading

import numpy as np


def _build(lib, m):
    from basicrenderer_b200 import artifacts as art

    h = lib.upload_geometry(art.interleave(m.positions, m.normals), m.indices, art.VERTEX_NORMALS)
    try:
        a = lib.build_artifacts_resident(h, views=False)
        return a.meshPages.tobytes(), a.groups.tobytes(), a.nodes.tobytes()
    finally:
        lib.free_geometry(h)


def test_concurrent_builds_match_serial_builds(lib):
    from basicrenderer_b200 import meshgen

    meshes = [meshgen.grid(40 + 7 * i, seed=i) if i % 2 else meshgen.icosphere(10 + 2 * i) for i in range(6)]
    serial = [_build(lib, m) for m in meshes]
    results = [None] * len(meshes)
    errors = []

    def work(k):
        try:
            for _ in range(2):  # a second build on the same thread reuses the thread's workspace
                results[k] = _build(lib, meshes[k])
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(meshes))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for k in range(len(meshes)):
        assert results[k] == serial[k], f"mesh {k} differs when built concurrently"
