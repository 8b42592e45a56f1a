"""include/clodb200_ingest.hpp: the C++ mirror of the reference's MeshIngestBuilder (ClusterLODTypes.h:354-434) above the C
ABI. A small C++ program feeds a mesh vertex by vertex as the importers do, checks the reference's error messages and
builds; its pages / groups / nodes must be byte-identical to the same build through the Python mirror."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _fnv1a(data: bytes) -> int:
    h = 1469598103934665603
    for chunk in (data[i:i + 1 << 16] for i in range(0, len(data), 1 << 16)):
        for b in chunk:
            h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def _run(lib_path: str, tmp_path, mesh):
    exe = str(tmp_path / "ingest_test")
    libdir, libname = os.path.dirname(lib_path), os.path.basename(lib_path)[3:-3]
    cuda = [f for d in ("/usr/local/cuda/lib64",) if os.path.isdir(d) for f in ("-Wl,-rpath-link," + d, "-Wl,-rpath," + d)]  # libcudart of the product library
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "ingest_test.cpp"), "-o", exe,
                           "-L", libdir, "-l" + libname, "-Wl,-rpath," + libdir] + cuda)
    path = str(tmp_path / "mesh.bin")
    v = np.ascontiguousarray(np.concatenate([mesh.positions, mesh.normals], axis=1), np.float32)
    with open(path, "wb") as f:
        f.write(np.array([v.shape[0], mesh.indices.size], np.uint32).tobytes())
        f.write(v.tobytes())
        f.write(np.ascontiguousarray(mesh.indices, np.uint32).tobytes())
    return subprocess.check_output([exe, path], text=True).splitlines()


def test_cpp_ingest_builder_matches_python_mirror(lib, tmp_path):
    from basicrenderer_b200 import artifacts as art
    from basicrenderer_b200 import meshgen

    mesh = meshgen.grid(48, seed=9)
    lines = _run(lib.path, tmp_path, mesh)
    fields = dict(zip(lines[0].split()[0::2], lines[0].split()[1::2]))
    assert fields["thrown"] == "2"  # both reference error messages reproduced
    ours = lib.build_artifacts(art.interleave(mesh.positions, mesh.normals), mesh.indices, art.VERTEX_NORMALS)
    assert int(fields["groups"]) == len(ours.groups) and int(fields["nodes"]) == len(ours.nodes) and int(fields["segments"]) == len(ours.segments)
    assert int(fields["pages_fnv"], 16) == _fnv1a(np.asarray(ours.meshPages).tobytes())
    assert int(fields["groups_fnv"], 16) == _fnv1a(np.asarray(ours.groups).tobytes())
    assert int(fields["nodes_fnv"], 16) == _fnv1a(np.asarray(ours.nodes).tobytes())
    assert lines[1] == "empty groups 0"
