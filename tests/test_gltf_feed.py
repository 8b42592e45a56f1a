"""Importer boundary: a glTF file fed through tools/gltf_cache_tool.py the way the reference's extractor
feeds its builder (GlTFGeometryExtractor.cpp:1008-1298), including the skip-if-cached resume. The .glb is written by the test from
synthetic meshes: one primitive with normals and two UV sets (u16 indices, interleaved buffer view), one without normals (u32)."""
import importlib.util
import json
import os
import struct

import numpy as np

from basicrenderer_b200 import artifacts as art
from basicrenderer_b200 import cache, meshgen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("gltf_cache_tool", os.path.join(ROOT, "tools", "gltf_cache_tool.py"))
tool = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(tool)


def _write_glb(path, prims):
    blob = bytearray()
    views, accessors, primitives = [], [], []

    def view(data: bytes, stride=None):
        while len(blob) % 4:
            blob.append(0)
        v = {"buffer": 0, "byteOffset": len(blob), "byteLength": len(data)}
        if stride:
            v["byteStride"] = stride
        views.append(v)
        blob.extend(data)
        return len(views) - 1

    def accessor(view_index, ctype, count, kind, offset=0):
        accessors.append({"bufferView": view_index, "byteOffset": offset, "componentType": ctype, "count": count, "type": kind})
        return len(accessors) - 1

    for p in prims:
        attrs = {}
        V = p["positions"].shape[0]
        if p.get("interleaved"):
            inter = np.concatenate([p["positions"], p["normals"], p["uv0"]], axis=1).astype(np.float32)
            vi = view(inter.tobytes(), stride=32)
            attrs["POSITION"] = accessor(vi, 5126, V, "VEC3", 0)
            attrs["NORMAL"] = accessor(vi, 5126, V, "VEC3", 12)
            attrs["TEXCOORD_0"] = accessor(vi, 5126, V, "VEC2", 24)
            attrs["TEXCOORD_1"] = accessor(view(p["uv1"].astype(np.float32).tobytes()), 5126, V, "VEC2")
        else:
            attrs["POSITION"] = accessor(view(p["positions"].astype(np.float32).tobytes()), 5126, V, "VEC3")
        idx = p["indices"]
        if idx.max() < 65536 and p.get("interleaved"):
            ii = accessor(view(idx.astype(np.uint16).tobytes()), 5123, idx.size, "SCALAR")
        else:
            ii = accessor(view(idx.astype(np.uint32).tobytes()), 5125, idx.size, "SCALAR")
        primitives.append({"attributes": attrs, "indices": ii, "mode": 4})
    doc = {"asset": {"version": "2.0"}, "buffers": [{"byteLength": len(blob)}], "bufferViews": views, "accessors": accessors,
           "meshes": [{"primitives": primitives[:1]}, {"primitives": primitives[1:]}]}
    js = json.dumps(doc).encode()
    js += b" " * (-len(js) % 4)
    blob.extend(b"\0" * (-len(blob) % 4))
    with open(path, "wb") as f:
        f.write(struct.pack("<III", 0x46546C67, 2, 12 + 8 + len(js) + 8 + len(blob)))
        f.write(struct.pack("<II", len(js), 0x4E4F534A) + js)
        f.write(struct.pack("<II", len(blob), 0x004E4942) + bytes(blob))


def test_gltf_file_builds_caches_and_resumes(lib, tmp_path):
    a = meshgen.icosphere(12, True, True)  # seams: positions repeat with different uvs
    b = meshgen.grid(48, seed=9)
    uv1 = np.random.default_rng(1).random((a.vertex_count, 2)).astype(np.float32)
    glb = str(tmp_path / "scene v1.glb")
    _write_glb(glb, [{"positions": a.positions, "normals": a.normals, "uv0": a.vertices[:, 6:8], "uv1": uv1, "indices": a.indices, "interleaved": True},
                     {"positions": b.positions, "indices": b.indices}])
    root = str(tmp_path / "cache")
    first = tool.run(glb, root, lib)
    assert [r[1] for r in first] == ["built", "built"] and [r[0] for r in first] == ["/glTF/Mesh/0/Primitive/0", "/glTF/Mesh/1/Primitive/0"]
    assert first[0][2] == a.triangle_count and first[1][2] == b.triangle_count
    source = os.path.normpath(glb).replace("\\", "/")
    directory = os.path.join(root, cache.cache_subdirectory(source))
    assert os.path.basename(directory).startswith("scene_v1_")
    # the cache of the first primitive holds what a direct build of the same streams gives
    direct = lib.build_artifacts(np.concatenate([a.positions, a.normals, a.vertices[:, 6:8]], axis=1), a.indices, art.VERTEX_NORMALS | art.VERTEX_TEXCOORDS,
                                 uv_sets=[a.vertices[:, 6:8], uv1])
    pages = cache.read_container(os.path.join(directory, first[0][3] + ".clodbin"))
    assert len(pages) == direct.page_count and all(np.array_equal(np.frombuffer(p, np.uint8), direct.page(i)) for i, p in enumerate(pages))
    meta = cache.read_metadata(open(os.path.join(directory, first[0][3] + ".clodblob"), "rb").read())
    assert meta["sourceIdentifier"] == source and meta["primPath"] == "/glTF/Mesh/0/Primitive/0" and meta["buildConfigHash"] == cache.build_config_hash()
    # resume: a second run finds both caches; removing one container rebuilds only that primitive
    assert [r[1] for r in tool.run(glb, root, lib)] == ["cached", "cached"]
    os.remove(os.path.join(directory, first[1][3] + ".clodbin"))
    assert [r[1] for r in tool.run(glb, root, lib)] == ["cached", "built"]
