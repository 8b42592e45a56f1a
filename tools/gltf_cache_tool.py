"""CLodCacheTool-shaped driver over the C ABI: reads a glTF 2.0 file (.gltf + buffers, or .glb), and for
every mesh primitive feeds the builder the way the reference's extractor does (BasicRenderer/src/Import/GlTFGeometryExtractor.cpp:
1008-1298): cache identity {normalised source path, "/glTF/Mesh/<m>/Primitive/<p>", ""} (:1013-1015), skip-if-cached probe
(:1025, CLodCacheLoader::TryLoadPrebuilt), interleaved MeshVertexLayout stream (position, normal, TEXCOORD_0), u32 indices, extra
TEXCOORD sets as separate UV sets, smooth normals generated when the primitive has none (:1072-1077), then
clodb200_buildArtifacts and clodb200_artifactsSaveCache under the reference's cache names.

  python tools/gltf_cache_tool.py scene.glb [--cache-root DIR] [--device 0] [--emu]

Scope: triangle primitives with float POSITION / NORMAL / TEXCOORD_n and u8/u16/u32 indices, sparse accessors not supported;
materials, nodes, skins and animations are the renderer's business. The `.usdc` wrapper of the blob is tools/wrap_usdc.py."""
import argparse
import base64
import json
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicrenderer_b200 import artifacts as art  # noqa: E402
from basicrenderer_b200 import cache  # noqa: E402

_COMPONENT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_WIDTH = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4}


class GltfDocument:
    def __init__(self, path: str):
        self.path = path
        data = open(path, "rb").read()
        self.glb_bin = None
        if data[:4] == b"glTF":
            _, version, length = struct.unpack_from("<III", data, 0)
            off = 12
            chunks = []
            while off < length:
                n, kind = struct.unpack_from("<II", data, off)
                chunks.append((kind, data[off + 8: off + 8 + n]))
                off += 8 + n
            self.json = json.loads(next(c for k, c in chunks if k == 0x4E4F534A))
            self.glb_bin = next((c for k, c in chunks if k == 0x004E4942), None)
        else:
            self.json = json.loads(data)
        self._buffers = {}

    def buffer(self, i: int) -> bytes:
        if i not in self._buffers:
            b = self.json["buffers"][i]
            uri = b.get("uri")
            if uri is None:
                self._buffers[i] = self.glb_bin
            elif uri.startswith("data:"):
                self._buffers[i] = base64.b64decode(uri.split(",", 1)[1])
            else:
                self._buffers[i] = open(os.path.join(os.path.dirname(self.path), uri), "rb").read()
        return self._buffers[i]

    def accessor(self, i: int) -> np.ndarray:
        a = self.json["accessors"][i]
        if "sparse" in a:
            raise ValueError("sparse accessors are not supported")
        view = self.json["bufferViews"][a["bufferView"]]
        dtype = np.dtype(_COMPONENT[a["componentType"]])
        width = _WIDTH[a["type"]]
        start = view.get("byteOffset", 0) + a.get("byteOffset", 0)
        stride = view.get("byteStride") or dtype.itemsize * width
        raw = self.buffer(view["buffer"])
        out = np.lib.stride_tricks.as_strided(np.frombuffer(raw, dtype, offset=start, count=((a["count"] - 1) * stride) // dtype.itemsize + width),
                                              shape=(a["count"], width), strides=(stride, dtype.itemsize)).copy()
        if a.get("normalized") and dtype.kind in "iu":
            out = out.astype(np.float32) / float(np.iinfo(dtype).max)
        return out


def smooth_normals(positions: np.ndarray, indices: np.ndarray) -> np.ndarray:
    """area-weighted vertex normals (ComputeSmoothNormals of the extractor)"""
    tri = indices.reshape(-1, 3)
    p = positions.astype(np.float64)
    fn = np.cross(p[tri[:, 1]] - p[tri[:, 0]], p[tri[:, 2]] - p[tri[:, 0]])
    n = np.zeros_like(p)
    for k in range(3):
        np.add.at(n, tri[:, k], fn)
    length = np.linalg.norm(n, axis=1, keepdims=True)
    n = np.where(length > 1e-20, n / np.maximum(length, 1e-20), np.array([[0.0, 0.0, 1.0]]))
    return n.astype(np.float32)


def primitives(doc: GltfDocument):
    """yields (mesh index, primitive index, vertices [V, 6 or 8] f32, indices u32, flags, extra uv sets)"""
    for mi, mesh in enumerate(doc.json.get("meshes", [])):
        for pi, prim in enumerate(mesh["primitives"]):
            if prim.get("mode", 4) != 4 or "POSITION" not in prim["attributes"]:
                continue
            attrs = prim["attributes"]
            pos = doc.accessor(attrs["POSITION"]).astype(np.float32)
            idx = doc.accessor(prim["indices"]).astype(np.uint32).reshape(-1) if "indices" in prim else np.arange(pos.shape[0], dtype=np.uint32)
            idx = idx[: idx.size // 3 * 3]
            nrm = doc.accessor(attrs["NORMAL"]).astype(np.float32) if "NORMAL" in attrs else smooth_normals(pos, idx)
            flags = art.VERTEX_NORMALS
            cols = [pos, nrm]
            sets = sorted(int(k.split("_")[1]) for k in attrs if k.startswith("TEXCOORD_"))
            uv_sets = []
            if sets:
                flags |= art.VERTEX_TEXCOORDS
                all_sets = [np.zeros((pos.shape[0], 2), np.float32) for _ in range(sets[-1] + 1)]
                for s in sets:
                    all_sets[s] = doc.accessor(attrs[f"TEXCOORD_{s}"]).astype(np.float32)
                cols.append(all_sets[0])
                uv_sets = all_sets
            yield mi, pi, np.ascontiguousarray(np.concatenate(cols, axis=1)), idx, flags, uv_sets


def run(path: str, cache_root: str, lib) -> list:
    """-> [(prim path, 'cached' | 'built', triangles, cache stem)] for every primitive of the file"""
    doc = GltfDocument(path)
    source = os.path.normpath(path).replace("\\", "/")
    config_hash = cache.build_config_hash()
    directory = os.path.join(cache_root, cache.cache_subdirectory(source))
    os.makedirs(directory, exist_ok=True)
    report = []
    for mi, pi, vertices, indices, flags, uv_sets in primitives(doc):
        prim_path = f"/glTF/Mesh/{mi}/Primitive/{pi}"  # GlTFGeometryExtractor.cpp:1014
        stem = cache.cache_file_name(source, prim_path, "", config_hash)[: -len(".usdc")]
        if lib.cache_probe(directory, stem + ".clodblob", source, prim_path, "", config_hash):
            report.append((prim_path, "cached", indices.size // 3, stem))
            continue
        a = lib.build_artifacts(vertices, indices, flags, uv_sets=uv_sets or None, keep_handle=True)
        lib.save_cache(a, directory, stem + ".clodbin", stem + ".clodblob", source, prim_path, "", config_hash)
        lib.free_artifacts(a)
        report.append((prim_path, "built", indices.size // 3, stem))
    return report


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("file")
    ap.add_argument("--cache-root", default="cache")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--emu", action="store_true", help="development emulation instead of the CUDA library")
    a = ap.parse_args()
    if a.emu:
        from basicrenderer_b200 import build
        from basicrenderer_b200.api import ClodLib

        lib = ClodLib(build.build_emu())
    else:
        from basicrenderer_b200 import load

        lib = load(a.device)
    for prim_path, what, tris, stem in run(a.file, a.cache_root, lib):
        print(f"{what:7} {prim_path}  {tris} triangles  {stem}")


if __name__ == "__main__":
    main()
